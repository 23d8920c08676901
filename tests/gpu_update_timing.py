"""Ad-hoc: distribution of the per-call times of the update replay (not a test)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blu_b200 import BLU, gen
m, bump, nupd = 20000, 1000, 60
cp, ri, v = gen.config3(m, bump, seed=7001)
pcp, pri, pv = gen.column_pool(7002, m, nupd)
dense = bump * bump
mem = (int(2.2 * (dense // 2 + 10 * m)) + 40 * m, int(1.2 * (dense // 2 + 10 * m)) + 40 * m, int(3.0 * dense + 40 * m))
g = BLU(m, len(v)); g.threads_per_basis = 1024
g.l_mem, g.u_mem, g.w_mem = mem
assert g.factorize(cp[:-1], cp[1:], ri, v) == 0
T = [[], [], []]
for it in range(nupd):
    idx, val = pri[pcp[it]:pcp[it + 1]], pv[pcp[it]:pcp[it + 1]]
    t = time.perf_counter(); assert g.solve_for_update(len(idx), idx, val, "N", 1) == 0; T[0].append(time.perf_counter() - t)
    j = int(np.argmax(np.abs(g.lhs))); xtbl = g.lhs[j]
    t = time.perf_counter(); assert g.solve_for_update(1, np.array([j]), None, "T", 0) == 0; T[1].append(time.perf_counter() - t)
    t = time.perf_counter(); st = g.update(xtbl); T[2].append(time.perf_counter() - t)
    assert st == 0
for name, t in zip(("ftran", "btran", "update"), T):
    t = np.array(t) * 1e3
    print(f"{name}: median {np.median(t):.2f} ms  mean {t.mean():.2f}  max {t.max():.2f}  min {t.min():.2f}; first 10: {np.round(t[:10], 1)}")
print("nrealloc", g.info("nrealloc"), "ngarbage", g.info("ngarbage"), "u_nz", g.info("u_nz"), "r_nz", g.info("r_nz"), "mem", g.get_param("l_mem"), g.get_param("u_mem"), g.get_param("w_mem"))
