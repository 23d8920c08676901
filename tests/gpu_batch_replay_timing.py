"""Ad-hoc: throughput of column replacements over a batch of bases (not a test)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from blu_b200 import BLUBatch, gen
from oracle_lib import Oracle
nmat, m, rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 2000, 10
bb, be, bi, bx, rhs = gen.batch(nmat, m, 700, 5.0, 2000, 3000)
b = BLUBatch(nmat, m, int((be - bb).reshape(nmat, m).sum(1).max()))
b.l_mem = 300000; b.u_mem = 300000; b.w_mem = 900000
st, status = b.factorize(bb, be, bi, bx)
assert st == 0 and (status == 0).all()
pool = gen.column_pool(7002, m, rounds, nnz_col=6)
ncheck = 4
oracles = []
for k in range(ncheck):
    cp, ri, v = gen.basis(2000 + k, m, 700, 5.0)
    o = Oracle(m, 400 * len(v)); o.set_param("check_file_diff", 0); assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
    oracles.append(o)
tg = to = 0.0
same = True
for it in range(rounds):
    col = (pool[1][pool[0][it]:pool[0][it + 1]], pool[2][pool[0][it]:pool[0][it + 1]])
    t = time.perf_counter()
    st, stat, out = b.solve_for_update([col] * nmat, "N", want_solution=1)
    tg += time.perf_counter() - t
    assert st == 0, (it, st, stat[:8])
    # leaving column per basis from the device's own solution (argmax |x|), as maxvolume does
    leave, xt = [], []
    for k in range(nmat):
        il, xl = out[k]
        q = int(np.argmax(np.abs(xl))); leave.append((np.array([il[q]]), None)); xt.append(xl[q])
    t = time.perf_counter()
    st, stat, _ = b.solve_for_update(leave, "T", want_solution=0)
    st2, stat2 = b.update(np.array(xt))
    tg += time.perf_counter() - t
    assert st == 0 and st2 in (0, -6), (it, st, st2)
    for k, o in enumerate(oracles):
        t = time.perf_counter()
        assert o.solve_for_update(len(col[0]), col[0], col[1], "N", want_solution=1) == 0
        n = o.nzlhs
        same = same and np.array_equal(out[k][0], o.ilhs[:n]) and np.array_equal(out[k][1], o.lhs[o.ilhs[:n]])
        assert o.solve_for_update(1, leave[k][0], None, "T", want_solution=0) == 0
        so = o.update(xt[k])
        to += time.perf_counter() - t
        same = same and so == stat2[k] and b.info(k, "nforrest") == o.info("nforrest") and b.info(k, "pivot_error") == o.info("pivot_error")
print(f"batch of {nmat} bases (m={m}), {rounds} rounds: GPU {1e3 * tg / rounds:.1f} ms per round = {1e6 * tg / rounds / nmat:.1f} us per replacement per basis "
      f"(host-side argmax excluded); CPU oracle {1e6 * to / rounds / ncheck:.0f} us per replacement (1 thread); sampled bases bit-identical: {same}", flush=True)
