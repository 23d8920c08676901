"""configs[2] (100,000 rows, bump 3,995) factorize for several dense-tail orders and store sizes (tuning, not a test)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blu_b200 import BLU, gen  # noqa: E402
m, bump = 100000, 4000
cp, ri, v = gen.config3(m, bump)
dense = bump * bump
big = (int(2.2 * (dense // 2 + 10 * m)), int(1.2 * (dense // 2 + 10 * m)), int(3.0 * dense + 40 * m))
names = ["validate", "singl", "setup", "search", "p_srow", "p_scol", "p_dbl", "p_small", "p_any", "build", "remove", "total", "dense", "dsearch+conv", "d_gather", "d_sweep"]
for kd in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,160,256").split(",")]:
    for mem in (None, big):
        g = BLU(m, len(v)); g.threads_per_basis = 1024; g.dense_k = kd
        if mem:
            g.l_mem, g.u_mem, g.w_mem = mem
        g.factorize(cp[:-1], cp[1:], ri, v)
        t = time.perf_counter(); st = g.factorize(cp[:-1], cp[1:], ri, v); dt = time.perf_counter() - t
        print(f"dense_k {kd} stores {'explicit' if mem else 'default'}: status {st} factorize {1e3 * dt:.0f} ms wall, gc {int(g.info('ngarbage'))}, realloc {int(g.info('nrealloc'))}, "
              + " ".join(f"{n}={g.info(f't_phase{q}') / 1e6:.0f}" for q, n in enumerate(names)) + " Mcycles", flush=True)
