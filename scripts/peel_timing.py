"""Singleton peel before/after: phase time (BluInfo.t_phase[1], SM cycles of thread 0) of the
level-synchronous parallel peel vs the serial warp-0 queue it replaced, on configs[2] (100,000 rows, 96,000
singletons) and on a configs[1] basis.  usage: python scripts/peel_timing.py  (BLU_B200_LIB selects the library)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blu_b200 import BLU, gen  # noqa: E402

for name, (cp, ri, v), m, nt in (("configs[2] 100k rows", gen.config3(100000, 4000), 100000, 1024),
                                 ("configs[1] basis 0", gen.config2_matrix(0)[0], 2000, 128)):
    g = BLU(m, len(v))
    g.threads_per_basis = nt
    g.factorize(cp[:-1], cp[1:], ri, v)
    t = time.perf_counter()
    st = g.factorize(cp[:-1], cp[1:], ri, v)
    dt = time.perf_counter() - t
    print(f"{os.environ.get('BLU_B200_LIB', 'libblu_b200.so (parallel peel)')}: {name}: status {st}, rank after peel = m - bump = {int(m - g.info('bump_size'))}, "
          f"singleton phase {g.info('t_phase1') / 1e3:.0f} kcycles = {g.info('t_phase1') / 1.965e6:.2f} ms at 1965 MHz, validate+transpose {g.info('t_phase0') / 1e3:.0f} kcycles, "
          f"whole factorize {1e3 * dt:.1f} ms wall", flush=True)
