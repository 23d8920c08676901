"""Bug hunt, not a test: many random settings of LU's tunables on random bases, CUDA path vs the oracle
(tests/parity.py:tunables_case).  usage: python scripts/tunables_hunt.py [first_seed] [count] [max_m]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from blu_b200 import BLU  # noqa: E402
from parity import tunables_case  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 300
max_m = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
bad, t0 = 0, time.time()
for seed in range(first, first + count):
    m = 30 + (seed * 37) % max_m
    try:
        tunables_case(lambda m, nnz: BLU(m, nnz), m, seed, nupd=15, dens=2.5 + (seed % 5))
    except AssertionError as e:
        bad += 1
        print("FAIL seed", seed, "m", m, str(e)[:400], flush=True)
        if bad > 8:
            break
print(f"{count} cases, {bad} failures, {time.time() - t0:.1f} s")
