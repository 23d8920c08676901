"""Bug hunt, not a test: random batches through the split factorization (head / tail / build launches) under random
orders of the two dense-tail stages, CTA sizes and tunables, CUDA path vs the oracle basis by basis.
usage: python scripts/batch_hunt.py [first_seed] [count] [--emu]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from blu_b200 import BLUBatch, gen, load_library  # noqa: E402
from oracle_lib import Oracle  # noqa: E402
from parity import STATS  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
first = int(args[0]) if len(args) > 0 else 41000
count = int(args[1]) if len(args) > 1 else 24
emu = "--emu" in sys.argv
lib = load_library(os.path.join(ROOT, "tests", "emu", "libblu_emu.so")) if emu else None
bad, t0 = 0, time.time()
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    m = int(rng.integers(120, 400)) if emu else int(rng.integers(300, 1600))
    nmat = 3 if emu else int(rng.integers(8, 80))
    nslack = int(m * rng.uniform(0.15, 0.45))
    dens = float(rng.uniform(3.5, 6.5))
    kd = int(rng.choice([32, 64, 96] if emu else [64, 96, 128, 160]))
    kbig = int(rng.choice([0, kd + 32, kd + 64, 256 if not emu else kd + 96]))
    tail = int(rng.choice([128, 256] if emu else [256, 512, 1024]))
    nt = int(rng.choice([64, 128] if emu else [64, 128, 256]))
    ms = int(rng.integers(1, 6))
    droptol = float(rng.choice([1e-20, 1e-12, 1e-3]))
    bb, be, bi, bx, rhs = gen.batch(nmat, m, nslack, dens, seed * 100, seed * 100 + 50000)
    cap = int((be - bb).reshape(nmat, m).sum(1).max())
    b = BLUBatch(nmat, m, cap, lib=lib) if emu else BLUBatch(nmat, m, cap)
    b.dense_k = kd; b.dense_k_big = kbig; b.tail_threads = tail; b.threads_per_basis = nt; b.split_min = 0
    b.maxsearch = ms; b.droptol = droptol
    desc = f"seed {seed} m {m} nmat {nmat} kd {kd} kbig {kbig}->{int(b.get_param('dense_k_big'))} tail {tail} nt {nt} maxsearch {ms} droptol {droptol}"
    try:
        st, status = b.factorize(bb, be, bi, bx)
        assert st == 0, st
        st, x, sst = b.solve_dense(rhs, "N")
        assert st == 0
        stages = 0
        for k in range(0, nmat, max(1, nmat // 6)):
            cp, ri, v = gen.basis(seed * 100 + k, m, nslack, dens)
            o = Oracle(m, 400 * len(v))
            o.set_param("maxsearch", ms); o.set_param("droptol", droptol)
            so = o.factorize(cp[:-1], cp[1:], ri, v)
            assert so == status[k], (k, so, status[k])
            if so not in (0, 2):
                continue
            _, fo = o.get_factors()
            _, fg = b.get_factors(k)
            for key in fo:
                assert np.array_equal(fo[key], fg[key]), (k, key)
            for name in STATS:
                assert o.info(name) == b.info(k, name), (k, name)
            _, xo = o.solve_dense(rhs[k * m:(k + 1) * m], "N")
            assert np.array_equal(x[k], xo), (k, "solve_dense")
            stages = max(stages, int(b.info(k, "n_kind6")))
        print("ok  ", desc, "dense entries+restages", stages, flush=True)
    except AssertionError as e:
        bad += 1
        print("FAIL", desc, str(e)[:300], flush=True)
    b.close()
print(f"{count} batches, {bad} failures, {time.time() - t0:.1f} s")
