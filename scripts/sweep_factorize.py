"""Tuning sweep, not a test: k_factorize of the configs[1] batch for several dense-tail orders and CTA sizes.
usage: python scripts/sweep_factorize.py [nmat] [dense_k,dense_k,...] [threads,threads,...] [tail_threads,...] [dense_k_big,...]
Prints ms per launch and the mean per-phase SM cycles of a sample of bases (BluInfo.t_phase)."""
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blu_b200 import BLUBatch, gen  # noqa: E402

nmat = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kds = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 128, 256, 384, 512]
nts = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [128]
tails = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [512]
kbigs = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [-1]
M = 2000
t0 = time.time()
bb, be, bi, bx, rhs = gen.batch(nmat, M, 700, 5.0, 2000, 3000)
print(f"generated {nmat} bases in {time.time() - t0:.1f} s", flush=True)
cap = int((be - bb).reshape(nmat, M).sum(1).max())
names = ["validate", "singl", "setup", "search", "p_srow", "p_scol", "p_dbl", "p_small", "p_any", "build", "remove", "total", "dense", "dsearch+convert", "d_gather", "d_sweep"]
for nt, kd, tail, kbig in [(a, b_, c, e) for a in nts for b_ in kds for c in (tails if b_ else tails[:1]) for e in (kbigs if b_ else kbigs[:1])]:
    if True:
        b = BLUBatch(nmat, M, cap, device=0)
        b.threads_per_basis = nt
        b.dense_k = kd
        b.tail_threads = tail
        if kbig >= 0:
            b.dense_k_big = kbig
        if os.environ.get("SWEEP_TREE_MIN"):
            b.tree_min = int(os.environ["SWEEP_TREE_MIN"])
        if os.environ.get("SWEEP_W_MEM"):
            b.w_mem = int(os.environ["SWEEP_W_MEM"])
        elif not os.environ.get("SWEEP_DEFAULT_MEM"):
            b.l_mem = 100000; b.u_mem = 100000; b.w_mem = 900000
        assert b.upload(bb, be, bi, bx, rhs) == 0
        ms = []
        for it in range(3):
            assert b.factorize_resident() == 0
            ms.append(b.last_kernel_ms(0) - b.last_kernel_ms(2))
        sample = range(0, nmat, max(1, nmat // 64))
        ph = np.array([[b.info(k, f"t_phase{q}") for q in range(16)] for k in sample]).mean(0)
        kinds = np.array([[b.info(k, f"n_kind{q}") for q in range(8)] for k in sample]).mean(0)
        bad = sum(int(b.info(k, "status")) != 0 for k in sample)
        gc = np.mean([b.info(k, "ngarbage") for k in sample]); nre = int(b.info(0, "nrealloc"))
        print(f"nt {nt} dense_k {kd} big {int(b.get_param('dense_k_big'))} tail {tail}: k_factorize {min(ms):.1f} ms (runs {[round(x, 1) for x in ms]}), norms {b.last_kernel_ms(2):.1f} ms, bad {bad}, head/tail/build {b.last_kernel_ms(3):.1f}/{b.last_kernel_ms(4):.1f}/{b.last_kernel_ms(5):.1f} ms, w_mem {int(b.get_param('w_mem'))} gc/basis {gc:.2f} realloc rounds {nre}", flush=True)
        print("   kcycles/basis: " + " ".join(f"{n}={v / 1e3:.0f}" for n, v in zip(names, ph)), flush=True)
        print("   pivots/basis: srow %.1f scol %.1f dbl %.1f small %.1f any %.1f | dense steps %.1f entries %.2f d_finish kcycles %.0f" % tuple(list(kinds[:7]) + [kinds[7] / 1e3]), flush=True)
        b.close()
