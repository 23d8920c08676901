"""Tuning sweep, not a test: one factorize+solve_dense step of configs[1] for the batch sizes a GPU gets when the
4,096-basis batch is sharded over 1/2/4/8 GPUs, for several CTA sizes (latency per basis vs bases in flight).
usage: python scripts/sweep_batchsize.py [nmat,nmat,...] [threads,...] [dense_k]"""
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blu_b200 import BLUBatch, gen  # noqa: E402

nmats = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4096, 2048, 1024, 512]
nts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [128, 256, 512]
kd = int(sys.argv[3]) if len(sys.argv) > 3 else -1
M = 2000
bb, be, bi, bx, rhs = gen.batch(max(nmats), M, 700, 5.0, 2000, 3000)
for nmat in nmats:
    nb = int(be[nmat * M - 1])
    cap = int((be[:nmat * M] - bb[:nmat * M]).reshape(nmat, M).sum(1).max())
    for nt in nts:
        b = BLUBatch(nmat, M, cap, device=0)
        b.threads_per_basis = nt
        if kd >= 0:
            b.dense_k = kd
        assert b.upload(bb[:nmat * M], be[:nmat * M], bi[:nb], bx[:nb], rhs[:nmat * M]) == 0
        ms = []
        for it in range(3):
            assert b.factorize_resident() == 0
            assert b.solve_dense_resident("N") == 0
            ms.append((b.last_kernel_ms(0), b.last_kernel_ms(3), b.last_kernel_ms(4), b.last_kernel_ms(5), b.last_kernel_ms(2), b.last_kernel_ms(1)))
        f = min(ms)
        gc = np.mean([b.info(k, "ngarbage") for k in range(0, nmat, max(1, nmat // 64))])
        print(f"nmat {nmat} nt {nt}: factorize {f[0]:.1f} ms (head {f[1]:.1f} tail {f[2]:.1f} build {f[3]:.1f} norms {f[4]:.1f}) solve {f[5]:.1f} ms -> {nmat / (f[0] + f[5]) * 1e3:.0f} bases/s; "
              f"realloc rounds {int(b.info(0, 'nrealloc'))}, garbage collections per basis {gc:.2f}, first pass {ms[0][0]:.1f} ms", flush=True)
        b.close()
