"""Ad-hoc: time of the condest/residual kernel alone (not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blu_b200 import BLUBatch, gen
nmat, m = 4096, 2000
bb, be, bi, bx, rhs = gen.batch(nmat, m, 700, 5.0, 2000, 3000)
b = BLUBatch(nmat, m, int((be - bb).reshape(nmat, m).sum(1).max()))
import os as _os
b.l_mem = int(_os.environ.get('L_MEM', 100000)); b.u_mem = int(_os.environ.get('U_MEM', 100000)); b.w_mem = int(_os.environ.get('W_MEM', 500000))
b.upload(bb, be, bi, bx, rhs)
for rep in range(2):
    b.factorize_resident()
    b.solve_dense_resident("N")
print(f"w_mem {b.get_param('w_mem'):.0f} l_mem {b.get_param('l_mem'):.0f} nrealloc {b.info(0, 'nrealloc'):.0f} ngarbage(0) {b.info(0, 'ngarbage'):.0f} | factorize total {b.last_kernel_ms(0):.1f} ms, norms kernel {b.last_kernel_ms(2):.2f} ms, solve_dense {b.last_kernel_ms(1):.2f} ms", flush=True)
import numpy as np
c = np.array([[b.info(k, f"norms_cyc{q}") for q in range(8)] for k in range(0, nmat, 97)])
print("norms warp cycles (mean over sampled bases): condestL %.3g condestU %.3g resF %.3g resT %.3g | resF pieces: Ldot %.3g Uaxpy %.3g Bpass %.3g onenorms %.3g" % tuple(c.mean(0)), flush=True)
