"""Ad-hoc timing sweep on the GPU box (not a test)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from blu_b200 import BLUBatch, gen
from oracle_lib import Oracle

nmat = int(sys.argv[1]) if len(sys.argv) > 1 else 512
m = 2000
t = time.time()
bb, be, bi, bx, rhs = gen.batch(nmat, m, 700, 5.0, 2000, 3000)
print("gen", time.time() - t, "nnz/mat", len(bi) / nmat)
cap = int((be - bb).reshape(nmat, m).sum(1).max())
# CPU oracle on a few
t = time.time()
for k in range(4):
    cp, ri, v = gen.basis(2000 + k, m, 700, 5.0)
    o = Oracle(m, 60 * len(v)); o.set_param("check_file_diff", 0)
    o.factorize(cp[:-1], cp[1:], ri, v); o.solve_dense(rhs[k * m:(k + 1) * m])
print("oracle ms/matrix (no file_diff)", (time.time() - t) / 4 * 1e3)
NTS = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [64, 128, 256, 512]
for nt in NTS:
    b = BLUBatch(nmat, m, cap)
    b.threads_per_basis = nt
    b.l_mem = 100000; b.u_mem = 100000; b.w_mem = 160000
    b.upload(bb, be, bi, bx, rhs)
    for rep in range(2):
        st = b.factorize_resident()
        tf = b.last_kernel_ms(0)
        b.solve_dense_resident("N")
        ts = b.last_kernel_ms(1)
    _, x, status = b.download()
    print(f"nt={nt} nmat={nmat} status={st} bad={(status != 0).sum()} factorize {tf:.2f} ms solve {ts:.2f} ms -> {nmat / (tf + ts) * 1e3:.0f} matrices/s; nrealloc {b.info(0, 'nrealloc')}", flush=True)
    names = ["validate+T", "singletons", "setup_bump", "search", "piv_srow", "piv_scol", "piv_dbl", "piv_small", "piv_any", "build", "remove", "total"]
    k = 0
    tot = b.info(k, "t_phase11")
    print("   phases(% of total cycles, matrix 0):", " ".join(f"{n}={100 * b.info(k, f't_phase{q}') / tot:.1f}" for q, n in enumerate(names)), f"total_cycles={tot:.3g}")
    print("   pivots by kind:", [int(b.info(k, f"n_kind{q}")) for q in range(5)], "rank0", int(b.info(k, "m") - b.info(k, "bump_size")))
    b.close()
