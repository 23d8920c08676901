"""Generates the committed fixtures of tests/golden/.

The reference (rwl/blu) is a Rust crate with no tests and no expected outputs in-tree, and it
cannot be compiled in this image (no Rust toolchain) -- so there are no reference-generated
vectors to commit.  What is pinned here instead:

  kat_simple.json      the one known-answer fixture of the reference, examples/simple.rs:21-33
                       (10x10, 32 nnz, rhs b); expected x solved independently with scipy
                       (SuperLU), plus the first two pivots hand-traced from the reference code
                       in SURVEY.md section 4.
  oracle_snapshot.npz  outputs of the C restatement (oracle/) on seeded inputs, so that any later
                       change of the oracle (or of the generator) is caught: permutations, rank,
                       nnz(L), nnz(U), factor_flops and a checksum of the factor values.

Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [  # (seed, m, nslack, pmean, cap)
    (11, 300, 90, 4.0, 20), (21, 120, 0, 3.0, 20), (31, 500, 150, 4.0, 20), (13, 160, 0, 10.0, 60), (1001, 1000, 300, 4.0, 20),
]


def checksum(a):
    return int(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def main():
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    from blu_b200 import gen
    from oracle_lib import Oracle
    arow = [0, 7, 8, 1, 4, 9, 2, 9, 3, 6, 7, 8, 9, 1, 4, 5, 3, 6, 9, 0, 3, 7, 8, 0, 3, 7, 8, 1, 2, 3, 6, 9]
    acolst = [0, 3, 6, 8, 13, 15, 16, 19, 23, 27, 32]
    a = [2.1, 0.14, 0.09, 1.1, 0.06, 0.03, 1.7, 0.04, 1.0, 0.32, 0.19, 0.32, 0.44, 0.06, 1.6, 2.2, 0.32, 1.9, 0.43,
         0.14, 0.19, 1.1, 0.22, 0.09, 0.32, 0.22, 2.4, 0.03, 0.04, 0.44, 0.43, 3.2]
    b = [0.403, 0.28, 0.55, 1.504, 0.812, 1.32, 1.888, 1.168, 2.473, 3.695]
    A = sp.csc_matrix((a, arow, acolst), shape=(10, 10))
    x = spl.splu(A).solve(np.array(b))
    xt = spl.splu(A.T.tocsc()).solve(np.array(b))
    json.dump({"source": "examples/simple.rs:21-33", "colptr": acolst, "rowidx": arow, "values": a, "rhs": b,
               "x_scipy": x.tolist(), "xT_scipy": xt.tolist(), "x_closed_form": [0.1 * k for k in range(1, 11)],
               "first_pivots": [[5, 5], [2, 2]]},
              open(os.path.join(HERE, "kat_simple.json"), "w"), indent=1)
    out = {}
    for (seed, m, nslack, pmean, cap) in CASES:
        cp, ri, v = gen.basis(seed, m, nslack, pmean, cap)
        o = Oracle(m, 400 * len(v) + 100)
        st = o.factorize(cp[:-1], cp[1:], ri, v)
        _, f = o.get_factors()
        tag = f"s{seed}_m{m}"
        out[tag + "_input_crc"] = np.array([checksum(cp), checksum(ri), checksum(v)], dtype=np.int64)
        out[tag + "_rowperm"] = f["rowperm"]
        out[tag + "_colperm"] = f["colperm"]
        out[tag + "_stats"] = np.array([st, o.info("rank"), o.info("l_nz"), o.info("u_nz"), o.info("factor_flops"),
                                        o.info("nsearch_pivot"), o.info("bump_size"), o.info("bump_nz")], dtype=np.int64)
        out[tag + "_value_crc"] = np.array([checksum(f["l_rowidx"]), checksum(f["l_value"]), checksum(f["u_rowidx"]), checksum(f["u_value"])], dtype=np.int64)
        _, xs = o.solve_dense(gen.rhs(seed + 1, m), "N")
        out[tag + "_x"] = xs
    np.savez_compressed(os.path.join(HERE, "oracle_snapshot.npz"), **out)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
