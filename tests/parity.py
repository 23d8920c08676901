"""Shared parity helpers: compare the CUDA path with the CPU oracle on the same input."""
import numpy as np
from oracle_lib import Oracle

STATS = ["rank", "bump_size", "bump_nz", "matrix_nz", "l_nz", "u_nz", "factor_flops", "nsearch_pivot",
         "min_pivot", "max_pivot", "nelim_div"]


def oracle_for(m, nnz, factor=60):
    # generous stores: the oracle then never takes the Reallocate path, whose retried
    # pivot steps double-count factor_flops (pivot.rs:108)
    return Oracle(m, factor * nnz + 100)


def assert_factor_parity(g, o, check_stats=True):
    """Pivot sequence, permutations, rank, L/U patterns and values: bit-exact."""
    sto, fo = o.get_factors()
    stg, fg = g.get_factors()
    assert sto == stg == 0
    for k in fo:
        assert fo[k].shape == fg[k].shape, k
        assert np.array_equal(fo[k], fg[k]), f"{k} differs from the oracle"
    if check_stats:
        for n in STATS:
            assert o.info(n) == g.info(n), n
    assert g.info("internal_error") == 0
    return fg


def backward_error(cp, ri, v, f, m, rank):
    """Backward error of B[rowperm,colperm] = L U, dependent columns replaced by unit columns
    (get_factors.rs:17-20).  Returns (||R||_F / || |L||U| ||_F, ||R||_F / ||B||_F).

    The first (growth-scaled) figure is the one bounded by 1e-14: threshold pivoting with
    reltol = 0.1 (lu.rs:252) admits element growth, and on the synthetic bases the REFERENCE
    algorithm's own factors give ||R||/||B|| of 1e-13 .. 1e-12 (measured with the oracle);
    the CUDA factors are bit-identical to those, so the second figure is only bounded loosely."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    B = sp.csc_matrix((v, ri, cp), shape=(m, m)).tolil()
    rp, cpm = f["rowperm"], f["colperm"]
    for k in range(rank, m):
        B[:, cpm[k]] = 0
        B[rp[k], cpm[k]] = 1.0
    B = B.tocsc()
    L = sp.csc_matrix((f["l_value"], f["l_rowidx"], f["l_colptr"]), shape=(m, m))
    U = sp.csc_matrix((f["u_value"], f["u_rowidx"], f["u_colptr"]), shape=(m, m))
    R = B[rp, :][:, cpm] - L @ U
    nr = spl.norm(R)
    return nr / max(spl.norm(abs(L) @ abs(U)), 1e-300), nr / max(spl.norm(B), 1e-300)


def assert_backward_error(cp, ri, v, f, m, rank):
    scaled, plain = backward_error(cp, ri, v, f, m, rank)
    assert scaled <= 1e-14, scaled
    assert plain <= 1e-10, plain
