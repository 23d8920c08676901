"""Shared parity helpers: compare the CUDA path with the CPU oracle on the same input."""
import numpy as np
from oracle_lib import Oracle

STATS = ["rank", "bump_size", "bump_nz", "matrix_nz", "l_nz", "u_nz", "factor_flops", "nsearch_pivot",
         "min_pivot", "max_pivot", "nelim_div"]


def oracle_for(m, nnz):
    # generous stores: the oracle then never takes the Reallocate path, whose retried
    # pivot steps double-count factor_flops (pivot.rs:108)
    return Oracle(m, 60 * nnz + 100)


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
    """||B[rowperm,colperm] - L U|| / ||B|| with dependent columns replaced by unit columns
    (get_factors.rs:17-20)."""
    import scipy.sparse as sp
    B = sp.csc_matrix((v, ri, cp), shape=(m, m)).tolil()
    rp, cpm = f["rowperm"], f["colperm"]
    for k in range(rank, m):
        B[:, cpm[k]] = 0
        B[rp[k], cpm[k]] = 1.0
    B = B.tocsc()
    L = sp.csc_matrix((f["l_value"], f["l_rowidx"], f["l_colptr"]), shape=(m, m))
    U = sp.csc_matrix((f["u_value"], f["u_rowidx"], f["u_colptr"]), shape=(m, m))
    P = B[rp, :][:, cpm]
    R = (P - L @ U)
    nb = abs(B).sum()
    return abs(R).sum() / (nb if nb else 1.0)
