"""CPU tests (`-m "not gpu"`): pin the oracle (the C restatement of rwl/blu in oracle/).

The reference has no tests, so the oracle is validated against
(1) the one known-answer fixture, examples/simple.rs, (2) algebraic invariants of get_factors,
(3) scipy's SuperLU for solution values, (4) the committed snapshot of its own outputs,
(5) invariants of the (repaired) Forrest-Tomlin update, (6) the status-code table."""
import json
import os
import zlib

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from blu_b200 import gen
from oracle_lib import Oracle, batch_factorize_solve
from parity import backward_error

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat_simple.json")))


def crc(a):
    return int(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def csc(cp, ri, v, m):
    return sp.csc_matrix((v, ri, cp), shape=(m, m))


def test_kat_simple_rs():
    """examples/simple.rs:21-44 -> x = 0.1 .. 1.0; pivots (5,5) then (2,2)."""
    cp, ri, v, b = (np.array(KAT[k]) for k in ("colptr", "rowidx", "values", "rhs"))
    o = Oracle(10, 32)
    assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
    st, x = o.solve_dense(b, "N")
    assert st == 0
    assert np.abs(x - np.array(KAT["x_closed_form"])).max() < 1e-14
    assert np.abs(x - np.array(KAT["x_scipy"])).max() < 1e-13
    st, xt = o.solve_dense(b, "T")
    assert np.abs(xt - np.array(KAT["xT_scipy"])).max() < 1e-13
    _, f = o.get_factors()
    assert [[int(f["rowperm"][k]), int(f["colperm"][k])] for k in range(2)] == KAT["first_pivots"]
    # hand-traced second pivot: L gets 0.04/1.7 in row 9, U gets 0.04 in column 9
    tr_o = Oracle(10, 32); tr_o.trace(True)
    tr_o.factorize(cp[:-1], cp[1:], ri, v)
    t = tr_o.get_trace()
    assert (t[0][0], t[0][1], t[0][2]) == (5, 5, 2.2)
    assert (t[1][0], t[1][1], t[1][2]) == (2, 2, 1.7) and t[1][3] == 4   # doubleton column variant


def test_snapshot():
    """The oracle still produces what was committed in tests/golden/oracle_snapshot.npz."""
    from golden.make_golden import CASES
    G = np.load(os.path.join(HERE, "golden", "oracle_snapshot.npz"))
    for (seed, m, nslack, pmean, cap) in CASES:
        cp, ri, v = gen.basis(seed, m, nslack, pmean, cap)
        tag = f"s{seed}_m{m}"
        assert [crc(cp), crc(ri), crc(v)] == list(G[tag + "_input_crc"]), "generator changed"
        o = Oracle(m, 400 * len(v) + 100)
        st = o.factorize(cp[:-1], cp[1:], ri, v)
        _, f = o.get_factors()
        assert np.array_equal(f["rowperm"], G[tag + "_rowperm"]) and np.array_equal(f["colperm"], G[tag + "_colperm"])
        stats = [st, o.info("rank"), o.info("l_nz"), o.info("u_nz"), o.info("factor_flops"), o.info("nsearch_pivot"),
                 o.info("bump_size"), o.info("bump_nz")]
        assert [int(s) for s in stats] == list(G[tag + "_stats"])
        assert [crc(f["l_rowidx"]), crc(f["l_value"]), crc(f["u_rowidx"]), crc(f["u_value"])] == list(G[tag + "_value_crc"])
        _, x = o.solve_dense(gen.rhs(seed + 1, m), "N")
        assert np.array_equal(x, G[tag + "_x"])


@pytest.mark.parametrize("seed", range(6))
def test_factors_reproduce_b(seed):
    """B[rowperm, colperm] = L U (get_factors.rs:17-20), growth-scaled backward error <= 1e-14."""
    m = 200 + 37 * seed
    cp, ri, v = gen.basis(500 + seed, m, int(0.3 * m) if seed % 2 else 0, 3.0 + seed)
    o = Oracle(m, 400 * len(v))
    st = o.factorize(cp[:-1], cp[1:], ri, v)
    assert st in (0, 2)
    _, f = o.get_factors()
    rank = int(o.info("rank"))
    assert sorted(f["rowperm"]) == list(range(m)) and sorted(f["colperm"]) == list(range(m))
    scaled, plain = backward_error(cp, ri, v, f, m, rank)
    assert scaled <= 1e-14 and plain <= 1e-10
    # L unit lower triangular with sorted rows, U upper triangular with the diagonal last
    L = sp.csc_matrix((f["l_value"], f["l_rowidx"], f["l_colptr"]), shape=(m, m))
    U = sp.csc_matrix((f["u_value"], f["u_rowidx"], f["u_colptr"]), shape=(m, m))
    assert sp.triu(L, 1).nnz == 0 and np.all(L.diagonal() == 1.0)
    assert sp.tril(U, -1).nnz == 0
    for k in range(m):
        assert f["l_rowidx"][f["l_colptr"][k]] == k
        assert f["u_rowidx"][f["u_colptr"][k + 1] - 1] == k


@pytest.mark.parametrize("seed", range(4))
def test_solves_match_superlu(seed):
    m = 400
    cp, ri, v = gen.basis(700 + seed, m, 120, 4.0)
    A = csc(cp, ri, v, m)
    o = Oracle(m, 400 * len(v))
    assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
    assert o.info("residual_test") < 1e-12          # lu.rs:610
    b = gen.rhs(800 + seed, m)
    lu = spl.splu(A)
    for tr, ref in (("N", lu.solve(b)), ("T", lu.solve(b, trans="T"))):
        st, x = o.solve_dense(b, tr)
        assert st == 0
        assert np.abs(x - ref).max() / np.abs(ref).max() < 1e-10
        # sparse solve of a sparse rhs equals the dense solve of the same rhs
        idx, val = gen.sparse_rhs(900 + seed, m, 7)
        rs = np.zeros(m); rs[idx] = val
        _, xd = o.solve_dense(rs, tr)
        assert o.solve_sparse(len(idx), idx, val, tr) == 0
        xs = o.lhs
        assert np.abs(xs - xd).max() <= 1e-12 * max(np.abs(xd).max(), 1.0)
        pat = set(o.ilhs[:o.nzlhs].tolist())
        assert pat == set(np.nonzero(xs)[0].tolist())


def test_singular_and_status_codes():
    m = 10
    cp, ri, v, b = (np.array(KAT[k]) for k in ("colptr", "rowidx", "values", "rhs"))
    o = Oracle(m, 32)
    assert o.solve_dense(b)[0] == -2                          # solve_dense.rs:25
    assert o.get_factors()[0] == -2                           # D9
    assert o.solve_sparse(1, np.array([0]), np.array([1.0])) == -2
    assert o.update(1.0) == -2                                # update.rs:50
    bad_end = cp[1:].copy(); bad_end[3] = cp[3] - 1
    assert o.factorize(cp[:-1], bad_end, ri, v) == -4         # singletons.rs:122-131
    ri2 = ri.copy(); ri2[5] = 10
    assert o.factorize(cp[:-1], cp[1:], ri2, v) == -4         # singletons.rs:157-173
    ri3 = ri.copy(); ri3[1] = ri3[0]
    assert o.factorize(cp[:-1], cp[1:], ri3, v) == -4         # singletons.rs:194-200
    v2 = v.copy(); v2[cp[4]:cp[5]] = 0.0                      # a zero column => rank m-1
    assert o.factorize(cp[:-1], cp[1:], ri, v2) == 2          # factorize.rs:176-178
    assert o.info("rank") == m - 1
    assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
    assert o.solve_sparse(11, np.arange(11) % 10, np.ones(11)) == -4   # solve_sparse.rs:45-58
    assert o.solve_sparse(1, np.array([10]), np.array([1.0])) == -4
    assert o.solve_for_update(1, np.array([0]), None, "N") == -3       # solve_for_update.rs:82-83
    assert o.update(1.0) == -2                                # no solve_for_update pair yet


def test_forrest_tomlin_update_invariants():
    """H7: after every (repaired) update the object must solve like a fresh factorization of the
    updated basis, and pivot_error must stay small (update.rs:513,942)."""
    m = 300
    cp, ri, v = gen.basis(41, m, 90, 4.0)
    pool_cp, pool_ri, pool_v = gen.basis(42, m, 0, 4.0)
    A = csc(cp, ri, v, m).tolil()
    o = Oracle(m, 400 * len(v))
    assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
    rng = np.random.default_rng(5)
    kinds = set()
    for it in range(40):
        q = it
        idx = pool_ri[pool_cp[q]:pool_cp[q + 1]]
        val = pool_v[pool_cp[q]:pool_cp[q + 1]]
        assert o.solve_for_update(len(idx), idx, val, "N", want_solution=1) == 0
        lhs = o.lhs
        j = int(np.argmax(np.abs(lhs)))                       # maxvolume.rs:120-131 rule
        xtbl = lhs[j]
        assert o.solve_for_update(1, np.array([j]), None, "T") == 0
        nf0 = o.info("nforrest")
        st = o.update(xtbl)
        assert st == 0, (it, st)
        kinds.add("ft" if o.info("nforrest") > nf0 else "perm")
        assert o.info("pivot_error") < 1e-8
        A[:, j] = 0
        for i_, x_ in zip(idx, val):
            A[i_, j] = x_
        b = rng.uniform(-1, 1, m)
        lu = spl.splu(A.tocsc())
        for tr, ref in (("N", lu.solve(b)), ("T", lu.solve(b, trans="T"))):
            _, x = o.solve_dense(b, tr)
            assert np.abs(x - ref).max() / np.abs(ref).max() < 1e-9, (it, tr)
    assert "ft" in kinds
    assert o.get_factors()[0] == -2                           # get_factors.rs:59-61: invalid after an update


def test_batch_driver_matches_single_instances():
    nmat, m = 6, 200
    bb, be, bi, bx, rhs = gen.batch(nmat, m, 60, 4.0, 100, 200)
    nt, x, st = batch_factorize_solve(nmat, m, bb, be, bi, bx, rhs, nthreads=3)
    assert nt == 3 and (st == 0).all()
    for k in range(nmat):
        cp, ri, v = gen.basis(100 + k, m, 60, 4.0)
        o = Oracle(m, 400 * len(v))
        o.factorize(cp[:-1], cp[1:], ri, v)
        _, xo = o.solve_dense(rhs[k * m:(k + 1) * m], "N")
        assert np.array_equal(x[k], xo)


def _ref_case(seed):
    """One random or structured matrix + one random setting of the tunables that steer pivot choice."""
    import scipy.sparse as sp
    from parity import structured_matrix
    rng = np.random.default_rng(seed)
    m = int(rng.integers(6, 70))
    kind = seed % 9
    if kind < 6:
        cp, ri, v = structured_matrix(kind, m, rng)
    else:
        dens = [0.08, 0.2, 0.5][kind - 6]
        A = sp.random(m, m, density=dens, format="csc", random_state=rng, data_rvs=lambda n: rng.uniform(-1, 1, n)) + sp.diags(rng.uniform(0.5, 2, m))
        A = sp.csc_matrix(A); A.sort_indices()
        cp, ri, v = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64)
    prm = dict(abstol=[1e-14, 1e-8, 1e-3][int(rng.integers(3))], reltol=[0.1, 0.01, 0.5, 1.0][int(rng.integers(4))],
               droptol=[1e-20, 1e-12, 1e-2][int(rng.integers(3))], maxsearch=int(rng.integers(1, 6)),
               search_rows=int(rng.integers(2)), nzbias=[1, -1][int(rng.integers(2))])
    return m, cp, ri, v, prm


def test_second_restatement_agrees_on_pivot_sequences():
    """oracle/blo_ref_pivots.py is a second reading of singletons.rs / setup_bump.rs / markowitz.rs / pivot.rs
    with different data structures.  On >= 1000 random and structured matrices (exact cancellation, rank
    deficiency, dense blocks, bad scaling; search_rows 0 and 1; maxsearch 1..5; several tolerances) it must
    choose the same pivots in the same order as the C oracle -- a shared misreading of tie-breaking or
    bucket order would show here."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from blo_ref_pivots import RefLU
    ncase, checked, outside = 1100, 0, 0
    for seed in range(20000, 20000 + ncase):
        m, cp, ri, v, prm = _ref_case(seed)
        try:
            ref = RefLU(m, cp, ri, v, **prm).factorize()
        except AssertionError:
            outside += 1          # no eligible pivot: factorize_bump.rs:22 asserts, the C oracle aborts there too
            continue
        o = Oracle(m, 200 * len(v) + 2000)
        for k, x in prm.items():
            o.set_param(k, x)
        st = o.factorize(cp[:-1], cp[1:], ri, v)
        assert st in (0, 2), (seed, st)
        _, f = o.get_factors()
        rows, cols = ref.permutations()
        assert int(o.info("rank")) == len(ref.pivots), (seed, prm)
        assert list(f["rowperm"]) == rows, (seed, prm)
        assert list(f["colperm"]) == cols, (seed, prm)
        assert int(o.info("nsearch_pivot")) == ref.nsearch, (seed, prm)
        checked += 1
    assert checked >= 1000, (checked, outside)
