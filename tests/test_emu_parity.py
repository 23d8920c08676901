"""CPU tests: the CUDA kernels compiled with g++ under the SIMT emulator of tests/emu
(barriers, warp votes/shuffles and atomics emulated with fibers) against the oracle.
This checks kernel LOGIC on the GPU-less build box; the `-m gpu` tests are the parity tests
proper.  The emulated library is test infrastructure and never part of the product."""
import os
import subprocess

import numpy as np
import pytest

from blu_b200 import BLU, BLUBatch, gen, load_library
from parity import oracle_for, assert_factor_parity

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(HERE, "emu")])
    return load_library(os.path.join(HERE, "emu", "libblu_emu.so"))


def pair(emu, cp, ri, v, m, nt, ofactor=60):
    o = oracle_for(m, len(v), ofactor)
    so = o.factorize(cp[:-1], cp[1:], ri, v)
    g = BLU(m, len(v), lib=emu)
    g.threads_per_basis = nt
    sg = g.factorize(cp[:-1], cp[1:], ri, v)
    assert so == sg
    return g, o, sg


@pytest.mark.parametrize("nt,seed", [(32, 1), (64, 2), (128, 3)])
def test_emu_factorize_solve(emu, nt, seed):
    m = 160
    cp, ri, v = gen.basis(60 + seed, m, 40, 4.0)
    g, o, st = pair(emu, cp, ri, v, m, nt)
    assert st == 0
    assert_factor_parity(g, o)
    b = gen.rhs(70 + seed, m)
    for tr in "NT":
        _, xo = o.solve_dense(b, tr)
        sg, xg = g.solve_dense(b, tr)
        assert sg == 0
        assert np.array_equal(xg, xo)      # ordered sums: bit-identical, not merely 1e-12


def test_emu_dense_columns(emu):
    m = 200
    cp, ri, v = gen.basis(13, m, 0, 14.0, cap=60)
    g, o, st = pair(emu, cp, ri, v, m, 64, ofactor=400)
    assert_factor_parity(g, o)
    assert g.info("n_kind4") > 0      # pivot_any exercised


def test_emu_singular_and_cancellation(emu):
    import scipy.sparse as sp
    m = 90
    rng = np.random.default_rng(3)
    A = sp.random(m, m, density=0.04, format="csc", random_state=3, data_rvs=lambda n: np.sign(rng.uniform(-1, 1, n)))
    A.sort_indices()
    cp, ri, v = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.copy()
    g, o, st = pair(emu, cp, ri, v, m, 64)
    assert st == 2
    assert_factor_parity(g, o)


def test_emu_batch(emu):
    nmat, m = 3, 100
    bb, be, bi, bx, rhs = gen.batch(nmat, m, 30, 4.0, 9000, 9500)
    b = BLUBatch(nmat, m, int((be - bb).reshape(nmat, m).sum(1).max()), lib=emu)
    st, status = b.factorize(bb, be, bi, bx)
    assert st == 0 and (status == 0).all()
    st, x, sst = b.solve_dense(rhs, "N")
    assert st == 0 and (sst == 0).all()
    for k in range(nmat):
        cp, ri, v = gen.basis(9000 + k, m, 30, 4.0)
        o = oracle_for(m, len(v))
        o.factorize(cp[:-1], cp[1:], ri, v)
        _, fo = o.get_factors()
        _, fg = b.get_factors(k)
        for key in fo:
            assert np.array_equal(fo[key], fg[key]), key
        _, xo = o.solve_dense(rhs[k * m:(k + 1) * m], "N")
        assert np.array_equal(x[k], xo)


def test_emu_solve_sparse(emu):
    from parity import assert_sparse_solve_parity
    m = 160
    cp, ri, v = gen.basis(81, m, 40, 4.0)
    g, o, st = pair(emu, cp, ri, v, m, 64)
    assert st == 0
    assert_sparse_solve_parity(g, o, m, 500)


@pytest.mark.parametrize("m,seed,nslack,niter,tight", [(120, 91, 36, 30, False), (80, 5, 60, 60, True), (150, 7, 75, 50, True)])
def test_emu_update_replay(emu, m, seed, nslack, niter, tight):
    """solve_for_update + update in lockstep with the oracle: symmetric/unsymmetric permutation
    updates, Forrest-Tomlin updates, eta/spike/row-file growth (tight = stores start at nnz(B), so the
    Reallocate protocol of blu.rs:268-291,319-334 runs), U and W compression."""
    from parity import replay_updates, assert_sparse_solve_parity
    cp, ri, v = gen.basis(seed, m, nslack, 4.0)
    pool = gen.basis(seed + 1, m, 0, 3.0)
    o = oracle_for(m, len(v), 400)
    g = BLU(m, len(v), lib=emu)
    g.threads_per_basis = 64
    if tight:
        g.l_mem = len(v); g.u_mem = len(v); g.w_mem = len(v)
    assert o.factorize(cp[:-1], cp[1:], ri, v) == g.factorize(cp[:-1], cp[1:], ri, v) == 0
    nr0 = g.info("nrealloc")
    kinds = replay_updates(g, o, m, pool, niter)
    assert "ft" in kinds and "perm" in kinds
    if tight:
        assert g.info("nrealloc") > nr0
    assert_sparse_solve_parity(g, o, m, 700)
    assert g.get_factors()[0] == o.get_factors()[0] == -2     # get_factors.rs:59-61
    # refactorize the same object: counters reset as LU::reset does (lu.rs:329-396)
    assert o.factorize(cp[:-1], cp[1:], ri, v) == g.factorize(cp[:-1], cp[1:], ri, v) == 0
    assert_factor_parity(g, o, check_stats=False)
    assert_sparse_solve_parity(g, o, m, 800, sizes=(2, 30))


def test_emu_sparse_status_codes(emu):
    """The Err sites of solve_sparse.rs:45-58, solve_for_update.rs:82-106 and update.rs:50, in the
    reference's order of checks."""
    m = 40
    cp, ri, v = gen.basis(3, m, 10, 3.0)
    g = BLU(m, len(v), lib=emu)
    one, x1 = np.array([0]), np.array([1.0])
    assert g.solve_sparse(1, one, x1) == -2                       # never factorized
    assert g.solve_for_update(1, one, None, "N") == -3            # ArgumentMissing comes first
    assert g.solve_for_update(1, one, x1, "N") == -2
    assert g.update(1.0) == -2
    assert g.factorize(cp[:-1], cp[1:], ri, v) == 0
    assert g.solve_sparse(m + 1, np.arange(m + 1) % m, np.ones(m + 1)) == -4
    assert g.solve_sparse(-1, one, x1) == -4
    assert g.solve_sparse(1, np.array([m]), x1) == -4
    assert g.solve_for_update(1, np.array([m]), None, "T") == -4
    assert g.update(1.0) == -2                                    # no solve_for_update pair yet
    assert g.solve_for_update(1, one, x1, "N") == 0
    assert g.update(1.0) == -2                                    # only the forward half was prepared
    assert g.solve_for_update(1, one, None, "T") == 0
    assert g.update(1.0) in (0, -6)
    assert g.solve_sparse(0, np.zeros(0, np.int64), np.zeros(0)) == 0 and g.nzlhs == 0


def test_emu_maxvolume(emu):
    """the maxvolume driver (host loop above the ABI) against the oracle's."""
    from parity import assert_maxvolume_parity
    m, ncol = 60, 100
    g = BLU(m, 600, lib=emu)
    g.threads_per_basis = 64
    o = oracle_for(m, 600, 400)
    nup = assert_maxvolume_parity(g, o, m, ncol, 300)
    assert nup > 0
    assert maxvolume_invalid(g)


def maxvolume_invalid(g):
    from blu_b200 import maxvolume
    st, n = maxvolume(g, 0, np.zeros(1, np.int64), np.zeros(0, np.int64), np.zeros(0), np.zeros(0, np.int64), np.zeros(0, np.int64), 0.5)
    return st == -4 and n == 0                                   # maxvolume.rs:83-90


@pytest.mark.parametrize("case", ["m1", "identity", "empty", "empty_cols", "dense40", "one_dense_row_col", "nonsquare_store"])
def test_emu_edge_shapes(emu, case, monkeypatch):
    """tests/test_gpu_parity.py::test_edge_shapes under the SIMT emulator."""
    import test_gpu_parity as tg
    monkeypatch.setattr(tg, "BLU", lambda m, nnz: BLU(m, nnz, lib=emu))
    tg.test_edge_shapes(case)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_emu_random_lane_order(emu, seed, monkeypatch):
    """The emulator visits the runnable lanes in a pseudo-random order between barriers (EMU_SEED): code
    that silently relies on lane order -- a missing __syncwarp between a read and a write of one word --
    fails parity under some seed.  One compact scenario through every kernel."""
    from parity import replay_updates, assert_sparse_solve_parity
    monkeypatch.setenv("EMU_SEED", str(seed))
    m = 90
    cp, ri, v = gen.basis(200 + seed, m, 30, 4.0)
    pool = gen.basis(300 + seed, m, 0, 3.0)
    o = oracle_for(m, len(v), 400)
    g = BLU(m, len(v), lib=emu)
    g.threads_per_basis = 64
    assert o.factorize(cp[:-1], cp[1:], ri, v) == g.factorize(cp[:-1], cp[1:], ri, v) == 0
    assert_factor_parity(g, o)
    assert_sparse_solve_parity(g, o, m, 900 + seed, sizes=(1, 4, 30))
    replay_updates(g, o, m, pool, 12)
    assert_sparse_solve_parity(g, o, m, 950 + seed, sizes=(2, 40))


def test_emu_solve_dense_multi(emu):
    """many right-hand sides against one factorization == that many solve_dense calls."""
    m = 120
    cp, ri, v = gen.basis(77, m, 30, 4.0)
    g, o, st = pair(emu, cp, ri, v, m, 64)
    assert st == 0
    R = np.stack([gen.rhs(800 + k, m) for k in range(5)])
    for tr in "NT":
        sg, X = g.solve_dense_multi(R, tr)
        assert sg == 0
        for k in range(5):
            _, xo = o.solve_dense(R[k], tr)
            assert np.array_equal(X[k], xo), (tr, k)
    g2 = BLU(m, len(v), lib=emu)
    assert g2.solve_dense_multi(R, "N")[0] == -2          # never factorized


def test_emu_solve_sparse_multi(emu):
    from parity import assert_sparse_multi_parity, replay_updates
    m = 140
    cp, ri, v = gen.basis(88, m, 40, 4.0)
    g, o, st = pair(emu, cp, ri, v, m, 64)
    assert st == 0
    assert_sparse_multi_parity(g, o, m, 1500)
    # also after updates (etas, permuted pivots): the factors are still only read
    replay_updates(g, o, m, gen.basis(89, m, 0, 3.0), 8)
    assert_sparse_multi_parity(g, o, m, 1600, sizes=(2, 25), reps=2)


@pytest.mark.parametrize("tight", [False, True])
def test_emu_batch_replay(emu, tight):
    """Many LPs advancing together: batch solve_for_update + update, one warp per basis.  tight: the stores
    start at nnz(B), so the batch Reallocate protocol (grow every basis' store, content kept, re-run only
    the bases that asked) is exercised."""
    from parity import batch_replay_parity
    nmat, m = 3, 70
    bb, be, bi, bx, rhs = gen.batch(nmat, m, 25, 4.0, 9100, 9600)
    b = BLUBatch(nmat, m, int((be - bb).reshape(nmat, m).sum(1).max()), lib=emu)
    if tight:
        b.l_mem = 300; b.u_mem = 300; b.w_mem = 300
    else:
        b.l_mem = 40000; b.u_mem = 40000; b.w_mem = 60000
    st, status = b.factorize(bb, be, bi, bx)
    assert st == 0 and (status == 0).all()
    oracles, pools = [], []
    for k in range(nmat):
        cp, ri, v = gen.basis(9100 + k, m, 25, 4.0)
        o = oracle_for(m, len(v), 400)
        assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
        oracles.append(o); pools.append(gen.basis(9700 + k, m, 0, 3.0))
    nr0 = b.info(0, "nrealloc")
    batch_replay_parity(b, oracles, m, pools, 8 if not tight else 20)
    if tight:
        assert b.info(0, "nrealloc") > nr0
    st, x, sst = b.solve_dense(rhs, "N")
    assert st == 0
    for k, o in enumerate(oracles):
        _, xo = o.solve_dense(rhs[k * m:(k + 1) * m], "N")
        assert np.array_equal(x[k], xo)


@pytest.mark.parametrize("seed", range(300, 306))
def test_emu_tunables(emu, seed):
    """Every `pub` tunable of LU (lu.rs:10-66) drawn at random on both sides: the CUDA path follows the
    oracle through each of them (thresholds, search depth, line padding, sparse/dense switch)."""
    from parity import tunables_case
    tunables_case(lambda m, nnz: BLU(m, nnz, lib=emu), 70, seed, nupd=6)


@pytest.mark.parametrize("seed,m", [(9004, 56), (9040, 88), (9000, 8), (9001, 45), (9002, 82), (9003, 19), (9005, 93)])
def test_emu_structures(emu, seed, m):
    """Structures the synthetic generator does not produce, under random tunables.  Seeds 9004 and 9040
    are regressions: badly scaled columns whose maximum falls below abstol while their pivot-row entry is
    dropped from U stay non-empty, and the search passes over them without counting them
    (markowitz.rs:88-90, defect D6 repaired); 9004 also caught a divergent barrier in post_remove_cols."""
    from parity import structured_case
    structured_case(lambda m, nnz: BLU(m, nnz, lib=emu), m, seed)


@pytest.mark.parametrize("m,seed0", [(41, 610), (96, 660), (151, 710)])
def test_emu_batch_structures(emu, m, seed0):
    from parity import batch_structured_case
    batch_structured_case(lambda n, m, cap: BLUBatch(n, m, cap, lib=emu), m, seed0)


@pytest.mark.parametrize("kd,nt", [(0, 64), (32, 64), (64, 128), (160, 32), (256, 64)])
def test_emu_dense_tail_orders(emu, kd, nt):
    """The dense tail (blu_factor_dense.cuh) never changes a result: the same basis for several switch
    orders -- 0 = sparse to the end, <= 160 = values in shared memory, larger = values in HBM -- pivot_any and
    pivot_small both running inside it."""
    m = 230
    cp, ri, v = gen.basis(313, m, 0, 12.0, cap=60)
    o = oracle_for(m, len(v), 400)
    assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
    g = BLU(m, len(v), lib=emu)
    g.threads_per_basis = nt
    g.dense_k = kd
    assert g.factorize(cp[:-1], cp[1:], ri, v) == 0
    assert_factor_parity(g, o)
    assert g.info("n_kind4") > 0 and g.info("n_kind3") > 0
    steps = g.info("n_kind5")
    assert (steps == 0) if kd == 0 else (steps >= min(kd, m) - 8)


@pytest.mark.parametrize("kd,tail", [(64, 128), (160, 256)])
def test_emu_split_batch(emu, kd, tail):
    """A batch run as three launches (sparse head, dense tail with the active submatrix in shared memory,
    build_factors: blu_factor_build.cuh k_factorize modes) equals the oracle basis by basis."""
    from parity import STATS
    nmat, m = 3, 170
    bb, be, bi, bx, rhs = gen.batch(nmat, m, 30, 5.0, 9100, 9600)
    b = BLUBatch(nmat, m, int((be - bb).reshape(nmat, m).sum(1).max()), lib=emu)
    b.dense_k = kd; b.split_min = 0; b.threads_per_basis = 64; b.tail_threads = tail
    l0 = b.launch_count()
    st, status = b.factorize(bb, be, bi, bx)
    assert st == 0 and (status == 0).all()
    assert b.launch_count() - l0 == 4       # head, tail, build + the condest/residual kernel
    st, x, sst = b.solve_dense(rhs, "T")
    assert st == 0 and (sst == 0).all()
    for k in range(nmat):
        cp, ri, v = gen.basis(9100 + k, m, 30, 5.0)
        o = oracle_for(m, len(v), 400)
        assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
        _, fo = o.get_factors()
        _, fg = b.get_factors(k)
        for key in fo:
            assert np.array_equal(fo[key], fg[key]), (k, key)
        for name in STATS:
            assert o.info(name) == b.info(k, name), (k, name)
        assert b.info(k, "n_kind5") > 0
        _, xo = o.solve_dense(rhs[k * m:(k + 1) * m], "T")
        assert np.array_equal(x[k], xo)


@pytest.mark.parametrize("kd,kbig,tail,m", [(64, 128, 128, 200), (32, 64, 64, 120), (96, 160, 256, 220), (160, 192, 256, 300)])
def test_emu_two_stage_tail(emu, kd, kbig, tail, m):
    """Two-stage dense tail (BLU_P_DENSE_K_BIG): the active submatrix turns dense at order kbig with its values in
    HBM/L2, is compacted into shared memory at order kd (dense_restage), and the factors stay those of the oracle."""
    from parity import STATS
    nmat = 3
    bb, be, bi, bx, rhs = gen.batch(nmat, m, 30, 5.0, 9300, 9800)
    b = BLUBatch(nmat, m, int((be - bb).reshape(nmat, m).sum(1).max()), lib=emu)
    b.dense_k = kd; b.dense_k_big = kbig; b.split_min = 0; b.threads_per_basis = 64; b.tail_threads = tail
    assert int(b.get_param("dense_k_big")) == kbig
    st, status = b.factorize(bb, be, bi, bx)
    assert st == 0 and (status == 0).all()
    st, x, sst = b.solve_dense(rhs, "N")
    assert st == 0 and (sst == 0).all()
    for k in range(nmat):
        cp, ri, v = gen.basis(9300 + k, m, 30, 5.0)
        o = oracle_for(m, len(v), 400)
        assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
        _, fo = o.get_factors()
        _, fg = b.get_factors(k)
        for key in fo:
            assert np.array_equal(fo[key], fg[key]), (k, key)
        for name in STATS:
            assert o.info(name) == b.info(k, name), (k, name)
        assert b.info(k, "n_kind6") >= 2, "both stages ran"      # dense_enter + dense_restage
        _, xo = o.solve_dense(rhs[k * m:(k + 1) * m], "N")
        assert np.array_equal(x[k], xo)


def test_emu_two_stage_tail_one_launch(emu):
    """The two-stage tail inside ONE launch (BLU_MODE_WHOLE: a batch below split_min): the CTA that ran the sparse
    head enters stage 1, restages and finishes; same factors as the oracle."""
    from parity import STATS
    nmat, m = 2, 220
    bb, be, bi, bx, rhs = gen.batch(nmat, m, 40, 5.0, 9350, 9850)
    b = BLUBatch(nmat, m, int((be - bb).reshape(nmat, m).sum(1).max()), lib=emu)
    b.dense_k = 64; b.dense_k_big = 128; b.split_min = 1000; b.threads_per_basis = 128
    l0 = b.launch_count()
    st, status = b.factorize(bb, be, bi, bx)
    assert st == 0 and (status == 0).all()
    assert b.launch_count() - l0 == 2       # one k_factorize + the condest/residual kernel
    for k in range(nmat):
        cp, ri, v = gen.basis(9350 + k, m, 40, 5.0)
        o = oracle_for(m, len(v), 400)
        assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
        _, fo = o.get_factors()
        _, fg = b.get_factors(k)
        for key in fo:
            assert np.array_equal(fo[key], fg[key]), (k, key)
        for name in STATS:
            assert o.info(name) == b.info(k, name), (k, name)
        assert b.info(k, "n_kind6") >= 2, "both stages ran"


def test_emu_dense_tail_structures(emu):
    """Exact cancellation, rank deficiency and columns that fall below abstol inside the dense tail
    (the paths that leave it early: dense_exit + pivot.rs:96-106 on the line file)."""
    from parity import structured_case
    os.environ["BLU_B200_DENSE_K"] = "64"
    try:
        for seed, m in [(9000, 120), (9003, 140), (9005, 90), (9006, 150), (9012, 100)]:
            structured_case(lambda mm, nnz: BLU(mm, nnz, lib=emu), m, seed, nupd=2)
    finally:
        del os.environ["BLU_B200_DENSE_K"]


def hungry_batch(nmat, m, heavy):
    """Sparse bases plus a few (`heavy`) much denser ones, in the batch ABI layout."""
    mats = [gen.basis(9400 + k, m, 0 if k in heavy else m // 2, 9.0 if k in heavy else 2.5, cap=40) for k in range(nmat)]
    off, bb, be = 0, [], []
    for cp, ri, v in mats:
        bb.append(cp[:-1] + off); be.append(cp[1:] + off); off += len(v)
    return mats, np.concatenate(bb), np.concatenate(be), np.concatenate([x[1] for x in mats]), np.concatenate([x[2] for x in mats])


def check_hungry(b, mats, heavy, m):
    for k, (cp, ri, v) in enumerate(mats):
        o = oracle_for(m, len(v), 400)
        assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
        _, fo = o.get_factors()
        _, fg = b.get_factors(k)
        for key in fo:
            assert np.array_equal(fo[key], fg[key]), (k, key)
        # blu.rs:95-118 re-runs the basis that asked for memory -- and only that one
        assert (b.info(k, "nruns") > 1) == (k in heavy), (k, b.info(k, "nruns"))
    assert b.info(0, "nrealloc") >= 1


@pytest.mark.parametrize("split", [False, True])
def test_emu_per_basis_reallocate(emu, split):
    """One hungry basis in a batch gets private, larger stores and runs again alone (grow_hungry_slots);
    afterwards a Reallocate in the update path folds the private stores back (grow_batch_stores)."""
    nmat, m, heavy = 5, 120, {3}
    mats, bb, be, bi, bx = hungry_batch(nmat, m, heavy)
    b = BLUBatch(nmat, m, max(len(x[2]) for x in mats), lib=emu)
    b.threads_per_basis = 64
    if split:
        b.split_min = 0; b.dense_k = 64; b.tail_threads = 128
    light = max(len(x[2]) for k, x in enumerate(mats) if k not in heavy)
    b.l_mem = 8 * light; b.u_mem = 8 * light; b.w_mem = 10 * light
    st, status = b.factorize(bb, be, bi, bx)
    assert st == 0 and (status == 0).all(), status
    check_hungry(b, mats, heavy, m)
    rhs = np.random.default_rng(1).uniform(-1, 1, nmat * m)
    st, x, sst = b.solve_dense(rhs, "N")
    assert st == 0 and (sst == 0).all()
    for k, (cp, ri, v) in enumerate(mats):
        o = oracle_for(m, len(v), 400)
        o.factorize(cp[:-1], cp[1:], ri, v)
        _, xo = o.solve_dense(rhs[k * m:(k + 1) * m], "N")
        assert np.array_equal(x[k], xo), k


def test_emu_bad_column_pointers(emu):
    """Column pointers outside b_i / b_x are an invalid argument for that basis (the reference would panic on
    the slice), never an out-of-bounds read."""
    nmat, m = 3, 60
    bb, be, bi, bx, rhs = gen.batch(nmat, m, 20, 3.0, 9500, 9550)
    b = BLUBatch(nmat, m, int((be - bb).reshape(nmat, m).sum(1).max()), lib=emu)
    be2 = be.copy(); be2[m + 5] = len(bi) + 7          # basis 1: a column ends beyond the arrays
    bb2 = bb.copy(); bb2[2 * m + 1] = -3               # basis 2: a column starts before them
    st, status = b.factorize(bb2, be2, bi, bx)
    assert st == 0 and list(status) == [0, -4, -4]
    g = BLU(m, 10, lib=emu)
    assert g.set_param("maxsearch", 33) == -4 and g.set_param("maxsearch", 32) == 0   # MAXCAND


def test_emu_batch_replay_after_private_stores(emu):
    """Tight stores: every basis gets private stores during factorize (grow_hungry_slots); the first
    Reallocate of the update path then folds them back into uniform ones with their content
    (grow_batch_stores / k_store_regrow) and the replay goes on, in lockstep with one oracle per basis."""
    from parity import batch_replay_parity
    nmat, m = 3, 90
    bb, be, bi, bx, rhs = gen.batch(nmat, m, 27, 4.0, 9100, 9600)
    b = BLUBatch(nmat, m, int((be - bb).reshape(nmat, m).sum(1).max()), lib=emu)
    b.l_mem = 450; b.u_mem = 450; b.w_mem = 450
    st, status = b.factorize(bb, be, bi, bx)
    assert st == 0 and (status == 0).all()
    nr0 = b.info(0, "nrealloc")
    assert nr0 >= 1
    oracles, pools = [], []
    for k in range(nmat):
        cp, ri, v = gen.basis(9100 + k, m, 27, 4.0)
        o = oracle_for(m, len(v), 400)
        assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
        oracles.append(o); pools.append(gen.basis(9700 + k, m, 0, 3.0))
    batch_replay_parity(b, oracles, m, pools, 25)
    assert b.info(0, "nrealloc") > nr0
    st, x, sst = b.solve_dense(rhs, "T")
    assert st == 0
    for k, o in enumerate(oracles):
        _, xo = o.solve_dense(rhs[k * m:(k + 1) * m], "T")
        assert np.array_equal(x[k], xo)


@pytest.mark.parametrize("m,seed,nslack,dens,nt,kd,ms", [(300, 5, 60, 3.0, 128, 0, 3), (400, 6, 80, 3.0, 256, 64, 4), (1200, 7, 100, 2.0, 128, 32, 1)])
def test_emu_tree_search(emu, m, seed, nslack, dens, nt, kd, ms):
    """Markowitz candidates through the min-tree over the column keys (blu_factor_bump.cuh ctree_*), forced on
    for small bumps with tree_min: same pivots, same nsearch_pivot as the bucket walk of markowitz.rs:80-123."""
    cp, ri, v = gen.basis(seed, m, nslack, dens)
    o = oracle_for(m, len(v), 100)
    o.set_param("maxsearch", ms)
    assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
    g = BLU(m, len(v), lib=emu)
    g.threads_per_basis = nt; g.dense_k = kd; g.maxsearch = ms; g.tree_min = 16
    assert g.factorize(cp[:-1], cp[1:], ri, v) == 0
    assert_factor_parity(g, o)


def test_emu_tree_search_structures(emu):
    """... with rank deficiency and emptied columns (the tree is repaired after remove_col, pivot.rs:1333-1381)."""
    from parity import structured_case
    os.environ["BLU_B200_TREE_MIN"] = "8"; os.environ["BLU_B200_DENSE_K"] = "32"
    try:
        for seed, m in [(9000, 120), (9005, 90), (9006, 150), (9003, 130)]:
            structured_case(lambda mm, nnz: BLU(mm, nnz, lib=emu), m, seed, nupd=2)
    finally:
        del os.environ["BLU_B200_TREE_MIN"]; del os.environ["BLU_B200_DENSE_K"]


@pytest.mark.parametrize("seed", [100, 101, 104, 107, 110, 113, 117, 131])
def test_emu_row_search(emu, seed):
    """search_rows != 0 (markowitz.rs:125-189 on the device: markowitz_search_rows): columns and rows visited
    count by count in bucket order, both early exits, rows parked in bucket m+1 until re-stamped."""
    from parity import structured_matrix
    rng = np.random.default_rng(seed)
    m = int(rng.integers(20, 160))
    if seed % 3 == 0:
        cp, ri, v = gen.basis(seed, m, m // 4, 3.0 + seed % 4)
    else:
        cp, ri, v = structured_matrix(seed % 6, m, rng)
    ms = int(rng.integers(1, 6)); reltol = [0.1, 0.5, 1.0, 0.01][seed % 4]
    o = oracle_for(m, len(v), 400)
    for k, x in (("search_rows", 1), ("maxsearch", ms), ("reltol", reltol)):
        o.set_param(k, x)
    so = o.factorize(cp[:-1], cp[1:], ri, v)
    g = BLU(m, len(v), lib=emu)
    g.threads_per_basis = [32, 64, 128][seed % 3]; g.search_rows = 1; g.maxsearch = ms; g.reltol = reltol
    assert g.factorize(cp[:-1], cp[1:], ri, v) == so
    if so in (0, 2):
        assert_factor_parity(g, o)


def test_emu_multi_device_entry(emu):
    """blu_multi_*: one batch split into contiguous ranges over a device list (two device slots here, both the
    emulator); results land at the bases' own offsets and equal the oracle."""
    from blu_b200 import BLUMulti
    nmat, m = 5, 70
    bb, be, bi, bx, rhs = gen.batch(nmat, m, 20, 3.5, 9300, 9350)
    mb = BLUMulti(nmat, m, int((be - bb).reshape(nmat, m).sum(1).max()), [0, 0], lib=emu)
    assert [mb.part(d)[1:] for d in range(2)] == [(0, 3), (3, 2)]
    st, x, status = mb.factorize_solve(bb, be, bi, bx, rhs, "N")
    assert st == 0 and (status == 0).all()
    for k in range(nmat):
        cp, ri, v = gen.basis(9300 + k, m, 20, 3.5)
        o = oracle_for(m, len(v))
        assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
        _, xo = o.solve_dense(rhs[k * m:(k + 1) * m], "N")
        assert np.array_equal(x[k], xo), k
        assert mb.info(k, "rank") == o.info("rank") and mb.info(k, "l_nz") == o.info("l_nz")
    mb.close()


def test_emu_factorize_c0ntinue(emu):
    """The free-function protocol (lib.rs:11-19, factorize.rs:34-119, blu.rs:345-377 done by the caller):
    Reallocate escapes with addmem_*, the caller grows the stores and calls again with c0ntinue; c0ntinue
    without a pending Reallocate is ErrorInvalidCall (factorize.rs:102-105)."""
    m = 110
    cp, ri, v = gen.basis(77, m, 20, 5.0)
    o = oracle_for(m, len(v), 400)
    assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
    g = BLU(m, len(v), lib=emu)
    g.threads_per_basis = 64
    assert g.factorize_c0ntinue(cp[:-1], cp[1:], ri, v, True) == -2
    g.l_mem = len(v); g.u_mem = len(v); g.w_mem = len(v)
    st, rounds = g.factorize_c0ntinue(cp[:-1], cp[1:], ri, v, False), 0
    while st == 1:
        rounds += 1
        assert rounds < 30
        add = [g.info(n) for n in ("addmem_l", "addmem_u", "addmem_w")]
        assert max(add) > 0
        for name, a in zip(("l_mem", "u_mem", "w_mem"), add):
            if a > 0:
                setattr(g, name, int(1.5 * (g.get_param(name) + a)) + 1)      # lu_realloc_obj, blu.rs:345-377
        st = g.factorize_c0ntinue(cp[:-1], cp[1:], ri, v, True)
    assert st == 0 and rounds >= 1
    assert_factor_parity(g, o, check_stats=False)
    assert g.factorize_c0ntinue(cp[:-1], cp[1:], ri, v, True) == -2      # nothing pending any more


def test_emu_batch_hunt_sample():
    """A few random batches through the split factorization under random dense-tail stage orders, CTA sizes and
    tunables (scripts/batch_hunt.py --emu): every sampled basis equals the oracle."""
    import subprocess
    import sys
    root = os.path.dirname(HERE)
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "batch_hunt.py"), "47000", "6", "--emu"],
                         capture_output=True, text=True, timeout=900).stdout
    assert out.strip().splitlines()[-1].startswith("6 batches, 0 failures"), out[-1500:]

