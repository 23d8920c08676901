"""Shared parity helpers: compare the CUDA path with the CPU oracle on the same input."""
import numpy as np
from oracle_lib import Oracle

STATS = ["rank", "bump_size", "bump_nz", "matrix_nz", "l_nz", "u_nz", "factor_flops", "nsearch_pivot",
         "min_pivot", "max_pivot", "nelim_div", "elim_bytes",      # elim_bytes: the numerator of bench.py's roofline
         # the tail of factorize (factorize.rs:121-147): condest x2, residual_test, matrix_norm
         "condest_l", "condest_u", "norm_l", "norm_u", "normest_l_inv", "normest_u_inv", "onenorm", "infnorm",
         "residual_test", "update_cost"]


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


SPARSE_STATS = ["nupdate", "nforrest", "u_nz", "r_nz", "l_flops", "u_flops", "r_flops", "pivot_error", "max_eta",
                "min_pivot", "max_pivot", "nsymperm_total", "nforrest_total", "nupdate_total", "pivotlen", "update_cost"]


def assert_same_solution(g, o, what=""):
    """lhs / ilhs / nzlhs of the last sparse solve: pattern ORDER and values bit-exact."""
    assert g.nzlhs == o.nzlhs, (what, g.nzlhs, o.nzlhs)
    n = o.nzlhs
    assert np.array_equal(g.ilhs[:n], o.ilhs[:n]), f"{what}: ilhs order differs"
    assert np.array_equal(g.lhs, o.lhs), f"{what}: lhs differs (max {np.abs(g.lhs - o.lhs).max():.3e})"


def assert_sparse_stats(g, o, what=""):
    for n in SPARSE_STATS:
        assert g.info(n) == o.info(n), (what, n, g.info(n), o.info(n))
    assert g.info("internal_error") == 0


def assert_sparse_solve_parity(g, o, m, seed, sizes=(1, 3, 17, 60)):
    """solve_sparse 'N' and 'T' for RHS of several densities (both sides of sparse_thres)."""
    from blu_b200 import gen
    for k, nz in enumerate(sizes):
        nz = min(nz, m)
        idx, val = gen.sparse_rhs(seed + k, m, nz)
        for tr in "NT":
            so = o.solve_sparse(nz, idx, val, tr)
            sg = g.solve_sparse(nz, idx, val, tr)
            assert so == sg == 0, (nz, tr, so, sg)
            assert_same_solution(g, o, f"solve_sparse nz={nz} trans={tr}")
    assert_sparse_stats(g, o, "after solve_sparse")


def replay_updates(g, o, m, pool, niter, rng_seed=5, check_dense=True):
    """Simplex-style replay on both objects in lockstep (C5 of BASELINE.json in miniature):
    entering column from the pool, leaving position = argmax |lhs| (maxvolume.rs:120-131),
    solve_for_update 'N' + 'T', update; everything observable must be identical."""
    pool_cp, pool_ri, pool_v = pool
    rng = np.random.default_rng(rng_seed)
    kinds = set()
    for it in range(niter):
        idx = pool_ri[pool_cp[it]:pool_cp[it + 1]]
        val = pool_v[pool_cp[it]:pool_cp[it + 1]]
        so = o.solve_for_update(len(idx), idx, val, "N", want_solution=1)
        sg = g.solve_for_update(len(idx), idx, val, "N", want_solution=1)
        assert so == sg, (it, so, sg)
        if so != 0:      # e.g. after a refused (singular) update: both sides now want a fresh factorization
            return kinds
        assert_same_solution(g, o, f"it {it} ftran")
        lhs = o.lhs
        j = int(np.argmax(np.abs(lhs)))
        xtbl = lhs[j]
        want = it % 2
        so = o.solve_for_update(1, np.array([j]), None, "T", want_solution=want)
        sg = g.solve_for_update(1, np.array([j]), None, "T", want_solution=want)
        assert so == sg == 0, (it, so, sg)
        if want:
            assert_same_solution(g, o, f"it {it} btran")
        nf0 = o.info("nforrest")
        so = o.update(xtbl)
        sg = g.update(xtbl)
        assert so == sg, (it, so, sg)
        assert_sparse_stats(g, o, f"it {it} update")
        if so != 0:
            continue
        kinds.add("ft" if o.info("nforrest") > nf0 else "perm")
        if check_dense:
            b = rng.uniform(-1, 1, m)
            for tr in "NT":
                _, xo = o.solve_dense(b, tr)
                sg, xg = g.solve_dense(b, tr)
                assert sg == 0
                assert np.array_equal(xg, xo), (it, tr, np.abs(xg - xo).max() / np.abs(xo).max())
    return kinds


def assert_maxvolume_parity(g, o, m, ncol, seed, volumetol=1.0):
    """maxvolume.rs:64-224 on both sides: same basis, same number of updates, same final solves."""
    from blu_b200 import gen, maxvolume
    # A = [slack-heavy starting basis | ncol - m candidate columns]
    cp, ri, v = gen.basis(seed, m, m // 2, 3.0)
    pcp, pri, pv = gen.basis(seed + 1, m, 0, 3.0)
    extra = ncol - m
    a_p = np.concatenate([cp, cp[-1] + pcp[1:extra + 1]])
    a_i = np.concatenate([ri, pri[:pcp[extra]]])
    a_x = np.concatenate([v, 3.0 * pv[:pcp[extra]]])       # candidates large enough to enter the basis
    bo, bg = np.arange(m, dtype=np.int64), np.arange(m, dtype=np.int64)
    io = np.zeros(ncol, dtype=np.int64); io[:m] = 1
    ig = io.copy()
    so, no = o.maxvolume(ncol, a_p, a_i, a_x, bo, io, volumetol)
    sg, ng = maxvolume(g, ncol, a_p, a_i, a_x, bg, ig, volumetol)
    assert (so, no) == (sg, ng), (so, no, sg, ng)
    assert np.array_equal(bo, bg) and np.array_equal(io, ig)
    for n in ("nupdate", "nforrest", "nfactorize", "nupdate_total", "nforrest_total", "nsymperm_total", "pivot_error"):
        assert g.info(n) == o.info(n), n
    b = gen.rhs(seed + 2, m)
    for tr in "NT":
        _, xo = o.solve_dense(b, tr)
        _, xg = g.solve_dense(b, tr)
        assert np.array_equal(xg, xo), tr
    return no


def assert_sparse_multi_parity(g, o, m, seed, sizes=(1, 3, 17, 60), reps=3):
    """blu_solve_sparse_multi == that many solve_sparse calls (pattern order and values bit-exact)."""
    from blu_b200 import gen
    for tr in "NT":
        rhs = []
        for rep in range(reps):
            for k, nz in enumerate(sizes):
                rhs.append(gen.sparse_rhs(seed + 10 * rep + k, m, min(nz, m)))
        st, out, stat = g.solve_sparse_multi(rhs, tr)
        assert st == 0 and (stat == 0).all()
        for (idx, val), (il, xl) in zip(rhs, out):
            assert o.solve_sparse(len(idx), idx, val, tr) == 0
            n = o.nzlhs
            assert len(il) == n and np.array_equal(il, o.ilhs[:n]), "pattern order"
            assert np.array_equal(xl, o.lhs[o.ilhs[:n]]), "values"
    # a bad right-hand side is reported for that unit only
    st, out, stat = g.solve_sparse_multi([(np.array([0]), np.array([1.0])), (np.array([m]), np.array([1.0]))], "N")
    assert st == -4 and list(stat) == [0, -4] and len(out[1][0]) == 0
    assert o.solve_sparse(1, np.array([0]), np.array([1.0]), "N") == 0      # keep the flop counters of both sides in step


def batch_replay_parity(b, oracles, m, pools, niter):
    """blu_batch_solve_for_update + blu_batch_update: every basis of the batch advances one column
    replacement per round, in lockstep with one oracle per basis; everything observable must be identical."""
    nmat = len(oracles)
    for it in range(niter):
        cols = [(pools[k][1][pools[k][0][it]:pools[k][0][it + 1]], pools[k][2][pools[k][0][it]:pools[k][0][it + 1]]) for k in range(nmat)]
        st, stat, out = b.solve_for_update(cols, "N", want_solution=1)
        assert st == 0 and (stat == 0).all(), (it, st, stat)
        leave, xt = [], []
        for k, o in enumerate(oracles):
            assert o.solve_for_update(len(cols[k][0]), cols[k][0], cols[k][1], "N", want_solution=1) == 0
            n = o.nzlhs
            assert np.array_equal(out[k][0], o.ilhs[:n]) and np.array_equal(out[k][1], o.lhs[o.ilhs[:n]]), (it, k)
            j = int(np.argmax(np.abs(o.lhs)))
            leave.append((np.array([j]), None)); xt.append(o.lhs[j])
        st, stat, _ = b.solve_for_update(leave, "T", want_solution=0)
        assert st == 0 and (stat == 0).all()
        st, stat = b.update(np.array(xt))
        for k, o in enumerate(oracles):
            assert o.solve_for_update(1, leave[k][0], None, "T", want_solution=0) == 0
            assert o.update(xt[k]) == stat[k], (it, k, stat[k])
            for name in ("nupdate", "nforrest", "u_nz", "r_nz", "pivot_error", "max_eta", "l_flops", "u_flops", "r_flops"):
                assert b.info(k, name) == o.info(name), (it, k, name)


TUNABLES = dict(            # the `pub` tunables of LU (lu.rs:10-66), values on both sides of the defaults
    droptol=[1e-20, 1e-12, 1e-6, 1e-2],
    abstol=[1e-14, 1e-8, 1e-3],
    reltol=[0.1, 0.01, 0.5, 0.9, 1.0],
    nzbias=[0, -1, 1, -3],
    maxsearch=[1, 2, 3, 4, 8],
    pad=[1, 2, 4, 8],
    stretch=[0.1, 0.3, 1.0],
    compress_thres=[0.05, 0.5, 1.0],
    sparse_thres=[0.0, 0.05, 0.5, 1.0],
    search_rows=[0, 0, 1],      # markowitz.rs:125-189; the crate's default is 0 (D8)
)


def draw_tunables(seed):
    rng = np.random.default_rng(seed)
    return {k: v[int(rng.integers(len(v)))] for k, v in TUNABLES.items()}


def tunables_case(make_gpu, m, seed, nupd=12, dens=4.0, matrix=None, pool=None):
    """One random setting of every tunable on both objects, then the whole public surface in lockstep:
    factorize, get_factors, dense + sparse solves, a few column replacements.  Returns the setting."""
    from blu_b200 import gen
    t = draw_tunables(seed)
    cp, ri, v = matrix if matrix is not None else gen.basis(seed, m, m // 3, dens)
    pool = pool if pool is not None else gen.basis(seed + 1, m, 0, 3.0)
    o = oracle_for(m, len(v), 400)
    g = make_gpu(m, len(v))
    for k, x in t.items():
        o.set_param(k, x)
        setattr(g, k, x)
    so = o.factorize(cp[:-1], cp[1:], ri, v)
    sg = g.factorize(cp[:-1], cp[1:], ri, v)
    assert so == sg, (t, so, sg)
    if so not in (0, 2):
        return t
    assert_factor_parity(g, o)
    b = gen.rhs(seed + 2, m)
    for tr in "NT":
        _, xo = o.solve_dense(b, tr)
        sg, xg = g.solve_dense(b, tr)
        assert sg == 0 and np.array_equal(xg, xo), (t, tr)
    assert_sparse_solve_parity(g, o, m, seed + 3, sizes=(1, 4, 30))
    if so == 0:
        replay_updates(g, o, m, pool, nupd, check_dense=False)
        for tr in "NT":
            _, xo = o.solve_dense(b, tr)
            sg, xg = g.solve_dense(b, tr)
            assert sg == 0 and np.array_equal(xg, xo), (t, tr, "after updates")
    return t


def _csc(A):
    import scipy.sparse as sp
    A = sp.csc_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64)


def structured_matrix(kind, m, rng):
    """Structures the synthetic bases of blu_b200.gen do not produce (see scripts/structure_hunt.py)."""
    import scipy.sparse as sp
    if kind == 0:      # +-1 entries: exact cancellation, often singular
        A = sp.random(m, m, density=min(1.0, 3.5 / m), format="csc", random_state=rng, data_rvs=lambda n: np.sign(rng.uniform(-1, 1, n)))
        A = A + sp.diags(np.where(rng.uniform(size=m) < 0.8, 1.0, 0.0))
    elif kind == 1:    # permuted triangle: singletons only
        A = sp.tril(sp.random(m, m, density=min(1.0, 4.0 / m), format="csc", random_state=rng)) + sp.diags(rng.uniform(0.5, 2, m))
        A = sp.csc_matrix(A)[rng.permutation(m), :][:, rng.permutation(m)]
    elif kind == 2:    # arrowhead + sparse noise
        A = sp.lil_matrix((m, m))
        A.setdiag(rng.uniform(0.5, 2, m))
        A[0, :] = rng.uniform(-1, 1, m)
        A[:, 0] = rng.uniform(-1, 1, (m, 1))
        A = sp.csc_matrix(A) + sp.random(m, m, density=min(1.0, 1.0 / m), format="csc", random_state=rng)
    elif kind == 3:    # dense block inside a sparse matrix
        k = min(m, int(rng.integers(5, 40)))
        A = sp.lil_matrix(sp.random(m, m, density=min(1.0, 2.5 / m), format="csc", random_state=rng) + sp.diags(rng.uniform(0.5, 2, m)))
        at = int(rng.integers(0, m - k + 1))
        A[at:at + k, at:at + k] = rng.uniform(-1, 1, (k, k))
        A = sp.csc_matrix(A)[rng.permutation(m), :]
    elif kind == 4:    # badly scaled: entries from 1e-16 to 1e4, some below abstol
        A = sp.random(m, m, density=min(1.0, 4.0 / m), format="csc", random_state=rng, data_rvs=lambda n: rng.uniform(-1, 1, n) * 10.0 ** rng.integers(-16, 5, n))
        A = A + sp.diags(10.0 ** rng.integers(-12, 3, m).astype(float))
    else:              # empty rows / columns, duplicate-free but structurally singular
        A = sp.lil_matrix(sp.random(m, m, density=min(1.0, 3.0 / m), format="csc", random_state=rng) + sp.diags(rng.uniform(0.5, 2, m)))
        for j in rng.integers(0, m, max(1, m // 20)):
            A[:, j] = 0
        for i in rng.integers(0, m, max(1, m // 25)):
            A[i, :] = 0
        A = sp.csc_matrix(A)
        A.eliminate_zeros()
    return _csc(A)


def structured_case(make_gpu, m, seed, nupd=12):
    """tunables_case on structured_matrix(seed % 6): sign matrices that cancel exactly, permuted triangles,
    arrowheads, dense blocks, badly scaled entries, empty rows/columns; entering columns from a random pool."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    M = structured_matrix(seed % 6, m, rng)
    nc = max(40, nupd)
    pool = _csc(sp.random(m, nc, density=min(1.0, 3.0 / m) + 0.01, format="csc", random_state=rng)
                + sp.csc_matrix((np.ones(nc), (rng.integers(0, m, nc), np.arange(nc))), shape=(m, nc)))
    return tunables_case(make_gpu, m, seed, nupd=nupd, matrix=M, pool=pool)


def batch_structured_case(make_batch, m, seed0):
    """One batch holding five different structures (structured_matrix kinds 0,1,2,3,5) under one random
    setting of the tunables: status, factors, counters and dense solves of every basis equal those of an
    oracle instance given the same basis alone."""
    t = draw_tunables(seed0)
    mats = [structured_matrix(kind, m, np.random.default_rng(seed0 + k)) for k, kind in enumerate([0, 1, 2, 3, 5])]
    n = len(mats)
    off, bb, be = 0, [], []
    for cp, ri, v in mats:
        bb.append(cp[:-1] + off); be.append(cp[1:] + off); off += len(v)
    bb, be = np.concatenate(bb), np.concatenate(be)
    bi, bx = np.concatenate([x[1] for x in mats]), np.concatenate([x[2] for x in mats])
    b = make_batch(n, m, max(len(x[2]) for x in mats))
    for a, x in t.items():
        setattr(b, a, x)
    st, status = b.factorize(bb, be, bi, bx)
    assert st == 0
    rhs = np.random.default_rng(seed0).uniform(-1, 1, n * m)
    outs = {tr: np.asarray(b.solve_dense(rhs, tr)[1]).reshape(-1) for tr in "NT"}
    for k, (cp, ri, v) in enumerate(mats):
        o = oracle_for(m, len(v), 400)
        for a, x in t.items():
            o.set_param(a, x)
        so = o.factorize(cp[:-1], cp[1:], ri, v)
        assert so == status[k], (k, so, status[k])
        if so not in (0, 2):
            continue
        _, fo = o.get_factors()
        stg, fg = b.get_factors(k)
        assert stg == 0
        for name in fo:
            assert np.array_equal(fo[name], fg[name]), (k, name)
        for name in STATS:
            assert o.info(name) == b.info(k, name), (k, name)
        for tr in "NT":
            _, xo = o.solve_dense(rhs[k * m:(k + 1) * m], tr)
            assert np.array_equal(outs[tr][k * m:(k + 1) * m], xo), (k, tr)
