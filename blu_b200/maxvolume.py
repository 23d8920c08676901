"""maxvolume driver (reference: src/maxvolume.rs:64-224): one pass over the columns of A, replacing a
basis position whenever |B^{-1} a_j| has an entry larger than `volumetol`.  It is the only in-tree
consumer of solve_for_update + update; here it is host code above the C ABI, exactly as the reference's
is host code above the LU kernels."""
import numpy as np

from .blu import Status


def _factorize(obj, a_p, a_i, a_x, basis):
    """maxvolume.rs:180-224: factorize the columns of A listed in `basis`"""
    begin = a_p[basis]
    end = a_p[basis + 1]
    return obj.factorize(begin, end, a_i, a_x)


def maxvolume(obj, ncol, a_p, a_i, a_x, basis, isbasic, volumetol):
    """Returns (status, nupdate).  `basis` (m column indices) and `isbasic` (ncol flags) are updated in place."""
    a_p = np.ascontiguousarray(a_p, dtype=np.int64)
    a_i = np.ascontiguousarray(a_i, dtype=np.int64)
    a_x = np.ascontiguousarray(a_x, dtype=np.float64)
    nupdate = 0
    if volumetol < 1.0:                                          # maxvolume.rs:83-90
        return Status.ERROR_INVALID_ARGUMENT, nupdate
    st = _factorize(obj, a_p, a_i, a_x, basis)
    if st != Status.OK:
        return st, nupdate
    m = obj.m
    for j in range(ncol):
        if isbasic[j]:
            continue
        b, e = a_p[j], a_p[j + 1]
        st = obj.solve_for_update(e - b, a_i[b:e], a_x[b:e], "N", 1)   # B^{-1} a_j, maxvolume.rs:111
        if st != Status.OK:
            return st, nupdate
        # first maximum in the order of ilhs (strict >), maxvolume.rs:120-131
        xmax, xtbl, imax = 0.0, 0.0, 0
        pat = obj.ilhs[:obj.nzlhs]
        vals = np.abs(obj.lhs[pat])
        if len(vals):
            k = int(np.argmax(vals))                             # argmax returns the first maximum
            if vals[k] > 0.0:
                imax = int(pat[k]); xtbl = float(obj.lhs[imax]); xmax = abs(xtbl)
        if xmax <= volumetol:
            continue
        isbasic[basis[imax]] = 0
        isbasic[j] = 1
        basis[imax] = j
        nupdate += 1
        st = obj.solve_for_update(0, np.array([imax], dtype=np.int64), None, "T", 0)   # maxvolume.rs:145
        if st != Status.OK:
            return st, nupdate
        st = obj.update(xtbl)                                    # maxvolume.rs:154
        if st != Status.OK:
            return st, nupdate
        # refactorization policy, maxvolume.rs:161-170
        if obj.info("nforrest") == m or obj.info("pivot_error") > 1e-8 or obj.info("update_cost") > 1.0:
            st = _factorize(obj, a_p, a_i, a_x, basis)
            if st != Status.OK:
                return st, nupdate
    return Status.OK, nupdate
