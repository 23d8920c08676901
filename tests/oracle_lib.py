"""ctypes wrapper of the CPU oracle (oracle/libblo.so).  Test infrastructure only."""
import ctypes
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ODIR = os.path.join(_ROOT, "oracle")
i64p = ctypes.POINTER(ctypes.c_int64)
f64p = ctypes.POINTER(ctypes.c_double)
_L = None

P = dict(droptol=0, abstol=1, reltol=2, nzbias=3, maxsearch=4, pad=5, stretch=6, compress_thres=7,
         sparse_thres=8, search_rows=9, check_file_diff=10, realloc_factor=11)
_INFO = ["m", "rank", "bump_size", "bump_nz", "matrix_nz", "l_nz", "u_nz", "r_nz", "nsearch_pivot", "nexpand",
         "ngarbage", "factor_flops", "min_pivot", "max_pivot", "max_eta", "nupdate", "nforrest", "nfactorize",
         "nupdate_total", "nforrest_total", "nsymperm_total", "l_flops", "u_flops", "r_flops", "condest_l",
         "condest_u", "norm_l", "norm_u", "normest_l_inv", "normest_u_inv", "onenorm", "infnorm", "residual_test",
         "pivot_error", "update_cost", "time_factorize", "time_solve", "time_update", "time_singletons",
         "time_search_pivot", "time_elim_pivot", "l_mem", "u_mem", "w_mem", "nzlhs", "elim_bytes", "nelim_div",
         "pivotlen", "rankdef"]
I = {n: 100 + k for k, n in enumerate(_INFO)}


class Trace(ctypes.Structure):
    _fields_ = [("row", ctypes.c_int64), ("col", ctypes.c_int64), ("pivot", ctypes.c_double),
                ("kind", ctypes.c_int), ("nz_row", ctypes.c_int64), ("nz_col", ctypes.c_int64)]


def lib():
    global _L
    if _L is None:
        so = os.path.join(_ODIR, "libblo.so")
        subprocess.check_call(["make", "-s", "-C", _ODIR])
        L = ctypes.CDLL(so)
        vp = ctypes.c_void_p
        L.blo_new.restype = vp; L.blo_new.argtypes = [ctypes.c_int64, ctypes.c_int64]
        L.blo_free.argtypes = [vp]
        L.blo_factorize.argtypes = [vp, i64p, i64p, i64p, f64p]
        L.blo_solve_dense.argtypes = [vp, f64p, f64p, ctypes.c_char]
        L.blo_solve_sparse.argtypes = [vp, ctypes.c_int64, i64p, f64p, ctypes.c_char]
        L.blo_solve_for_update.argtypes = [vp, ctypes.c_int64, i64p, f64p, ctypes.c_char, ctypes.c_int64]
        L.blo_update.argtypes = [vp, ctypes.c_double]
        L.blo_get_factors.argtypes = [vp, i64p, i64p, i64p, i64p, f64p, i64p, i64p, f64p]
        L.blo_maxvolume.argtypes = [vp, ctypes.c_int64, i64p, i64p, f64p, i64p, i64p, ctypes.c_double, i64p]
        L.blo_get_info.restype = ctypes.c_double; L.blo_get_info.argtypes = [vp, ctypes.c_int]
        L.blo_set_param.argtypes = [vp, ctypes.c_int, ctypes.c_double]
        L.blo_trace_enable.argtypes = [vp, ctypes.c_int]
        L.blo_trace_len.restype = ctypes.c_int64; L.blo_trace_len.argtypes = [vp]
        L.blo_trace_data.restype = ctypes.POINTER(Trace); L.blo_trace_data.argtypes = [vp]
        L.blo_lhs.restype = f64p; L.blo_lhs.argtypes = [vp]
        L.blo_ilhs.restype = i64p; L.blo_ilhs.argtypes = [vp]
        L.blo_batch_factorize_solve.argtypes = [ctypes.c_int64, ctypes.c_int64, i64p, i64p, i64p, f64p, f64p, f64p,
                                                ctypes.c_char, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                                ctypes.POINTER(ctypes.c_int)]
        _L = L
    return _L


def batch_factorize_solve(nmat, m, b_begin, b_end, b_i, b_x, rhs, trans="N", store_nz=0, nthreads=1, check_file_diff=1):
    """One oracle instance per thread over a batch (oracle/blo_batch.c).  Returns (threads, lhs[nmat,m], status)."""
    L = lib()
    bb = np.ascontiguousarray(b_begin, dtype=np.int64); be = np.ascontiguousarray(b_end, dtype=np.int64)
    bi = np.ascontiguousarray(b_i, dtype=np.int64); bx = np.ascontiguousarray(b_x, dtype=np.float64)
    r = np.ascontiguousarray(rhs, dtype=np.float64).reshape(-1)
    x = np.zeros(nmat * m)
    status = np.zeros(nmat, dtype=np.int32)
    if store_nz <= 0:
        store_nz = int((be - bb).reshape(nmat, m).sum(1).max())
    nt = L.blo_batch_factorize_solve(nmat, m, _pi(bb), _pi(be), _pi(bi), _pf(bx), _pf(r), _pf(x), str(trans).encode()[:1],
                                     int(store_nz), int(nthreads), int(check_file_diff),
                                     status.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    return nt, x.reshape(nmat, m), status


def _pi(a):
    return a.ctypes.data_as(i64p) if a is not None else None


def _pf(a):
    return a.ctypes.data_as(f64p) if a is not None else None


class Oracle:
    """Same method names as blu_b200.BLU so parity tests read symmetrically."""

    def __init__(self, m, b_nz):
        self._L = lib()
        self._h = self._L.blo_new(int(m), int(b_nz))
        self.m = int(m)

    def __del__(self):
        try:
            if self._h:
                self._L.blo_free(self._h)
                self._h = None
        except Exception:
            pass

    def set_param(self, name, v):
        self._L.blo_set_param(self._h, P[name], float(v))

    def info(self, name):
        return self._L.blo_get_info(self._h, I[name])

    def trace(self, on=True):
        self._L.blo_trace_enable(self._h, 1 if on else 0)

    def get_trace(self):
        n = self._L.blo_trace_len(self._h)
        d = self._L.blo_trace_data(self._h)
        return [(d[k].row, d[k].col, d[k].pivot, d[k].kind, d[k].nz_row, d[k].nz_col) for k in range(n)]

    def factorize(self, b_begin, b_end, b_i, b_x):
        bb = np.ascontiguousarray(b_begin, dtype=np.int64); be = np.ascontiguousarray(b_end, dtype=np.int64)
        bi = np.ascontiguousarray(b_i, dtype=np.int64); bx = np.ascontiguousarray(b_x, dtype=np.float64)
        return self._L.blo_factorize(self._h, _pi(bb), _pi(be), _pi(bi), _pf(bx))

    def solve_dense(self, rhs, trans="N"):
        r = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.zeros(self.m)
        st = self._L.blo_solve_dense(self._h, _pf(r), _pf(x), str(trans).encode()[:1])
        return st, x

    def solve_sparse(self, nzrhs, irhs, xrhs, trans="N"):
        ir = np.ascontiguousarray(irhs, dtype=np.int64); xr = np.ascontiguousarray(xrhs, dtype=np.float64)
        st = self._L.blo_solve_sparse(self._h, int(nzrhs), _pi(ir), _pf(xr), str(trans).encode()[:1])
        return st

    def solve_for_update(self, nzrhs, irhs, xrhs, trans="N", want_solution=0):
        ir = np.ascontiguousarray(irhs, dtype=np.int64)
        xr = np.ascontiguousarray(xrhs, dtype=np.float64) if xrhs is not None else None
        return self._L.blo_solve_for_update(self._h, int(nzrhs), _pi(ir), _pf(xr), str(trans).encode()[:1], int(want_solution))

    def update(self, xtbl):
        return self._L.blo_update(self._h, float(xtbl))

    def maxvolume(self, ncol, a_p, a_i, a_x, basis, isbasic, volumetol):
        """maxvolume.rs:64-224; basis / isbasic (int64 arrays) are updated in place.  Returns (status, nupdate)."""
        ap = np.ascontiguousarray(a_p, dtype=np.int64); ai = np.ascontiguousarray(a_i, dtype=np.int64)
        ax = np.ascontiguousarray(a_x, dtype=np.float64)
        nup = ctypes.c_int64(0)
        st = self._L.blo_maxvolume(self._h, int(ncol), _pi(ap), _pi(ai), _pf(ax), _pi(basis), _pi(isbasic), float(volumetol), ctypes.byref(nup))
        return st, nup.value

    @property
    def nzlhs(self):
        return int(self.info("nzlhs"))

    @property
    def lhs(self):
        return np.ctypeslib.as_array(self._L.blo_lhs(self._h), shape=(self.m,)).copy()

    @property
    def ilhs(self):
        return np.ctypeslib.as_array(self._L.blo_ilhs(self._h), shape=(self.m,)).copy()

    def get_factors(self):
        m = self.m
        lnz, unz = int(self.info("l_nz")), int(self.info("u_nz"))
        out = dict(rowperm=np.zeros(m, np.int64), colperm=np.zeros(m, np.int64),
                   l_colptr=np.zeros(m + 1, np.int64), l_rowidx=np.zeros(m + lnz, np.int64), l_value=np.zeros(m + lnz),
                   u_colptr=np.zeros(m + 1, np.int64), u_rowidx=np.zeros(m + unz, np.int64), u_value=np.zeros(m + unz))
        st = self._L.blo_get_factors(self._h, _pi(out["rowperm"]), _pi(out["colperm"]),
                                     _pi(out["l_colptr"]), _pi(out["l_rowidx"]), _pf(out["l_value"]),
                                     _pi(out["u_colptr"]), _pi(out["u_rowidx"]), _pf(out["u_value"]))
        return st, out
