"""Host-side mirror of the reference's public surface (src/blu.rs, src/lib.rs:11-64).

``BLU(m, b_nz)`` has the methods of ``impl BLU`` (blu.rs:61-334) with the same argument
meaning; instead of ``Result<(), Status>`` every method returns the integer status
(``Status.OK`` == ``Ok(())``), numbered as in include/blu_b200.h.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

i64p = ctypes.POINTER(ctypes.c_int64)
f64p = ctypes.POINTER(ctypes.c_double)
i32p = ctypes.POINTER(ctypes.c_int)


class Status:
    OK = 0
    REALLOCATE = 1
    WARNING_SINGULAR_MATRIX = 2
    ERROR_INVALID_CALL = -2
    ERROR_ARGUMENT_MISSING = -3
    ERROR_INVALID_ARGUMENT = -4
    ERROR_MAXIMUM_UPDATES = -5
    ERROR_SINGULAR_UPDATE = -6
    ERROR_INTERNAL = -100
    ERROR_CUDA = -101
    ERROR_OUT_OF_MEMORY = -102


# selectors (include/blu_b200.h)
P = dict(droptol=0, abstol=1, reltol=2, nzbias=3, maxsearch=4, pad=5, stretch=6, compress_thres=7,
         sparse_thres=8, search_rows=9, realloc_factor=10, l_mem=11, u_mem=12, w_mem=13,
         threads_per_basis=14, norms=15, dense_k=16, tail_threads=17, split_min=18, tree_min=19, dense_k_big=20)
_INFO_NAMES = ["m", "rank", "bump_size", "bump_nz", "matrix_nz", "l_nz", "u_nz", "r_nz", "nsearch_pivot",
               "nexpand", "ngarbage", "factor_flops", "min_pivot", "max_pivot", "max_eta", "nupdate",
               "nforrest", "nfactorize", "nupdate_total", "nforrest_total", "nsymperm_total", "l_flops",
               "u_flops", "r_flops", "condest_l", "condest_u", "norm_l", "norm_u", "normest_l_inv",
               "normest_u_inv", "onenorm", "infnorm", "residual_test", "pivot_error", "update_cost",
               "time_factorize", "time_solve", "time_update", "elim_bytes", "nelim_div", "pivotlen",
               "rankdef", "internal_error", "status", "nrealloc", "elim_bytes_head", "nruns", "addmem_l", "addmem_u", "addmem_w"]
I = {n: 100 + k for k, n in enumerate(_INFO_NAMES)}
I.update({f"t_phase{q}": 200 + q for q in range(16)})
I.update({f"n_kind{q}": 220 + q for q in range(8)})
I.update({f"norms_cyc{q}": 230 + q for q in range(16)})


def library_path():
    return os.environ.get("BLU_B200_LIB", os.path.join(_HERE, "libblu_b200.so"))


def load_library(path=None):
    """Load the C-ABI library.  Raises if it is missing: there is no fallback."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or library_path()
    if not os.path.exists(path):
        raise RuntimeError(f"blu_b200: CUDA library not built: {path} (run __graft_entry__.build())")
    L = ctypes.CDLL(path)
    vp = ctypes.c_void_p
    L.blu_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
    L.blu_destroy.argtypes = [vp]
    L.blu_set_param.argtypes = [vp, ctypes.c_int, ctypes.c_double]
    L.blu_get_param.argtypes = [vp, ctypes.c_int]; L.blu_get_param.restype = ctypes.c_double
    L.blu_get_info.argtypes = [vp, ctypes.c_int]; L.blu_get_info.restype = ctypes.c_double
    L.blu_factorize.argtypes = [vp, i64p, i64p, i64p, f64p]
    L.blu_get_factors.argtypes = [vp, i64p, i64p, i64p, i64p, f64p, i64p, i64p, f64p]
    L.blu_factorize_c0ntinue.argtypes = [vp, i64p, i64p, i64p, f64p, ctypes.c_int]
    L.blu_solve_dense.argtypes = [vp, f64p, f64p, ctypes.c_char]
    L.blu_solve_dense_multi.argtypes = [vp, ctypes.c_int64, f64p, f64p, ctypes.c_char]
    L.blu_solve_sparse_multi.argtypes = [vp, ctypes.c_int64, i64p, i64p, f64p, i64p, i64p, f64p, i32p, ctypes.c_char]
    L.blu_solve_sparse.argtypes = [vp, ctypes.c_int64, i64p, f64p, i64p, i64p, f64p, ctypes.c_char]
    L.blu_solve_for_update.argtypes = [vp, ctypes.c_int64, i64p, f64p, i64p, i64p, f64p, ctypes.c_char]
    L.blu_update.argtypes = [vp, ctypes.c_double]
    L.blu_batch_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
    L.blu_batch_destroy.argtypes = [vp]
    L.blu_batch_factorize.argtypes = [vp, i64p, i64p, i64p, f64p, ctypes.c_int64, i32p]
    L.blu_batch_solve_dense.argtypes = [vp, f64p, f64p, ctypes.c_char, i32p]
    L.blu_batch_solve_for_update.argtypes = [vp, i64p, i64p, f64p, ctypes.c_int, i64p, i64p, f64p, i32p, ctypes.c_char]
    L.blu_batch_update.argtypes = [vp, f64p, i32p]
    L.blu_batch_get_info.argtypes = [vp, ctypes.c_int64, ctypes.c_int]; L.blu_batch_get_info.restype = ctypes.c_double
    L.blu_batch_get_factors.argtypes = [vp, ctypes.c_int64, i64p, i64p, i64p, i64p, f64p, i64p, i64p, f64p]
    L.blu_batch_upload.argtypes = [vp, i64p, i64p, i64p, f64p, ctypes.c_int64, f64p]
    L.blu_batch_factorize_resident.argtypes = [vp]
    L.blu_batch_solve_dense_resident.argtypes = [vp, ctypes.c_char]
    L.blu_batch_download.argtypes = [vp, f64p, i32p]
    L.blu_batch_factorize_dev.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int64]
    L.blu_batch_solve_dense_dev.argtypes = [vp, vp, vp, ctypes.c_char, vp]
    L.blu_batch_graph_capture.argtypes = [vp, ctypes.c_char]
    L.blu_batch_graph_launch.argtypes = [vp]
    L.blu_batch_stream.argtypes = [vp]; L.blu_batch_stream.restype = vp
    L.blu_batch_set_stream.argtypes = [vp, vp]
    L.blu_batch_synchronize.argtypes = [vp]
    L.blu_batch_last_kernel_ms.argtypes = [vp, ctypes.c_int]; L.blu_batch_last_kernel_ms.restype = ctypes.c_double
    L.blu_batch_launch_count.argtypes = [vp]; L.blu_batch_launch_count.restype = ctypes.c_int64
    L.blu_multi_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, i32p, ctypes.c_int]
    L.blu_multi_destroy.argtypes = [vp]
    L.blu_multi_factorize_solve.argtypes = [vp, i64p, i64p, i64p, f64p, ctypes.c_int64, f64p, f64p, ctypes.c_char, i32p]
    L.blu_multi_part.argtypes = [vp, ctypes.c_int, i64p, i64p]; L.blu_multi_part.restype = vp
    L.blu_version.restype = ctypes.c_char_p
    if path == library_path():
        _LIB = L
    return L


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _pi(a):
    return a.ctypes.data_as(i64p) if a is not None else None


def _pf(a):
    return a.ctypes.data_as(f64p) if a is not None else None


def _ch(trans):
    return (trans if isinstance(trans, bytes) else str(trans).encode())[:1]


class _Base:
    _get_info_fn = "blu_get_info"

    def set_param(self, name, value):
        return self._L.blu_set_param(self._h, P[name], float(value))

    def get_param(self, name):
        return self._L.blu_get_param(self._h, P[name])

    def __setattr__(self, name, value):
        # the tunables of `LU` are plain pub fields in the reference (lu.rs:10-66)
        if name in P and "_h" in self.__dict__:
            st = self.set_param(name, value)
            if st != 0:
                raise ValueError(f"set {name}={value}: status {st}")
        else:
            object.__setattr__(self, name, value)


class BLU(_Base):
    """struct BLU, blu.rs:9-20: `lu` state on the device, `lhs/ilhs/nzlhs` on the host."""

    def __init__(self, m, b_nz, device=-1, lib=None):
        self._L = lib or load_library()
        h = ctypes.c_void_p()
        st = self._L.blu_create(ctypes.byref(h), int(m), int(b_nz), int(device))
        if st != 0:
            raise RuntimeError(f"blu_create failed with status {st} (no CUDA device? there is no CPU path)")
        self._h = h
        self.m = int(m)
        self.lhs = np.zeros(self.m)
        self.ilhs = np.zeros(self.m, dtype=np.int64)
        self.nzlhs = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.blu_destroy(self._h)
            object.__setattr__(self, "_h", None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self, name):
        return self._L.blu_get_info(self._h, I[name])

    # blu.rs:95
    def factorize(self, b_begin, b_end, b_i, b_x):
        bb, be, bi, bx = _i64(b_begin), _i64(b_end), _i64(b_i), _f64(b_x)
        if len(bb) < self.m or len(be) < self.m:
            raise IndexError("b_begin/b_end shorter than m")
        return self._L.blu_factorize(self._h, _pi(bb), _pi(be), _pi(bi), _pf(bx))

    def factorize_c0ntinue(self, b_begin, b_end, b_i, b_x, c0ntinue):
        """factorize() of the crate's free-function surface (lib.rs:11-19, factorize.rs:34): Reallocate escapes
        (status 1, `addmem_l/u/w` say how much is missing); grow `l_mem`/`u_mem`/`w_mem` and call again with
        c0ntinue=True."""
        bb, be, bi, bx = _i64(b_begin), _i64(b_end), _i64(b_i), _f64(b_x)
        return self._L.blu_factorize_c0ntinue(self._h, _pi(bb), _pi(be), _pi(bi), _pf(bx), 1 if c0ntinue else 0)

    # blu.rs:139
    def get_factors(self, want_l=True, want_u=True):
        m = self.m
        lnz, unz = int(self.info("l_nz")), int(self.info("u_nz"))
        out = dict(rowperm=np.zeros(m, np.int64), colperm=np.zeros(m, np.int64))
        if want_l:
            out.update(l_colptr=np.zeros(m + 1, np.int64), l_rowidx=np.zeros(m + lnz, np.int64), l_value=np.zeros(m + lnz))
        if want_u:
            out.update(u_colptr=np.zeros(m + 1, np.int64), u_rowidx=np.zeros(m + unz, np.int64), u_value=np.zeros(m + unz))
        st = self._L.blu_get_factors(self._h, _pi(out["rowperm"]), _pi(out["colperm"]),
                                     _pi(out.get("l_colptr")), _pi(out.get("l_rowidx")), _pf(out.get("l_value")),
                                     _pi(out.get("u_colptr")), _pi(out.get("u_rowidx")), _pf(out.get("u_value")))
        return st, out

    # blu.rs:182
    def solve_dense(self, rhs, trans="N"):
        r = _f64(rhs)
        if len(r) != self.m:
            raise IndexError("rhs length != m")   # the reference panics here (lu/solve_dense.rs:37)
        x = np.zeros(self.m)
        st = self._L.blu_solve_dense(self._h, _pf(r), _pf(x), _ch(trans))
        return st, x

    def _clear_lhs(self):
        """lu_clear_lhs, blu.rs:380-395: zero what the previous solve scattered."""
        nz = self.nzlhs
        if nz:
            if nz <= int(self.get_param("sparse_thres") * self.m):
                self.lhs[self.ilhs[:nz]] = 0.0
            else:
                self.lhs[:] = 0.0
            self.nzlhs = 0

    def solve_dense_multi(self, rhs, trans="N"):
        """rhs[nrhs, m] -> (status, x[nrhs, m]): solve_dense for every row, all in flight together."""
        r = _f64(rhs).reshape(-1, self.m)
        x = np.zeros_like(r)
        st = self._L.blu_solve_dense_multi(self._h, r.shape[0], _pf(r), _pf(x), _ch(trans))
        return st, x

    def solve_sparse_multi(self, rhs_list, trans="N"):
        """rhs_list = [(irhs, xrhs), ...] -> (status, [(ilhs, values), ...]): solve_sparse for each, all in
        flight together; entry n of a result is lhs[ilhs[n]] = values[n], in solve_sparse's order."""
        n = len(rhs_list)
        begin = np.zeros(n + 1, dtype=np.int64)
        for r, (ii, _) in enumerate(rhs_list):
            begin[r + 1] = begin[r] + len(ii)
        irhs = _i64(np.concatenate([np.asarray(ii, dtype=np.int64) for ii, _ in rhs_list])) if n else np.zeros(0, np.int64)
        xrhs = _f64(np.concatenate([np.asarray(xx, dtype=np.float64) for _, xx in rhs_list])) if n else np.zeros(0)
        nz = np.zeros(n, dtype=np.int64)
        il = np.zeros(n * self.m, dtype=np.int64)
        xl = np.zeros(n * self.m)
        stat = np.zeros(n, dtype=np.int32)
        st = self._L.blu_solve_sparse_multi(self._h, n, _pi(begin), _pi(irhs), _pf(xrhs), _pi(nz), _pi(il), _pf(xl),
                                            stat.ctypes.data_as(i32p), _ch(trans))
        out = []
        for r in range(n):
            k = max(int(nz[r]), 0)
            out.append((il[r * self.m:r * self.m + k].copy(), xl[r * self.m:r * self.m + k].copy()))
        return st, out, stat

    # blu.rs:207: result in self.lhs / self.ilhs / self.nzlhs
    def solve_sparse(self, nzrhs, irhs, xrhs, trans="N"):
        ir, xr = _i64(irhs), _f64(xrhs)
        self._clear_lhs()
        nz = ctypes.c_int64(0)
        st = self._L.blu_solve_sparse(self._h, int(nzrhs), _pi(ir), _pf(xr), ctypes.byref(nz), _pi(self.ilhs), _pf(self.lhs), _ch(trans))
        self.nzlhs = nz.value
        return st

    # blu.rs:257
    def solve_for_update(self, nzrhs, irhs, xrhs, trans="N", want_solution=0):
        ir = _i64(irhs)
        xr = _f64(xrhs) if xrhs is not None else None
        self._clear_lhs()
        # D13 repaired: blu.rs:268-283 passes Some(lhs)
        # even when want_solution == 0; BASICLU semantics = no solution unless asked for.
        if want_solution:
            nz = ctypes.c_int64(0)
            st = self._L.blu_solve_for_update(self._h, int(nzrhs), _pi(ir), _pf(xr), ctypes.byref(nz), _pi(self.ilhs), _pf(self.lhs), _ch(trans))
            self.nzlhs = nz.value
        else:
            st = self._L.blu_solve_for_update(self._h, int(nzrhs), _pi(ir), _pf(xr), None, None, None, _ch(trans))
        return st

    # blu.rs:319
    def update(self, xtbl):
        return self._L.blu_update(self._h, float(xtbl))


class BLUBatch(_Base):
    """Many independent bases of one dimension on one GPU (one `BLU` per basis on the CPU)."""

    def __init__(self, nmat, m, bnz_cap, device=-1, lib=None):
        self._L = lib or load_library()
        h = ctypes.c_void_p()
        st = self._L.blu_batch_create(ctypes.byref(h), int(nmat), int(m), int(bnz_cap), int(device))
        if st != 0:
            raise RuntimeError(f"blu_batch_create failed with status {st} (no CUDA device? there is no CPU path)")
        self._h = h
        self.nmat, self.m = int(nmat), int(m)

    def close(self):
        if getattr(self, "_h", None):
            self._L.blu_batch_destroy(self._h)
            object.__setattr__(self, "_h", None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self, k, name):
        return self._L.blu_batch_get_info(self._h, int(k), I[name])

    def factorize(self, b_begin, b_end, b_i, b_x):
        bb, be, bi, bx = _i64(b_begin), _i64(b_end), _i64(b_i), _f64(b_x)
        status = np.zeros(self.nmat, dtype=np.int32)
        st = self._L.blu_batch_factorize(self._h, _pi(bb), _pi(be), _pi(bi), _pf(bx), len(bi), status.ctypes.data_as(i32p))
        return st, status

    def solve_dense(self, rhs, trans="N"):
        r = _f64(rhs).reshape(-1)
        x = np.zeros(self.nmat * self.m)
        status = np.zeros(self.nmat, dtype=np.int32)
        st = self._L.blu_batch_solve_dense(self._h, _pf(r), _pf(x), _ch(trans), status.ctypes.data_as(i32p))
        return st, x.reshape(self.nmat, self.m), status

    def solve_for_update(self, rhs_list, trans="N", want_solution=0):
        """rhs_list[k] = (irhs, xrhs or None) for basis k -> (status, per-basis status, [(ilhs, values)] or None)"""
        n = self.nmat
        begin = np.zeros(n + 1, dtype=np.int64)
        for k, (ii, _) in enumerate(rhs_list):
            begin[k + 1] = begin[k] + len(ii)
        irhs = _i64(np.concatenate([np.asarray(ii, dtype=np.int64).reshape(-1) for ii, _ in rhs_list]))
        have_x = rhs_list[0][1] is not None
        xrhs = _f64(np.concatenate([np.asarray(xx, dtype=np.float64) for _, xx in rhs_list])) if have_x else None
        stat = np.zeros(n, dtype=np.int32)
        if want_solution:
            nz = np.zeros(n, dtype=np.int64); il = np.zeros(n * self.m, dtype=np.int64); xl = np.zeros(n * self.m)
            st = self._L.blu_batch_solve_for_update(self._h, _pi(begin), _pi(irhs), _pf(xrhs), 1, _pi(nz), _pi(il), _pf(xl),
                                                    stat.ctypes.data_as(i32p), _ch(trans))
            out = [(il[k * self.m:k * self.m + max(int(nz[k]), 0)].copy(), xl[k * self.m:k * self.m + max(int(nz[k]), 0)].copy()) for k in range(n)]
            return st, stat, out
        st = self._L.blu_batch_solve_for_update(self._h, _pi(begin), _pi(irhs), _pf(xrhs), 0, None, None, None,
                                                stat.ctypes.data_as(i32p), _ch(trans))
        return st, stat, None

    def update(self, xtbl):
        x = _f64(xtbl)
        stat = np.zeros(self.nmat, dtype=np.int32)
        st = self._L.blu_batch_update(self._h, _pf(x), stat.ctypes.data_as(i32p))
        return st, stat

    def get_factors(self, k):
        m = self.m
        lnz, unz = int(self.info(k, "l_nz")), int(self.info(k, "u_nz"))
        out = dict(rowperm=np.zeros(m, np.int64), colperm=np.zeros(m, np.int64),
                   l_colptr=np.zeros(m + 1, np.int64), l_rowidx=np.zeros(m + lnz, np.int64), l_value=np.zeros(m + lnz),
                   u_colptr=np.zeros(m + 1, np.int64), u_rowidx=np.zeros(m + unz, np.int64), u_value=np.zeros(m + unz))
        st = self._L.blu_batch_get_factors(self._h, int(k), _pi(out["rowperm"]), _pi(out["colperm"]),
                                           _pi(out["l_colptr"]), _pi(out["l_rowidx"]), _pf(out["l_value"]),
                                           _pi(out["u_colptr"]), _pi(out["u_rowidx"]), _pf(out["u_value"]))
        return st, out

    # device-resident path (bench)
    def upload(self, b_begin=None, b_end=None, b_i=None, b_x=None, rhs=None):
        bb = _i64(b_begin) if b_begin is not None else None
        be = _i64(b_end) if b_end is not None else None
        bi = _i64(b_i) if b_i is not None else None
        bx = _f64(b_x) if b_x is not None else None
        r = _f64(rhs).reshape(-1) if rhs is not None else None
        return self._L.blu_batch_upload(self._h, _pi(bb), _pi(be), _pi(bi), _pf(bx), len(bi) if bi is not None else 0, _pf(r))

    def factorize_resident(self):
        return self._L.blu_batch_factorize_resident(self._h)

    def factorize_dev(self, d_b_begin, d_b_end, d_b_i, d_b_x, bnz_total):
        """B in caller-owned device memory (raw device pointers as ints, e.g. torch.Tensor.data_ptr())."""
        return self._L.blu_batch_factorize_dev(self._h, d_b_begin, d_b_end, d_b_i, d_b_x, int(bnz_total))

    def solve_dense_dev(self, d_rhs, d_lhs, trans="N", d_status=None):
        return self._L.blu_batch_solve_dense_dev(self._h, d_rhs, d_lhs, _ch(trans), d_status)

    def graph_capture(self, trans="N"):
        return self._L.blu_batch_graph_capture(self._h, _ch(trans))

    def graph_launch(self):
        return self._L.blu_batch_graph_launch(self._h)

    def solve_dense_resident(self, trans="N"):
        return self._L.blu_batch_solve_dense_resident(self._h, _ch(trans))

    def download(self):
        x = np.zeros(self.nmat * self.m)
        status = np.zeros(self.nmat, dtype=np.int32)
        st = self._L.blu_batch_download(self._h, _pf(x), status.ctypes.data_as(i32p))
        return st, x.reshape(self.nmat, self.m), status

    def last_kernel_ms(self, which):
        return self._L.blu_batch_last_kernel_ms(self._h, int(which))

    def launch_count(self):
        return self._L.blu_batch_launch_count(self._h)

    def set_stream(self, stream_ptr):
        return self._L.blu_batch_set_stream(self._h, ctypes.c_void_p(stream_ptr))

    def synchronize(self):
        return self._L.blu_batch_synchronize(self._h)


class BLUMulti:
    """One batch of independent bases spread over several GPUs of this process (blu_multi_*: contiguous
    ranges per device, one host thread and stream each, no collective -- SURVEY.md 8e)."""

    def __init__(self, nmat, m, bnz_cap, devices, lib=None):
        self._L = lib or load_library()
        self.nmat, self.m = int(nmat), int(m)
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        h = ctypes.c_void_p()
        st = self._L.blu_multi_create(ctypes.byref(h), self.nmat, self.m, int(bnz_cap), dev.ctypes.data_as(i32p), len(dev))
        if st != 0:
            raise RuntimeError(f"blu_multi_create failed with status {st} (no CUDA device? there is no CPU path)")
        self._h = h
        self.ndev = len(dev)

    def factorize_solve(self, b_begin, b_end, b_i, b_x, rhs=None, trans="N"):
        bb, be, bi, bx = _i64(b_begin), _i64(b_end), _i64(b_i), _f64(b_x)
        status = np.zeros(self.nmat, dtype=np.int32)
        r = _f64(rhs).reshape(-1) if rhs is not None else None
        x = np.zeros(self.nmat * self.m) if rhs is not None else None
        st = self._L.blu_multi_factorize_solve(self._h, _pi(bb), _pi(be), _pi(bi), _pf(bx), len(bi), _pf(r), _pf(x), _ch(trans), status.ctypes.data_as(i32p))
        return st, (x.reshape(self.nmat, self.m) if x is not None else None), status

    def part(self, d):
        """(batch handle usable with blu_batch_get_info / get_factors, first basis, number of bases) of device slot d"""
        first, count = ctypes.c_int64(), ctypes.c_int64()
        h = self._L.blu_multi_part(self._h, int(d), ctypes.byref(first), ctypes.byref(count))
        return h, first.value, count.value

    def info(self, k, name):
        for d in range(self.ndev):
            h, first, count = self.part(d)
            if first <= k < first + count:
                return self._L.blu_batch_get_info(h, int(k - first), I[name])
        raise IndexError(k)

    def close(self):
        if getattr(self, "_h", None):
            self._L.blu_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
