/* blu_b200.h -- C ABI of the B200-native BLU hot path (libblu_b200.so).
 *
 * Drop-in boundary for the public surface of the rwl/blu crate (paths relative to
 * /root/reference/src).  Every entry point takes plain host pointers and sizes; the
 * library owns all device memory, does the H2D/D2H copies and launches hand-written
 * sm_100a kernels.  There is no CPU fallback: without a CUDA device every call
 * returns BLU_ERROR_CUDA.
 *
 * Index type: the reference uses usize/LUInt = 64-bit (lib.rs:32); so does this ABI.
 * Return codes: the Rust `Status` enum (lib.rs:39-64) has no discriminants; BASICLU's
 * numbering is used.  Ok(()) == BLU_OK.  WarningSingularMatrix is returned like an
 * error by the crate but the factorization IS valid (factorize.rs:115-119,176-178).
 * `Reallocate` never escapes the object API (blu.rs:95-118 loops on it); it does not
 * escape this ABI either -- the library grows device memory and re-runs.
 */
#ifndef BLU_B200_H
#define BLU_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    BLU_OK = 0,
    BLU_REALLOCATE = 1,                 /* lib.rs:41  (internal, never returned) */
    BLU_WARNING_SINGULAR_MATRIX = 2,    /* lib.rs:44 */
    BLU_ERROR_INVALID_CALL = -2,        /* lib.rs:48 */
    BLU_ERROR_ARGUMENT_MISSING = -3,    /* lib.rs:51 */
    BLU_ERROR_INVALID_ARGUMENT = -4,    /* lib.rs:54 */
    BLU_ERROR_MAXIMUM_UPDATES = -5,     /* lib.rs:58 */
    BLU_ERROR_SINGULAR_UPDATE = -6,     /* lib.rs:63 */
    BLU_ERROR_INTERNAL = -100,          /* a device-side invariant (a reference assert!) failed */
    BLU_ERROR_CUDA = -101,              /* no device / CUDA runtime error */
    BLU_ERROR_OUT_OF_MEMORY = -102
};

/* parameter / info selectors for blu_set_param, blu_get_param, blu_get_info.
 * Parameters: pub fields of `LU` (lu/lu.rs:10-66) and BLU.realloc_factor (blu.rs:19).
 * Info: the getters of lu/lu.rs:399-683. */
enum {
    BLU_P_DROPTOL = 0, BLU_P_ABSTOL, BLU_P_RELTOL, BLU_P_NZBIAS, BLU_P_MAXSEARCH, BLU_P_PAD,
    BLU_P_STRETCH, BLU_P_COMPRESS_THRES, BLU_P_SPARSE_THRES, BLU_P_SEARCH_ROWS,
    BLU_P_REALLOC_FACTOR, BLU_P_L_MEM, BLU_P_U_MEM, BLU_P_W_MEM,
    BLU_P_THREADS_PER_BASIS,            /* CTA size of the factorization kernel (32..1024) */
    BLU_P_NORMS,                        /* 1 (default): factorize also runs condest x2 + residual_test as factorize.rs:121-147 does; 0: skip them (their getters then read 0) */
    BLU_P_DENSE_K,                      /* order at which the active submatrix switches to the dense-tail representation (multiple of 32, <= 256; 0 = never; default: the largest order whose values fit in shared memory, 160 on B200).  Results do not depend on it. */
    BLU_P_TAIL_THREADS,                 /* CTA size of the dense-tail launch of a split batch factorization (default 512) */
    BLU_P_SPLIT_MIN,                    /* batches of more bases than this run as three launches: sparse head, dense tail with one CTA per SM, build_factors (default 0: always, when the dense tail is shared-memory resident -- a single launch sized for 200+ KB of shared memory would leave the sparse head without L1) */
    BLU_P_TREE_MIN,                     /* bumps with more active columns than this find their Markowitz candidates through a min-tree over the column keys instead of a scan (default 4096; needs maxsearch <= 4) */
    BLU_P_DENSE_K_BIG,                  /* two-stage dense tail: the active submatrix turns dense already at this order (multiple of 32, <= 256, > BLU_P_DENSE_K) with its values in HBM/L2, and moves into shared memory when it has shrunk to BLU_P_DENSE_K (0 = one stage; default 256 for batches when BLU_P_DENSE_K is shared-memory resident).  Results do not depend on it. */
    BLU_I_M = 100, BLU_I_RANK, BLU_I_BUMP_SIZE, BLU_I_BUMP_NZ, BLU_I_MATRIX_NZ, BLU_I_L_NZ,
    BLU_I_U_NZ, BLU_I_R_NZ, BLU_I_NSEARCH_PIVOT, BLU_I_NEXPAND, BLU_I_NGARBAGE,
    BLU_I_FACTOR_FLOPS, BLU_I_MIN_PIVOT, BLU_I_MAX_PIVOT, BLU_I_MAX_ETA, BLU_I_NUPDATE,
    BLU_I_NFORREST, BLU_I_NFACTORIZE, BLU_I_NUPDATE_TOTAL, BLU_I_NFORREST_TOTAL,
    BLU_I_NSYMPERM_TOTAL, BLU_I_L_FLOPS, BLU_I_U_FLOPS, BLU_I_R_FLOPS, BLU_I_CONDEST_L,
    BLU_I_CONDEST_U, BLU_I_NORM_L, BLU_I_NORM_U, BLU_I_NORMEST_L_INV, BLU_I_NORMEST_U_INV,
    BLU_I_ONENORM, BLU_I_INFNORM, BLU_I_RESIDUAL_TEST, BLU_I_PIVOT_ERROR, BLU_I_UPDATE_COST,
    BLU_I_TIME_FACTORIZE, BLU_I_TIME_SOLVE, BLU_I_TIME_UPDATE, BLU_I_ELIM_BYTES, BLU_I_NELIM_DIV,
    BLU_I_PIVOTLEN, BLU_I_RANKDEF, BLU_I_INTERNAL_ERROR, BLU_I_STATUS, BLU_I_NREALLOC,
    BLU_I_ELIM_BYTES_HEAD,              /* part of BLU_I_ELIM_BYTES done by the head launch of a split batch factorization */
    BLU_I_NRUNS,                        /* factorization passes started on this basis since creation (a Reallocate re-run counts) */
    BLU_I_ADDMEM_L, BLU_I_ADDMEM_U, BLU_I_ADDMEM_W, /* lu.rs:308-314: entries missing in L / U / W when the last call answered Reallocate */
    BLU_I_T_PHASE0 = 200, /* +0..15: SM cycles per phase of the factorization kernel (diagnostic) */
    BLU_I_N_KIND0 = 220,  /* +0..4: pivots taken by singleton-row / singleton-col / doubleton / small / any; +5: of those, steps taken in the dense tail; +6: entries into it */
    BLU_I_NORMS_CYC0 = 230 /* +0..3: SM cycles of condest(L), condest(U), residual forward, residual transposed (diagnostic) */
};

/* ------------------------------------------------------------------ */
/* object API: struct BLU, blu.rs:9-334                                 */
/* ------------------------------------------------------------------ */
typedef struct blu_b200 blu_t;

/* BLU::new(m, b_nz), blu.rs:61-70.  device < 0: current device.
 * 1 <= m < 2^23 (BLU_ERROR_INVALID_ARGUMENT otherwise); each of the L, U and W stores holds fewer than 2^30
 * (W: 2^29) entries per basis, beyond which growing them reports BLU_ERROR_OUT_OF_MEMORY. */
int blu_create(blu_t **out, int64_t m, int64_t b_nz, int device);
void blu_destroy(blu_t *o);

int blu_set_param(blu_t *o, int what, double value);
double blu_get_param(const blu_t *o, int what);
double blu_get_info(blu_t *o, int what);

/* BLU::factorize, blu.rs:95-118 -> factorize(), factorize.rs:34.  Column j of B is
 * b_i/b_x[b_begin[j] .. b_end[j]); pass (colptr, colptr+1) for CSC. */
int blu_factorize(blu_t *o, const int64_t *b_begin, const int64_t *b_end,
                  const int64_t *b_i, const double *b_x);

/* factorize() of the crate's FREE-FUNCTION surface (lib.rs:11-19, factorize.rs:34-119), where Reallocate escapes:
 * returns BLU_REALLOCATE with BLU_I_ADDMEM_L/U/W set; the caller grows the stores (blu_set_param with
 * BLU_P_L_MEM / BLU_P_U_MEM / BLU_P_W_MEM -- lu_realloc_obj, blu.rs:345-377) and calls again with c0ntinue != 0.
 * c0ntinue != 0 without a pending Reallocate: BLU_ERROR_INVALID_CALL (factorize.rs:102-105). */
int blu_factorize_c0ntinue(blu_t *o, const int64_t *b_begin, const int64_t *b_end,
                           const int64_t *b_i, const double *b_x, int c0ntinue);

/* BLU::get_factors, blu.rs:139-160 -> get_factors.rs:48.  Every output may be NULL. */
int blu_get_factors(blu_t *o, int64_t *rowperm, int64_t *colperm,
                    int64_t *l_colptr, int64_t *l_rowidx, double *l_value,
                    int64_t *u_colptr, int64_t *u_rowidx, double *u_value);

/* BLU::solve_dense, blu.rs:182-184 -> solve_dense.rs:24.  trans 't'/'T' = transposed. */
int blu_solve_dense(blu_t *o, const double *rhs, double *lhs, char trans);

/* Many right-hand sides against one factorization: rhs and lhs hold nrhs vectors of
 * length m back to back; the result is bit-identical to nrhs calls of blu_solve_dense. */
int blu_solve_dense_multi(blu_t *o, int64_t nrhs, const double *rhs, double *lhs, char trans);

/* Many sparse right-hand sides against one factorization: right-hand side r is
 * irhs/xrhs[rhs_begin[r] .. rhs_begin[r+1]).  For every r: nzlhs[r] (or -1 and status[r], which may be NULL),
 * the pattern ilhs[r*m .. r*m+nzlhs[r]) in the order blu_solve_sparse returns it, and the VALUES of those
 * entries compacted in xlhs[r*m + n].  Bit-identical to nrhs calls of blu_solve_sparse; returns the first
 * non-OK status. */
int blu_solve_sparse_multi(blu_t *o, int64_t nrhs, const int64_t *rhs_begin, const int64_t *irhs, const double *xrhs,
                           int64_t *nzlhs, int64_t *ilhs, double *xlhs, int *status, char trans);

/* BLU::solve_sparse, blu.rs:207-230 -> solve_sparse.rs:35.  The result is returned like
 * BLU.lhs / ilhs / nzlhs: lhs[m] scattered (zero elsewhere), ilhs[0..*nzlhs) its pattern. */
int blu_solve_sparse(blu_t *o, int64_t nzrhs, const int64_t *irhs, const double *xrhs,
                     int64_t *nzlhs, int64_t *ilhs, double *lhs, char trans);

/* BLU::solve_for_update, blu.rs:257-294 -> solve_for_update.rs:72.  xrhs may be NULL for
 * trans; nzlhs/ilhs/lhs may all be NULL when the solution is not wanted. */
int blu_solve_for_update(blu_t *o, int64_t nzrhs, const int64_t *irhs, const double *xrhs,
                         int64_t *nzlhs, int64_t *ilhs, double *lhs, char trans);

/* BLU::update, blu.rs:319-334 -> update.rs:49 */
int blu_update(blu_t *o, double xtbl);

/* ------------------------------------------------------------------ */
/* batch API: many independent bases of one dimension on one GPU       */
/* (what "one BLU instance per core" is on the CPU)      */
/* ------------------------------------------------------------------ */
typedef struct blu_b200 blu_batch_t;

/* nmat objects of dimension m whose B has at most bnz_cap entries each. */
int blu_batch_create(blu_batch_t **out, int64_t nmat, int64_t m, int64_t bnz_cap, int device);
void blu_batch_destroy(blu_batch_t *b);

/* Column j of basis k is b_i/b_x[b_begin[k*m+j] .. b_end[k*m+j]); b_i/b_x hold `bnz_total`
 * entries in all.  status[k] (may be NULL) receives the per-basis code; the return value
 * is BLU_OK unless a call-level error occurred. */
int blu_batch_factorize(blu_batch_t *b, const int64_t *b_begin, const int64_t *b_end,
                        const int64_t *b_i, const double *b_x, int64_t bnz_total, int *status);
/* rhs, lhs: nmat*m, basis k at offset k*m */
int blu_batch_solve_dense(blu_batch_t *b, const double *rhs, double *lhs, char trans, int *status);
double blu_batch_get_info(blu_batch_t *b, int64_t k, int what);
int blu_batch_get_factors(blu_batch_t *b, int64_t k, int64_t *rowperm, int64_t *colperm,
                          int64_t *l_colptr, int64_t *l_rowidx, double *l_value,
                          int64_t *u_colptr, int64_t *u_rowidx, double *u_value);

/* One basis change on every basis of the batch at once (many LPs advancing together): solve_for_update
 * (blu.rs:257) and update (blu.rs:319) with one warp per basis.  Right-hand side of basis k:
 * irhs/xrhs[rhs_begin[k] .. rhs_begin[k+1]); for trans 't'/'T' one index (the leaving column), xrhs may be
 * NULL.  With want_solution: nzlhs[k], the pattern ilhs[k*m ..] and the values of those entries compacted
 * in xlhs[k*m + n].  status[k] carries the per-basis code (may be NULL); the return value is the first
 * non-OK one.  Reallocate never escapes (blu.rs:268-291, 319-334): the stores of the whole batch are grown,
 * content kept, and the bases that asked run again. */
int blu_batch_solve_for_update(blu_batch_t *b, const int64_t *rhs_begin, const int64_t *irhs, const double *xrhs,
                               int want_solution, int64_t *nzlhs, int64_t *ilhs, double *xlhs, int *status, char trans);
int blu_batch_update(blu_batch_t *b, const double *xtbl, int *status);

/* Device-resident variants for throughput measurement: upload once, then time kernels only. */
int blu_batch_upload(blu_batch_t *b, const int64_t *b_begin, const int64_t *b_end,
                     const int64_t *b_i, const double *b_x, int64_t bnz_total, const double *rhs);
int blu_batch_factorize_resident(blu_batch_t *b);
int blu_batch_solve_dense_resident(blu_batch_t *b, char trans);
int blu_batch_download(blu_batch_t *b, double *lhs, int *status);
/* SURVEY.md 8(f) N4.  Caller-owned DEVICE buffers (same layout as the host variants): no staging, no copies;
 * B must stay valid and unchanged until the next factorization of the batch.  d_status (nmat ints on the device)
 * may be NULL. */
int blu_batch_factorize_dev(blu_batch_t *b, const int64_t *d_b_begin, const int64_t *d_b_end,
                            const int64_t *d_b_i, const double *d_b_x, int64_t bnz_total);
int blu_batch_solve_dense_dev(blu_batch_t *b, const double *d_rhs, double *d_lhs, char trans, int *d_status);
/* The steady-state step of a resident batch (factorization launches + condest/residual_test + solve_dense of the
 * uploaded rhs) captured into a CUDA graph once, then replayed with one launch per step; a basis that answers
 * Reallocate during a replay is grown and re-run by the classic path before the call returns. */
int blu_batch_graph_capture(blu_batch_t *b, char trans);
int blu_batch_graph_launch(blu_batch_t *b);
/* cudaStream_t the library launches on (so callers can bracket it with their own events),
 * and a way to make it use the caller's stream instead */
void *blu_batch_stream(blu_batch_t *b);
int blu_batch_set_stream(blu_batch_t *b, void *cuda_stream);
int blu_batch_synchronize(blu_batch_t *b);
/* device time of the last factorize / solve kernels, measured with CUDA events on the launching stream */
double blu_batch_last_kernel_ms(blu_batch_t *b, int which /*0 factorize (all kernels of the call), 1 solve_dense, 2 the condest/residual_test kernel alone, 3 / 4 / 5 the head / tail / build launch of the last split factorization pass (3 = the single launch when not split)*/);
/* number of kernels this library launched since creation */
int64_t blu_batch_launch_count(blu_batch_t *b);

/* ------------------------------------------------------------------ */
/* one batch over several GPUs of this process ("a device list")  */
/* Bases are split into contiguous ranges, one per device (sizes differ by at most   */
/* one); every device has its own blu_batch_t, host thread and stream; there is no   */
/* collective -- the results land in the caller's arrays at the bases' own offsets.  */
/* ------------------------------------------------------------------ */
typedef struct blu_multi blu_multi_t;
int blu_multi_create(blu_multi_t **out, int64_t nmat, int64_t m, int64_t bnz_cap, const int *devices, int ndev);
void blu_multi_destroy(blu_multi_t *mb);
/* factorize every basis and (if rhs and lhs are given) solve with it; layout as blu_batch_factorize /
 * blu_batch_solve_dense.  status[k]: per-basis code of the factorization, or of the solve if that failed. */
int blu_multi_factorize_solve(blu_multi_t *mb, const int64_t *b_begin, const int64_t *b_end, const int64_t *b_i,
                              const double *b_x, int64_t bnz_total, const double *rhs, double *lhs, char trans, int *status);
/* the batch object of device slot d and the range of bases it owns (for the getters / further calls) */
blu_batch_t *blu_multi_part(blu_multi_t *mb, int d, int64_t *first, int64_t *count);

const char *blu_version(void);

#ifdef __cplusplus
}
#endif
#endif
