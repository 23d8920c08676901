/*
 * blo.h -- CPU ORACLE for the BLU sparse-LU hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the algorithm of rwl/blu (a Rust port of
 * BASICLU).  It exists so that the CUDA product in blu_b200/ can be checked
 * bit-for-bit (pivot sequence, permutations, rank, patterns) and to 1e-12
 * (values, solves).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call it.  The product never does.
 *
 * PARITY PINNING: the reference ships no tests and cannot be compiled in this
 * environment (no Rust toolchain), so this oracle is "parity unpinned" against
 * reference *executions*.  It is pinned against (a) the only known-answer
 * fixture of the reference, examples/simple.rs:21-33 (solution 0.1..1.0 and the
 * hand-traced first two pivots), (b) algebraic invariants
 * (B[rowperm,colperm] = L*U, residuals), (c) scipy SuperLU for solution values.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference/src).  Defects of the Rust port that are REPAIRED here
 * (SURVEY.md section 0): D1 eta_row is its own array; D2-D4 update() reach
 * vectors; D5 64-bit cancellation mask; D6 markowitz loop advance; D7 sentinel
 * refresh after reallocation; D9 get_factors before factorize -> INVALID_CALL;
 * D12 signed arithmetic in update() compression tests; D13 BLU::solve_for_update
 * computes no solution when want_solution == 0 (blu.rs:268-283 always passes Some(lhs)); D14 bfs_path's
 * `for front in 0..tail` (update.rs:68) fixes the range at entry although `tail` grows inside the loop, so the
 * reference's augmenting-path search never leaves j0 -- the oracle uses BASICLU's dynamic bound.  Every repair
 * sits behind a compile-time flag BLO_REPAIR_Dn (blo_int.h, default 1; 0 = the Rust source's behaviour, with
 * a named trap where that is a panic or an endless loop).  D8 (search_rows
 * default 0) is REPRODUCED.  D11 (release-mode file_diff asserts in
 * setup_bump) is behind the run-time flag `check_file_diff` (default on, as
 * in the reference).
 */
#ifndef BLO_H
#define BLO_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t lint; /* lib.rs:32  LUInt = i64 */

/* Return codes.  The Rust enum (lib.rs:39-64) has no discriminants; we use
 * BASICLU's numbering so that C callers can share it with the product ABI. */
enum {
    BLO_OK = 0,
    BLO_REALLOCATE = 1,
    BLO_WARNING_SINGULAR_MATRIX = 2,
    BLO_ERROR_INVALID_CALL = -2,
    BLO_ERROR_ARGUMENT_MISSING = -3,
    BLO_ERROR_INVALID_ARGUMENT = -4,
    BLO_ERROR_MAXIMUM_UPDATES = -5,
    BLO_ERROR_SINGULAR_UPDATE = -6
};

/* def.rs:6-12 */
enum { BLO_TASK_NONE = 0, BLO_TASK_SINGLETONS, BLO_TASK_SETUP_BUMP,
       BLO_TASK_FACTORIZE_BUMP, BLO_TASK_BUILD_FACTORS };

/* per-pivot trace record (oracle-only instrumentation) */
typedef struct {
    lint row, col;
    double pivot;
    int kind; /* 0 singleton-col phase, 1 singleton-row phase, 2 bump singleton row,
                 3 bump singleton col, 4 doubleton col, 5 small, 6 any, 7 rank-deficient drop */
    lint nz_row, nz_col;
} blo_trace;

/* struct LU, lu.rs:9-171 (aliasing of lu.rs:173-233 undone: every role has its own array) */
typedef struct blo_lu {
    /* memory sizes, lu.rs:10-15 */
    lint l_mem, u_mem, w_mem;
    /* parameters, lu.rs:17-66, defaults lu.rs:250-259 */
    double droptol, abstol, reltol;
    lint nzbias;      /* >= 0 <=> Some(_), < 0 <=> None */
    lint maxsearch, pad;
    double stretch, compress_thres, sparse_thres;
    lint search_rows;
    int check_file_diff; /* D11 */

    lint m;
    lint addmem_l, addmem_u, addmem_w;
    lint nupdate; /* -1 <=> None */
    lint nforrest, nfactorize, nupdate_total, nforrest_total, nsymperm_total;
    lint l_nz, u_nz, r_nz;
    double min_pivot, max_pivot, max_eta;
    double update_cost_numer, update_cost_denom;
    double time_factorize, time_solve, time_update;
    double time_factorize_total, time_solve_total, time_update_total;
    lint l_flops, u_flops, r_flops;
    double condest_l, condest_u, norm_l, norm_u, normest_l_inv, normest_u_inv;
    double onenorm, infnorm, residual_test;
    lint matrix_nz, rank, bump_size, bump_nz, nsearch_pivot, nexpand, ngarbage, factor_flops;
    double time_singletons, time_search_pivot, time_elim_pivot;
    double pivot_error;

    /* private, lu.rs:123-133 */
    int task;
    lint pivot_row, pivot_col;          /* -1 <=> None */
    lint ftran_for_update, btran_for_update; /* -1 <=> None */
    lint marker, pivotlen, rankdef, min_colnz, min_rownz;

    lint *l_index, *u_index, *w_index;
    double *l_value, *u_value, *w_value;

    lint *colcount_flink, *colcount_blink; /* 2m+2 */
    lint *rowcount_flink, *rowcount_blink; /* 2m+2 */
    lint *w_begin, *w_end, *w_flink, *w_blink; /* 2m+2 */
    lint *pinv, *qinv;   /* m; become pmap, qmap after build_factors */
    lint *l_begin_p, *u_begin; /* m+1 */
    lint *pivotcol, *pivotrow; /* 2m+2 */
    lint *l_begin, *lt_begin, *lt_begin_p, *p; /* m+1 */
    lint *r_begin, *eta_row; /* m+1 each (D1 repaired) */
    lint *iwork1;  /* 2m+2 */
    lint *iwork0;  /* m, == marked */
    double *work0, *work1, *col_pivot, *row_pivot; /* m */
    uint64_t *cancelled; /* m; D5: real 64-bit masks (BASICLU semantics) */
    lint *pstack;        /* m; dfs position stack (the reference reuses work1 as f64, dfs.rs:70) */

    /* accounting for the roofline numerators */
    double elim_bytes; /* algorithmic bytes of the elimination phase */
    lint nelim_div;    /* sum (nz_col-1) divides */

    /* oracle-only trace */
    blo_trace *trace; lint trace_cap, trace_len; int trace_on;
} blo_lu;

/* struct BLU, blu.rs:9-20 */
typedef struct blo {
    blo_lu lu;
    double *lhs; lint *ilhs; lint nzlhs;
    double realloc_factor;
} blo;

/* ---- L3 object API (blu.rs) ---- */
blo *blo_new(lint m, lint b_nz);                                   /* blu.rs:61 */
void blo_free(blo *o);
int blo_factorize(blo *o, const lint *b_begin, const lint *b_end,
                  const lint *b_i, const double *b_x);             /* blu.rs:95 */
int blo_solve_dense(blo *o, const double *rhs, double *lhs, char trans); /* blu.rs:182 */
int blo_solve_sparse(blo *o, lint nzrhs, const lint *irhs, const double *xrhs, char trans); /* blu.rs:207 */
int blo_solve_for_update(blo *o, lint nzrhs, const lint *irhs, const double *xrhs /*nullable*/,
                         char trans, lint want_solution);          /* blu.rs:257 */
int blo_update(blo *o, double xtbl);                               /* blu.rs:319 */
int blo_get_factors(blo *o, lint *rowperm, lint *colperm,
                    lint *l_colptr, lint *l_rowidx, double *l_value,
                    lint *u_colptr, lint *u_rowidx, double *u_value); /* blu.rs:139 */
int blo_maxvolume(blo *o, lint ncol, const lint *a_p, const lint *a_i, const double *a_x,
                  lint *basis, lint *isbasic, double volumetol, lint *p_nupdate); /* maxvolume.rs:64 */

/* ---- L2 free functions on the LU state (lib.rs:11-19) ---- */
void blo_lu_init(blo_lu *lu, lint m, lint b_nz);                   /* lu.rs:243 */
void blo_lu_release(blo_lu *lu);
void blo_lu_reset(blo_lu *lu);                                     /* lu.rs:329 */
int blo_lu_factorize(blo_lu *lu, const lint *b_begin, const lint *b_end,
                     const lint *b_i, const double *b_x, int c0ntinue); /* factorize.rs:34 */
int blo_lu_solve_dense(blo_lu *lu, const double *rhs, double *lhs, char trans); /* solve_dense.rs:24 */
int blo_lu_solve_sparse(blo_lu *lu, lint nzrhs, const lint *irhs, const double *xrhs,
                        lint *p_nzlhs, lint *ilhs, double *lhs, char trans); /* solve_sparse.rs:35 */
int blo_lu_solve_for_update(blo_lu *lu, lint nzrhs, const lint *irhs, const double *xrhs,
                            lint *p_nzlhs, lint *ilhs, double *lhs, char trans); /* solve_for_update.rs:72 */
int blo_lu_update(blo_lu *lu, double xtbl);                        /* update.rs:49 */
int blo_lu_get_factors(blo_lu *lu, lint *rowperm, lint *colperm,
                       lint *l_colptr, lint *l_rowidx, double *l_value,
                       lint *u_colptr, lint *u_rowidx, double *u_value); /* get_factors.rs:48 */
double blo_lu_update_cost(const blo_lu *lu);                       /* lu.rs:324 */

/* one instance per core over a batch of bases (blo_batch.c); returns the threads used */
int blo_batch_factorize_solve(lint nmat, lint m, const lint *b_begin, const lint *b_end,
                              const lint *b_i, const double *b_x, const double *rhs, double *lhs,
                              char trans, lint store_nz, int nthreads, int check_file_diff, int *status);

/* oracle-only helpers for the tests */
void blo_trace_enable(blo *o, int on);
lint blo_trace_len(const blo *o);
const blo_trace *blo_trace_data(const blo *o);
/* scalar getter by name index, so ctypes users need not mirror the struct */
double blo_get_info(const blo *o, int what);
void blo_set_param(blo *o, int what, double v);
enum { /* blo_get_info / blo_set_param selectors */
    BLO_P_DROPTOL = 0, BLO_P_ABSTOL, BLO_P_RELTOL, BLO_P_NZBIAS, BLO_P_MAXSEARCH, BLO_P_PAD,
    BLO_P_STRETCH, BLO_P_COMPRESS_THRES, BLO_P_SPARSE_THRES, BLO_P_SEARCH_ROWS,
    BLO_P_CHECK_FILE_DIFF, BLO_P_REALLOC_FACTOR,
    BLO_I_M = 100, BLO_I_RANK, BLO_I_BUMP_SIZE, BLO_I_BUMP_NZ, BLO_I_MATRIX_NZ, BLO_I_L_NZ,
    BLO_I_U_NZ, BLO_I_R_NZ, BLO_I_NSEARCH_PIVOT, BLO_I_NEXPAND, BLO_I_NGARBAGE,
    BLO_I_FACTOR_FLOPS, BLO_I_MIN_PIVOT, BLO_I_MAX_PIVOT, BLO_I_MAX_ETA, BLO_I_NUPDATE,
    BLO_I_NFORREST, BLO_I_NFACTORIZE, BLO_I_NUPDATE_TOTAL, BLO_I_NFORREST_TOTAL,
    BLO_I_NSYMPERM_TOTAL, BLO_I_L_FLOPS, BLO_I_U_FLOPS, BLO_I_R_FLOPS, BLO_I_CONDEST_L,
    BLO_I_CONDEST_U, BLO_I_NORM_L, BLO_I_NORM_U, BLO_I_NORMEST_L_INV, BLO_I_NORMEST_U_INV,
    BLO_I_ONENORM, BLO_I_INFNORM, BLO_I_RESIDUAL_TEST, BLO_I_PIVOT_ERROR, BLO_I_UPDATE_COST,
    BLO_I_TIME_FACTORIZE, BLO_I_TIME_SOLVE, BLO_I_TIME_UPDATE, BLO_I_TIME_SINGLETONS,
    BLO_I_TIME_SEARCH_PIVOT, BLO_I_TIME_ELIM_PIVOT, BLO_I_L_MEM, BLO_I_U_MEM, BLO_I_W_MEM,
    BLO_I_NZLHS, BLO_I_ELIM_BYTES, BLO_I_NELIM_DIV, BLO_I_PIVOTLEN, BLO_I_RANKDEF
};
const double *blo_lhs(const blo *o);
const lint *blo_ilhs(const blo *o);

#ifdef __cplusplus
}
#endif
#endif
