/* blo_int.h -- internal declarations of the CPU oracle (test infrastructure). */
#ifndef BLO_INT_H
#define BLO_INT_H

#include "blo.h"
#include <stdio.h>
#include <stdlib.h>
#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* The reference keeps every assert! live in release builds (Rust does not
 * strip them), so the oracle keeps them live too: never compile with NDEBUG. */
#ifdef NDEBUG
#error "the oracle must be built with asserts enabled"
#endif

static inline double blo_now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static inline void blo_iswap(lint *x, lint i, lint j) { lint t = x[i]; x[i] = x[j]; x[j] = t; }   /* def.rs:20 */
static inline void blo_fswap(double *x, lint i, lint j) { double t = x[i]; x[i] = x[j]; x[j] = t; } /* def.rs:26 */

/* list.rs */
void blo_list_init(lint *flink, lint *blink, lint nelem, lint nlist, lint *min_list);
void blo_list_add(lint elem, lint list, lint *flink, lint *blink, lint nelem, lint *min_list);
void blo_list_remove(lint *flink, lint *blink, lint elem);
void blo_list_move(lint elem, lint list, lint *flink, lint *blink, lint nelem, lint *min_list);
void blo_list_swap(lint *flink, lint *blink, lint e1, lint e2);

/* file.rs */
void blo_file_empty(lint nlines, lint *begin, lint *end, lint *next, lint *prev, lint fmem);
void blo_file_reappend(lint line, lint nlines, lint *begin, lint *end, lint *next, lint *prev,
                       lint *index, double *value, lint extra_space);
lint blo_file_compress(lint nlines, lint *begin, lint *end, const lint *next,
                       lint *index, double *value, double stretch, lint pad);
lint blo_file_diff(lint nrow, const lint *begin_row, const lint *end_row,
                   const lint *begin_col, const lint *end_col,
                   const lint *index, const double *value);

/* factorization phases */
int blo_singletons(blo_lu *lu, const lint *b_begin, const lint *b_end, const lint *b_i, const double *b_x);
int blo_setup_bump(blo_lu *lu, const lint *b_begin, const lint *b_end, const lint *b_i, const double *b_x);
int blo_markowitz(blo_lu *lu);
int blo_pivot(blo_lu *lu);
int blo_factorize_bump(blo_lu *lu);
int blo_build_factors(blo_lu *lu);
double blo_condest(lint m, const lint *u_begin, const lint *u_i, const double *u_x,
                   const double *pivot, const lint *perm, int upper, double *work,
                   double *norm, double *norminv);
void blo_residual_test(blo_lu *lu, const lint *b_begin, const lint *b_end, const lint *b_i, const double *b_x);
void blo_matrix_norm(blo_lu *lu, const lint *b_begin, const lint *b_end, const lint *b_i, const double *b_x);

/* solves */
void blo_garbage_perm(blo_lu *lu);
lint blo_dfs(lint i, const lint *begin, const lint *end, const lint *index, lint top,
             lint *xi, lint *pstack, lint *marked, lint marker);
lint blo_solve_symbolic(lint m, const lint *begin, const lint *end, const lint *index,
                        lint nrhs, const lint *irhs, lint *ilhs, lint *pstack,
                        lint *marked, lint marker);
lint blo_solve_triangular(lint nz_symb, const lint *pattern_symb, const lint *begin, const lint *end,
                          const lint *index, const double *value, const double *pivot,
                          double droptol, double *lhs, lint *pattern, lint *flops);
void blo_k_solve_dense(blo_lu *lu, const double *rhs, double *lhs, char trans);
void blo_k_solve_sparse(blo_lu *lu, lint nrhs, const lint *irhs, const double *xrhs,
                        lint *p_nlhs, lint *ilhs, double *xlhs, char trans);
int blo_k_solve_for_update(blo_lu *lu, lint nrhs, const lint *irhs, const double *xrhs,
                           lint *p_nlhs, lint *ilhs, double *xlhs, char trans);
int blo_k_update(blo_lu *lu, double xtbl);

void blo_trace_push(blo_lu *lu, lint row, lint col, double pivot, int kind, lint nz_row, lint nz_col);


/* ---- repairs of the Rust port's defects ------------------------------------------
 * Every repair is behind a named compile-time flag, default ON (= BASICLU semantics, what the device path
 * implements).  Building with -DBLO_REPAIR_Dn=0 restores what the Rust source does at that place; where that
 * is a panic or an endless loop the oracle stops with a message naming the defect (BLO_DEFECT_TRAP) instead
 * of corrupting memory or hanging.  `make repairs-off-check` compiles every such variant. */
#ifndef BLO_REPAIR_D1   /* lu.rs:184-193: eta_row aliases r_begin */
#define BLO_REPAIR_D1 1
#endif
#ifndef BLO_REPAIR_D2   /* update.rs:422-423, 877-878: vec![0; ipivot] instead of [ipivot] */
#define BLO_REPAIR_D2 1
#endif
#ifndef BLO_REPAIR_D3   /* update.rs:634-643: reach vectors one element short */
#define BLO_REPAIR_D3 1
#endif
#ifndef BLO_REPAIR_D4   /* update.rs:797 vs 176-192: permute() handed nswap entries, reads nswap+1 */
#define BLO_REPAIR_D4 1
#endif
#ifndef BLO_REPAIR_D5   /* pivot.rs:645-664, 750: 32-bit cancellation mask round-tripped through f64 */
#define BLO_REPAIR_D5 1
#endif
#ifndef BLO_REPAIR_D6   /* markowitz.rs:90-92: `continue` without advancing j */
#define BLO_REPAIR_D6 1
#endif
#ifndef BLO_REPAIR_D7   /* blu.rs:345-377, lu.rs:308-314: sentinel / addmem not refreshed after reallocation */
#define BLO_REPAIR_D7 1
#endif
#ifndef BLO_REPAIR_D9   /* get_factors.rs:59: unwrap() on None instead of ErrorInvalidCall */
#define BLO_REPAIR_D9 1
#endif
#ifndef BLO_REPAIR_D12  /* update.rs:917, 925: usize subtraction wraps */
#define BLO_REPAIR_D12 1
#endif
#ifndef BLO_REPAIR_D13  /* blu.rs:268-283: solve_for_update scatters a solution nobody asked for */
#define BLO_REPAIR_D13 1
#endif
#ifndef BLO_REPAIR_D14  /* update.rs:68: `for front in 0..tail` fixes the range at entry, the BFS never leaves j0 */
#define BLO_REPAIR_D14 1
#endif
#define BLO_DEFECT_TRAP(name, what) do { fprintf(stderr, "blo oracle: reference defect %s reproduced: %s\n", name, what); abort(); } while (0)

#endif
