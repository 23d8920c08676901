/* blo_batch.c -- CPU oracle (TEST / BASELINE INFRASTRUCTURE ONLY): "one independent BLU
 * instance per core" driver for batches of bases.  Every worker thread
 * owns one blo object (struct BLU, blu.rs:9-20) and pulls basis indices from a shared
 * counter; per basis it does what examples/simple.rs:38-42 does: factorize + solve_dense.
 * The reference has no threads (SURVEY.md 8b "Threading"): independence of objects is what
 * makes this legal.  Used by bench.py (cpu_baseline leg, --impl reference) and tests. */
#include "blo_int.h"
#include <pthread.h>
#include <stdatomic.h>

typedef struct {
    lint nmat, m;
    const lint *b_begin, *b_end, *b_i;
    const double *b_x, *rhs;
    double *lhs;
    int *status;
    int check_file_diff;
    char trans;
    lint store_nz;
    atomic_long next;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *J = arg;
    blo *o = blo_new(J->m, J->store_nz);
    if (!o) return NULL;
    o->lu.check_file_diff = J->check_file_diff;
    double *x = malloc((size_t)J->m * sizeof(double));
    for (;;) {
        long k = atomic_fetch_add(&J->next, 1);
        if (k >= J->nmat) break;
        const lint *bb = J->b_begin + (size_t)k * J->m, *be = J->b_end + (size_t)k * J->m;
        int st = blo_factorize(o, bb, be, J->b_i, J->b_x);
        if ((st == BLO_OK || st == BLO_WARNING_SINGULAR_MATRIX) && J->rhs) {
            blo_solve_dense(o, J->rhs + (size_t)k * J->m, J->lhs ? J->lhs + (size_t)k * J->m : x, J->trans);
        }
        if (J->status) J->status[k] = st;
    }
    free(x);
    blo_free(o);
    return NULL;
}

/* Column j of basis k is b_i/b_x[b_begin[k*m+j] .. b_end[k*m+j]) (same layout as
 * blu_batch_factorize in include/blu_b200.h).  store_nz: initial size of each of the
 * L/U/W stores (BLU::new(m, b_nz), lu.rs:245-247); they grow on demand and stay grown,
 * as a long-lived per-core instance's would.  Returns the number of threads used. */
int blo_batch_factorize_solve(lint nmat, lint m, const lint *b_begin, const lint *b_end,
                              const lint *b_i, const double *b_x, const double *rhs, double *lhs,
                              char trans, lint store_nz, int nthreads, int check_file_diff, int *status) {
    if (nthreads < 1) nthreads = 1;
    if ((lint)nthreads > nmat) nthreads = (int)(nmat > 0 ? nmat : 1);
    batch_job J = { nmat, m, b_begin, b_end, b_i, b_x, rhs, lhs, status, check_file_diff, trans, store_nz, 0 };
    atomic_init(&J.next, 0);
    pthread_t *th = malloc((size_t)nthreads * sizeof *th);
    int started = 0;
    for (int t = 0; t < nthreads; t++) {
        if (pthread_create(&th[t], NULL, batch_worker, &J) != 0) break;
        started++;
    }
    if (started == 0) batch_worker(&J);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
    return started ? started : 1;
}
