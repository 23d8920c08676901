/* abi_check.c -- a plain C program compiled against include/blu_b200.h only: proves that the header is valid C,
 * that every declared entry point links against libblu_b200.so, and that calls fail loudly (BLU_ERROR_CUDA,
 * never a CPU fallback) on a machine without a GPU.  With a GPU it factorizes the 10 x 10 example of
 * /root/reference/examples/simple.rs:21-33 and checks the solution 0.1 .. 1.0.
 * Build: gcc -std=c99 -I include tests/abi/abi_check.c -L blu_b200 -lblu_b200 -Wl,-rpath,blu_b200 */
#include <stdio.h>
#include <math.h>
#include "blu_b200.h"

int main(void) {
    /* examples/simple.rs:21-33 (CSC) */
    static const int64_t bp[11] = {0, 3, 6, 8, 13, 15, 16, 19, 23, 27, 32};
    static const int64_t bi[32] = {0, 7, 8, 1, 4, 9, 2, 9, 3, 6, 7, 8, 9, 1, 4, 5, 3, 6, 9, 0, 3, 7, 8, 0, 3, 7, 8, 1, 2, 3, 6, 9};
    static const double bx[32] = {2.1, 0.14, 0.09, 1.1, 0.06, 0.03, 1.7, 0.04, 1.0, 0.32, 0.19, 0.32, 0.44, 0.06, 1.6, 2.2, 0.32, 1.9, 0.43,
                                  0.14, 0.19, 1.1, 0.22, 0.09, 0.32, 0.22, 2.4, 0.03, 0.04, 0.44, 0.43, 3.2};
    double rhs[10], lhs[10];
    blu_t *o = NULL;
    int k, st;
    printf("%s\n", blu_version());
    st = blu_create(&o, 10, 32, -1);
    if (st == BLU_ERROR_CUDA) { printf("no CUDA device: blu_create -> BLU_ERROR_CUDA (no CPU fallback)\n"); return 0; }
    if (st != BLU_OK) { printf("blu_create failed: %d\n", st); return 1; }
    st = blu_factorize(o, bp, bp + 1, bi, bx);
    if (st != BLU_OK) { printf("blu_factorize: %d\n", st); return 1; }
    /* rhs = B * (0.1 .. 1.0) */
    for (k = 0; k < 10; k++) rhs[k] = 0.0;
    for (k = 0; k < 10; k++) { int64_t p; for (p = bp[k]; p < bp[k + 1]; p++) rhs[bi[p]] += bx[p] * 0.1 * (double)(k + 1); }
    st = blu_solve_dense(o, rhs, lhs, 'N');
    if (st != BLU_OK) { printf("blu_solve_dense: %d\n", st); return 1; }
    for (k = 0; k < 10; k++) if (fabs(lhs[k] - 0.1 * (double)(k + 1)) > 1e-12) { printf("x[%d] = %.17g\n", k, lhs[k]); return 1; }
    printf("rank %d, solution 0.1..1.0 recovered, residual_test %.3g\n", (int)blu_get_info(o, BLU_I_RANK), blu_get_info(o, BLU_I_RESIDUAL_TEST));
    blu_destroy(o);
    return 0;
}
