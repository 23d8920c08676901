/* blu_solve.cuh -- dense solves and factor export on the device.
 * Reference: src/lu/solve_dense.rs:7-120, src/lu/garbage_perm.rs:16-48, src/get_factors.rs:48-180. */
#ifndef BLU_SOLVE_CUH
#define BLU_SOLVE_CUH
#include "blu_dev_common.cuh"

__device__ __forceinline__ bool is_trans(char t) { return t == 't' || t == 'T'; }

/* garbage_perm.rs:16-48; one thread (only after > m updates worth of pivots accumulated) */
__device__ __forceinline__ void dev_garbage_perm(Mat &M) {
    BluInfo *I = M.info;
    const int m = M.m, pivotlen = I->pivotlen;
    if (pivotlen > m) {
        int marker = ++I->marker;
        int put = pivotlen;
        for (int get = pivotlen - 1; get >= 0; get--) {
            int j = M.pivotcol[get];
            if (M.marked[j] != marker) {
                M.marked[j] = marker;
                --put;
                M.pivotcol[put] = j;
                M.pivotrow[put] = M.pivotrow[get];
            }
        }
        for (int k = 0; k < m; k++) { M.pivotcol[k] = M.pivotcol[put + k]; M.pivotrow[k] = M.pivotrow[put + k]; }
        I->pivotlen = m;
    }
}

/* acc += term of lanes 0..n-1 in lane order, one rounding per step: exactly the sum the
 * reference's sequential dot-product loop produces (lu/solve_dense.rs:64-66,83-85). */
__device__ __forceinline__ double ordered_acc(double acc, double term, int n) {
    for (int t = 0; t < n; t++) acc = __dadd_rn(acc, __shfl_sync(FULLMASK, term, t));
    return acc;
}

/* One warp per basis.  The sweeps are sequential over the pivot order (as in the
 * reference); the lanes share the dot product / axpy of each step and the pointer
 * loads are batched 32 pivots at a time.  Every multiply, add, subtract and divide is
 * rounded once and sums run in the reference's order, so the result is bit-identical. */
__global__ void __launch_bounds__(32) k_solve_dense(BluDev D, const double *rhs_all, double *lhs_all, char trans, int *status) {
    __shared__ Mat M;
    const int lane = threadIdx.x & 31;
    for (int s = blockIdx.x; s < D.nmat; s += gridDim.x) {
        if (lane == 0) mat_view(M, D, s);
        __syncwarp();
        const int m = M.m;
        BluInfo *I = M.info;
        if (I->nupdate < 0) { if (lane == 0 && status) status[s] = BLU_ERROR_INVALID_CALL; __syncwarp(); continue; }
        if (lane == 0) dev_garbage_perm(M);
        __syncwarp();
        const double *rhs = rhs_all + (size_t)s * m;
        double *lhs = lhs_all + (size_t)s * m;
        double *work = M.work1;
        const int nforrest = I->nforrest;
        for (int i = lane; i < m; i += 32) work[i] = rhs[i];
        __syncwarp();
        if (is_trans(trans)) {
            /* U', lu/solve_dense.rs:40-48 */
            for (int kb = 0; kb < m; kb += 32) {
                int k = kb + lane;
                int jp = k < m ? M.pivotcol[k] : 0, ip = k < m ? M.pivotrow[k] : 0;
                int b = k < m ? M.lbeg[jp] : 0, e = k < m ? M.lend[jp] : 0;
                double piv = k < m ? M.colpiv[jp] : 1.0;
                int n = m - kb < 32 ? m - kb : 32;
                for (int t = 0; t < n; t++) {
                    int jj = __shfl_sync(FULLMASK, jp, t), ii = __shfl_sync(FULLMASK, ip, t);
                    int bb = __shfl_sync(FULLMASK, b, t), ee = __shfl_sync(FULLMASK, e, t);
                    double pv = __shfl_sync(FULLMASK, piv, t);
                    double x = __ddiv_rn(work[jj], pv);
                    __syncwarp();
                    for (int pos = bb + lane; pos < ee; pos += 32) { const int r = M.w_idx[pos]; work[r] = __dsub_rn(work[r], __dmul_rn(x, M.w_val[pos])); }
                    if (lane == 0) lhs[ii] = x;
                    __syncwarp();
                }
            }
            /* etas backwards, :52-59 */
            for (int t = nforrest - 1; t >= 0; t--) {
                double x = lhs[M.eta_row[t]];
                __syncwarp();
                for (int pos = M.r_begin[t] + lane; pos < M.r_begin[t + 1]; pos += 32) { const int r = M.l_idx[pos]; lhs[r] = __dsub_rn(lhs[r], __dmul_rn(x, M.l_val[pos])); }
                __syncwarp();
            }
            /* L', :63-73 */
            for (int kb = ((m - 1) / 32) * 32; kb >= 0; kb -= 32) {
                int k = kb + lane;
                int b = k < m ? M.l_begin_p[k] : 0, e = k < m ? M.l_begin_p[k + 1] - 1 : 0;
                int ip = k < m ? M.p[k] : 0;
                unsigned ne = __ballot_sync(FULLMASK, e > b);
                while (ne) {
                    int t = 31 - __clz((int)ne);
                    ne &= ~(1u << t);
                    int bb = __shfl_sync(FULLMASK, b, t), ee = __shfl_sync(FULLMASK, e, t), ii = __shfl_sync(FULLMASK, ip, t);
                    double x = 0.0;
                    for (int cb = bb; cb < ee; cb += 32) {
                        const int pos = cb + lane;
                        const double term = pos < ee ? __dmul_rn(lhs[M.l_idx[pos]], M.l_val[pos]) : 0.0;
                        x = ordered_acc(x, term, ee - cb < 32 ? ee - cb : 32);
                    }
                    if (lane == 0) lhs[ii] = __dsub_rn(lhs[ii], x);
                    __syncwarp();
                }
            }
        } else {
            /* L, :81-90 (row-wise dot form) */
            for (int kb = 0; kb < m; kb += 32) {
                int k = kb + lane;
                int b = k < m ? M.lt_begin_p[k] : 0, e = k < m ? M.lt_begin_p[k + 1] - 1 : 0;
                int ip = k < m ? M.p[k] : 0;
                unsigned ne = __ballot_sync(FULLMASK, e > b);
                while (ne) {
                    int t = __ffs((int)ne) - 1;
                    ne &= ne - 1;
                    int bb = __shfl_sync(FULLMASK, b, t), ee = __shfl_sync(FULLMASK, e, t), ii = __shfl_sync(FULLMASK, ip, t);
                    double x = 0.0;
                    for (int cb = bb; cb < ee; cb += 32) {
                        const int pos = cb + lane;
                        const double term = pos < ee ? __dmul_rn(work[M.l_idx[pos]], M.l_val[pos]) : 0.0;
                        x = ordered_acc(x, term, ee - cb < 32 ? ee - cb : 32);
                    }
                    if (lane == 0) work[ii] = __dsub_rn(work[ii], x);
                    __syncwarp();
                }
            }
            /* etas, :93-102 */
            for (int t = 0; t < nforrest; t++) {
                double x = 0.0;
                const int rb = M.r_begin[t], re = M.r_begin[t + 1];
                for (int cb = rb; cb < re; cb += 32) {
                    const int pos = cb + lane;
                    const double term = pos < re ? __dmul_rn(work[M.l_idx[pos]], M.l_val[pos]) : 0.0;
                    x = ordered_acc(x, term, re - cb < 32 ? re - cb : 32);
                }
                if (lane == 0) work[M.eta_row[t]] = __dsub_rn(work[M.eta_row[t]], x);
                __syncwarp();
            }
            /* U, :106-118 (column-wise axpy form, terminator-delimited) */
            for (int kb = ((m - 1) / 32) * 32; kb >= 0; kb -= 32) {
                int k = kb + lane;
                int jp = k < m ? M.pivotcol[k] : 0, ip = k < m ? M.pivotrow[k] : 0;
                int b = k < m ? M.u_begin[ip] : 0;
                double piv = k < m ? M.rowpiv[ip] : 1.0;
                int n = m - kb < 32 ? m - kb : 32;
                for (int t = n - 1; t >= 0; t--) {
                    int jj = __shfl_sync(FULLMASK, jp, t), ii = __shfl_sync(FULLMASK, ip, t);
                    int bb = __shfl_sync(FULLMASK, b, t);
                    double pv = __shfl_sync(FULLMASK, piv, t);
                    double x = __ddiv_rn(work[ii], pv);
                    __syncwarp();
                    for (int pos = bb;; pos += 32) {
                        int idx = M.u_idx[pos + lane];
                        unsigned term = __ballot_sync(FULLMASK, idx < 0);
                        int nvalid = term ? __ffs((int)term) - 1 : 32;
                        if (lane < nvalid) work[idx] = __dsub_rn(work[idx], __dmul_rn(x, M.u_val[pos + lane]));
                        if (term) break;
                    }
                    if (lane == 0) lhs[jj] = x;
                    __syncwarp();
                }
            }
        }
        if (lane == 0 && status) status[s] = BLU_OK;
        __syncwarp();
    }
}

/* get_factors.rs:48-180.  Output in 64-bit indices straight into device staging buffers:
 *   rowperm[m] colperm[m] l_colptr[m+1] l_rowidx/l_value[m+l_nz] u_colptr[m+1] u_rowidx/u_value[m+u_nz]
 * One CTA; the fills run in pivot order on one warp so the rows inside every output column
 * come out ascending, as get_factors.rs:97-113 and :150-167 produce them. */
template <int NT> __global__ void __launch_bounds__(NT) k_get_factors(BluDev D, int s, i64 *rowperm, i64 *colperm,
                                                                       i64 *l_colptr, i64 *l_rowidx, double *l_value,
                                                                       i64 *u_colptr, i64 *u_rowidx, double *u_value, int *status) {
    __shared__ Mat M;
    __shared__ int iscr[40];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) mat_view(M, D, s);
    bsync<NT>();
    const int m = M.m;
    if (M.info->nupdate != 0) { if (tid == 0) *status = BLU_ERROR_INVALID_CALL; return; }
    for (int k = tid; k < m; k += NT) { rowperm[k] = M.pivotrow[k]; colperm[k] = M.pivotcol[k]; }
    int *fill = M.iwork1;   /* m */
    /* L */
    {
        int put = 0;
        for (int base = 0; base < m; base += NT) {
            int k = base + tid;
            int c = k < m ? M.l_begin_p[k + 1] - M.l_begin_p[k] : 0; /* entries + 1 (terminator <-> unit diagonal) */
            int tot, ex = block_excl_scan<NT>(c, &tot, iscr);
            if (k < m) {
                l_colptr[k] = put + ex;
                l_rowidx[put + ex] = k; l_value[put + ex] = 1.0;
                fill[k] = put + ex + 1;
            }
            put += tot;
        }
        if (tid == 0) l_colptr[m] = put;
    }
    bsync<NT>();
    if (wid == 0) {
        for (int k = 0; k < m; k++) {
            const int b = M.lt_begin_p[k], e = M.lt_begin_p[k + 1] - 1;
            for (int pos = b + lane; pos < e; pos += 32) {
                int c = M.prank[M.l_idx[pos]];
                int dst = fill[c]; fill[c] = dst + 1;
                l_rowidx[dst] = k; l_value[dst] = M.l_val[pos];
            }
            __syncwarp();
        }
    }
    bsync<NT>();
    /* U */
    for (int k = tid; k < m; k += NT) fill[k] = 0;
    bsync<NT>();
    for (int j = wid; j < m; j += NT / 32)
        for (int pos = M.lbeg[j] + lane; pos < M.lend[j]; pos += 32) atomicAdd(&fill[M.qrank[M.w_idx[pos]]], 1);
    bsync<NT>();
    {
        int put = 0;
        for (int base = 0; base < m; base += NT) {
            int k = base + tid;
            int c = k < m ? fill[k] + 1 : 0;
            int tot, ex = block_excl_scan<NT>(c, &tot, iscr);
            if (k < m) {
                u_colptr[k] = put + ex;
                u_rowidx[put + ex + c - 1] = k; u_value[put + ex + c - 1] = M.colpiv[M.pivotcol[k]];
                fill[k] = put + ex;
            }
            put += tot;
        }
        if (tid == 0) u_colptr[m] = put;
    }
    bsync<NT>();
    if (wid == 0) {
        for (int k = 0; k < m; k++) {
            const int j = M.pivotcol[k];
            for (int pos = M.lbeg[j] + lane; pos < M.lend[j]; pos += 32) {
                int c = M.qrank[M.w_idx[pos]];
                int dst = fill[c]; fill[c] = dst + 1;
                u_rowidx[dst] = k; u_value[dst] = M.w_val[pos];
            }
            __syncwarp();
        }
        if (lane == 0) *status = BLU_OK;
    }
}

#endif
