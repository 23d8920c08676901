/* blu_factor_setup.cuh -- phases 1 and 2 of the factorization on the device:
 * validation + row-wise copy + singleton peeling (reference: src/lu/singletons.rs:81-503)
 * and the bump set-up (src/lu/setup_bump.rs:55-264).  One CTA per basis matrix. */
#ifndef BLU_FACTOR_SETUP_CUH
#define BLU_FACTOR_SETUP_CUH
#include "blu_dev_common.cuh"

/* Sort one short line of (idx,val) ascending by idx; one thread.  Rows of the row-wise
 * copy are filled with atomics, so their order is restored here: the reference fills
 * them by scanning j = 0..m (singletons.rs:186-198), i.e. ascending column index. */
__device__ __forceinline__ void insertion_sort_line(int *idx, double *val, int n) {
    for (int a = 1; a < n; a++) {
        int k = idx[a]; double v = val[a];
        int b = a - 1;
        while (b >= 0 && idx[b] > k) { idx[b + 1] = idx[b]; val[b + 1] = val[b]; b--; }
        idx[b + 1] = k; val[b + 1] = v;
    }
}

/* Warp-cooperative rank sort for long lines: O(n^2/32), scratch in (sidx,sval). */
__device__ __forceinline__ void warp_rank_sort_line(int *idx, double *val, int n, int *sidx, double *sval) {
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < n; base += 32) {
        int e = base + lane;
        int ke = e < n ? idx[e] : 0x7fffffff;
        int rank = 0;
        for (int f = 0; f < n; f++) {
            int kf = idx[f];
            rank += (kf < ke) || (kf == ke && f < e);
        }
        if (e < n) { sidx[rank] = ke; sval[rank] = val[e]; }
    }
    __syncwarp();
    for (int e = lane; e < n; e += 32) { idx[e] = sidx[e]; val[e] = sval[e]; }
    __syncwarp();
}

/* singletons.rs:116-201: validate B, count, build the row-wise copy. */
template <int NT> __device__ void phase_validate_transpose(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    int bad = 0;
    i64 nnz = 0;
    for (int j = tid; j < m; j += NT) {
        i64 bb = M.b_begin[j], be = M.b_end[j];
        /* (a pointer outside b_i / b_x is an invalid argument too: the reference would panic on the slice) */
        if (be < bb || bb < 0 || be > M.b_total) bad = 1; else nnz += be - bb;
    }
    bad = block_max<NT>(bad, S.iscr);
    nnz = block_sum64<NT>(nnz, S.kscr);
    if (bad) { if (tid == 0) S.status = BLU_ERROR_INVALID_ARGUMENT; bsync<NT>(); return; }
    if (tid == 0) {
        BluInfo *I = M.info;
        I->matrix_nz = nnz;
        int ok = 1;
        if ((i64)M.l_mem < nnz) { I->addmem_l = nnz - M.l_mem; ok = 0; }
        if ((i64)M.u_mem < nnz) { I->addmem_u = nnz - M.u_mem; ok = 0; }
        if ((i64)M.w_mem < nnz) { I->addmem_w = nnz - M.w_mem; ok = 0; }
        if (!ok) S.status = BLU_REALLOCATE;
        else if ((i64)M.bnz_cap < nnz) { S.status = BLU_ERROR_INTERNAL; I->internal_error = __LINE__; }
    }
    bsync<NT>();
    if (S.status != BLU_OK) return;

    /* row counts + index range, singletons.rs:154-173 */
    int *cnt = M.iwork1;            /* m */
    for (int i = tid; i < m; i += NT) cnt[i] = 0;
    bsync<NT>();
    for (int j = tid; j < m; j += NT) {
        for (i64 pos = M.b_begin[j]; pos < M.b_end[j]; pos++) {
            i64 i = M.b_i[pos];
            if (i < 0 || i >= m) bad = 1; else atomicAdd(&cnt[(int)i], 1);
        }
    }
    bad = block_max<NT>(bad, S.iscr);
    if (bad) { if (tid == 0) S.status = BLU_ERROR_INVALID_ARGUMENT; bsync<NT>(); return; }

    /* exclusive scan of the counts -> bt_ptr; cnt becomes the fill pointer */
    int running = 0;
    for (int base = 0; base < m; base += NT) {
        int i = base + tid;
        int c = i < m ? cnt[i] : 0, tot;
        int ex = block_excl_scan<NT>(c, &tot, S.iscr);
        if (i < m) { M.bt_ptr[i] = running + ex; cnt[i] = running + ex; }
        running += tot;
    }
    if (tid == 0) M.bt_ptr[m] = running;
    bsync<NT>();
    for (int j = tid; j < m; j += NT) {
        for (i64 pos = M.b_begin[j]; pos < M.b_end[j]; pos++) {
            int i = (int)M.b_i[pos];
            int put = atomicAdd(&cnt[i], 1);
            M.bt_idx[put] = j;
            M.bt_val[put] = M.b_x[pos];
        }
    }
    bsync<NT>();
    /* restore ascending-j order inside each row; detect duplicates (singletons.rs:194-200) */
    for (int i = tid; i < m; i += NT) {
        int rb = M.bt_ptr[i], n = M.bt_ptr[i + 1] - rb;
        if (n > 1 && n <= 32) insertion_sort_line(M.bt_idx + rb, M.bt_val + rb, n);
    }
    bsync<NT>();
    {   /* long rows: one warp each, scratch in the (still unused) W arena split per warp */
        const int chunk = (2 * M.w_mem) / NW;
        for (int i = wid; i < m; i += NW) {
            int rb = M.bt_ptr[i], n = M.bt_ptr[i + 1] - rb;
            if (n > 32) {
                if (n > chunk) { if (lane == 0) BLU_CHECK(S, 0); }
                else warp_rank_sort_line(M.bt_idx + rb, M.bt_val + rb, n, M.w_idx + wid * chunk, M.w_val + wid * chunk);
            }
        }
    }
    bsync<NT>();
    for (int i = tid; i < m; i += NT) {
        int rb = M.bt_ptr[i], re = M.bt_ptr[i + 1];
        for (int pos = rb + 1; pos < re; pos++) if (M.bt_idx[pos] == M.bt_idx[pos - 1]) bad = 1;
    }
    bad = block_max<NT>(bad, S.iscr);
    if (bad) { if (tid == 0) S.status = BLU_ERROR_INVALID_ARGUMENT; }
    bsync<NT>();
}

/* singletons.rs:287-393 (columns) and 398-503 (rows).  The FIFO queue is processed
 * by warp 0 in the reference's order; the lanes share the scan of each pivot row /
 * column so one queue entry costs a handful of dependent loads instead of O(nz).
 * iset = iwork1[0..m), queue = iwork1[m..2m). */
template <int NT> __device__ void singleton_cols(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int *iset = M.iwork1, *queue = M.iwork1 + m;
    const double abstol = M.prm.abstol;
    int tail = 0;
    for (int base = 0; base < m; base += NT) {
        int j = base + tid;
        int is1 = 0;
        if (j < m && M.qinv[j] < 0) {
            i64 bb = M.b_begin[j], be = M.b_end[j];
            int x = 0;
            for (i64 pos = bb; pos < be; pos++) x ^= (int)M.b_i[pos];
            iset[j] = x;
            M.qinv[j] = -(int)(be - bb) - 1;
            is1 = (be - bb) == 1;
        }
        int tot, ex = block_excl_scan<NT>(is1, &tot, S.iscr);
        if (is1) queue[tail + ex] = j;
        tail += tot;
    }
    bsync<NT>();
    if (wid == 0) {
        int rank = S.rank;
        const int rk0 = rank;
        int uput = M.u_begin[rank];
        for (int front = 0; front < tail; front++) {
            const int j = queue[front];
            if (M.qinv[j] == -1) continue;          /* column became empty meanwhile */
            const int i = iset[j];
            const int rb = M.bt_ptr[i], re = M.bt_ptr[i + 1];
            double piv = 0.0;
            for (int base = rb; base < re; base += 32) {
                int pos = base + lane;
                int hit = pos < re && M.bt_idx[pos] == j;
                unsigned hm = __ballot_sync(FULLMASK, hit);
                if (hm) {
                    double v = hit ? M.bt_val[pos] : 0.0;
                    piv = __shfl_sync(FULLMASK, v, __ffs((int)hm) - 1);
                    break;
                }
            }
            if (piv == 0.0 || fabs(piv) < abstol) continue; /* leave to the bump */
            if (lane == 0) { M.qinv[j] = rank; M.pinv[i] = rank; }
            __syncwarp();
            for (int base = rb; base < re; base += 32) {
                int pos = base + lane;
                int j2 = -1; double v = 0.0; int act = 0;
                if (pos < re) { j2 = M.bt_idx[pos]; v = M.bt_val[pos]; act = M.qinv[j2] < 0; }
                unsigned am = __ballot_sync(FULLMASK, act);
                int enq = 0;
                if (act) {
                    int dst = uput + __popc(am & lanemask_lt());
                    M.u_idx[dst] = j2; M.u_val[dst] = v;
                    iset[j2] ^= i;
                    int q = M.qinv[j2] + 1;
                    M.qinv[j2] = q;
                    enq = q == -2;
                }
                unsigned em = __ballot_sync(FULLMASK, enq);
                if (enq) queue[tail + __popc(em & lanemask_lt())] = j2;
                uput += __popc(am);
                tail += __popc(em);
            }
            if (lane == 0) { M.u_begin[rank + 1] = uput; M.colpiv[j] = piv; }
            rank++;
            __syncwarp();
        }
        /* empty L columns, singletons.rs:385-391 */
        int lpos = M.l_begin_p[rk0];
        for (int rk = rk0 + lane; rk < rank; rk += 32) {
            M.l_idx[lpos + (rk - rk0)] = -1;
            M.l_begin_p[rk + 1] = lpos + (rk - rk0) + 1;
        }
        if (lane == 0) S.rank = rank;
    }
    bsync<NT>();
}

template <int NT> __device__ void singleton_rows(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int *iset = M.iwork1, *queue = M.iwork1 + m;
    const double abstol = M.prm.abstol;
    int tail = 0;
    for (int base = 0; base < m; base += NT) {
        int i = base + tid;
        int is1 = 0;
        if (i < m && M.pinv[i] < 0) {
            int rb = M.bt_ptr[i], re = M.bt_ptr[i + 1];
            int x = 0;
            for (int pos = rb; pos < re; pos++) x ^= M.bt_idx[pos];
            iset[i] = x;
            M.pinv[i] = -(re - rb) - 1;
            is1 = (re - rb) == 1;
        }
        int tot, ex = block_excl_scan<NT>(is1, &tot, S.iscr);
        if (is1) queue[tail + ex] = i;
        tail += tot;
    }
    bsync<NT>();
    if (wid == 0) {
        int rank = S.rank;
        const int rk0 = rank;
        int lput = M.l_begin_p[rank];
        for (int front = 0; front < tail; front++) {
            const int i = queue[front];
            if (M.pinv[i] == -1) continue;
            const int j = iset[i];
            const i64 cb = M.b_begin[j], ce = M.b_end[j];
            double piv = 0.0;
            for (i64 base = cb; base < ce; base += 32) {
                i64 pos = base + lane;
                int hit = pos < ce && (int)M.b_i[pos] == i;
                unsigned hm = __ballot_sync(FULLMASK, hit);
                if (hm) {
                    double v = hit ? M.b_x[pos] : 0.0;
                    piv = __shfl_sync(FULLMASK, v, __ffs((int)hm) - 1);
                    break;
                }
            }
            if (piv == 0.0 || fabs(piv) < abstol) continue;
            if (lane == 0) { M.qinv[j] = rank; M.pinv[i] = rank; }
            __syncwarp();
            for (i64 base = cb; base < ce; base += 32) {
                i64 pos = base + lane;
                int i2 = -1; double v = 0.0; int act = 0;
                if (pos < ce) { i2 = (int)M.b_i[pos]; v = M.b_x[pos]; act = M.pinv[i2] < 0; }
                unsigned am = __ballot_sync(FULLMASK, act);
                int enq = 0;
                if (act) {
                    int dst = lput + __popc(am & lanemask_lt());
                    M.l_idx[dst] = i2; M.l_val[dst] = __ddiv_rn(v, piv);
                    iset[i2] ^= j;
                    int q = M.pinv[i2] + 1;
                    M.pinv[i2] = q;
                    enq = q == -2;
                }
                unsigned em = __ballot_sync(FULLMASK, enq);
                if (enq) queue[tail + __popc(em & lanemask_lt())] = i2;
                lput += __popc(am);
                tail += __popc(em);
            }
            if (lane == 0) { M.l_idx[lput] = -1; M.l_begin_p[rank + 1] = lput + 1; M.colpiv[j] = piv; }
            lput++;
            rank++;
            __syncwarp();
        }
        /* empty U rows, singletons.rs:495-500 */
        int upos = M.u_begin[rk0];
        for (int rk = rk0 + lane; rk < rank; rk += 32) M.u_begin[rk + 1] = upos;
        if (lane == 0) S.rank = rank;
    }
    bsync<NT>();
}

/* singletons.rs:81-264 */
template <int NT> __device__ void phase_singletons(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x;
    i64 t0 = clock64();
    phase_validate_transpose<NT>(S);
    if (tid == 0) S.t_phase[0] += clock64() - t0;
    if (S.status != BLU_OK) return;
    t0 = clock64();
    for (int i = tid; i < m; i += NT) { M.pinv[i] = -1; M.qinv[i] = -1; }
    if (tid == 0) { M.l_begin_p[0] = 0; M.u_begin[0] = 0; S.rank = 0; }
    bsync<NT>();
    if (M.prm.nzbias >= 0) { singleton_cols<NT>(S); singleton_rows<NT>(S); }
    else { singleton_rows<NT>(S); singleton_cols<NT>(S); }
    for (int i = tid; i < m; i += NT) {
        if (M.pinv[i] < 0) M.pinv[i] = -1;
        if (M.qinv[i] < 0) M.qinv[i] = -1;
    }
    bsync<NT>();
    if (tid == 0) S.t_phase[1] += clock64() - t0;
}

/* setup_bump.rs:55-264.  Lines of the W file get `stretch*nz + pad` slack; the bucket
 * lists of the reference (list.rs) are replaced by keys (count<<40 | stamp) with the
 * initial stamp = index, which is the order list_add produces at setup_bump.rs:161-168
 * and 202-209. */
template <int NT> __device__ void phase_setup_bump(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x;
    const int rank = S.rank;
    const double abstol = M.prm.abstol;
    const i64 l_nz = M.l_begin_p[rank] - rank, u_nz = M.u_begin[rank];
    i64 bump_nz = M.info->matrix_nz - l_nz - u_nz - rank;
    {
        i64 need = bump_nz + (i64)(M.prm.stretch * (double)bump_nz) + (i64)(m - rank) * M.prm.pad;
        need *= 2;
        if (need > (i64)M.w_mem) {
            if (tid == 0) { M.info->addmem_w = need - M.w_mem; S.status = BLU_REALLOCATE; }
            bsync<NT>();
            return;
        }
    }
    bsync<NT>();
    /* pass 1: active count and max per bump column (setup_bump.rs:132-160) */
    int *ccnt = M.tmpi;         /* m */
    int *rcnt = M.tmpi + m;     /* m */
    i64 dropped = 0;
    for (int j = tid; j < m; j += NT) {
        int cnz = 0; double cmx = 0.0;
        if (M.qinv[j] < 0) {
            for (i64 pos = M.b_begin[j]; pos < M.b_end[j]; pos++) {
                int i = (int)M.b_i[pos];
                if (M.pinv[i] >= 0) continue;
                cmx = fmax(cmx, fabs(M.b_x[pos]));
                cnz++;
            }
            if (cmx == 0.0 || cmx < abstol) { dropped += cnz; cnz = 0; cmx = 0.0; }
            M.colpiv[j] = cmx;
        }
        ccnt[j] = cnz;
    }
    dropped = block_sum64<NT>(dropped, S.kscr);
    bump_nz -= dropped;
    /* column lines */
    int put = 0, nact = 0;
    for (int base = 0; base < m; base += NT) {
        int j = base + tid;
        int active = j < m && M.qinv[j] < 0;
        int cnz = active ? ccnt[j] : 0;
        int sz = (active && cnz > 0) ? cnz + slack_of(M.prm, cnz) : 0;
        int tot, ex = block_excl_scan<NT>(sz, &tot, S.iscr);
        int tota, exa = block_excl_scan<NT>(active, &tota, S.iscr);
        if (j < m) {
            if (active) {
                int b = cnz > 0 ? put + ex : 0;
                M.lbeg[j] = b; M.lcap[j] = b + sz;
                int w = b;
                if (cnz > 0) {
                    for (i64 pos = M.b_begin[j]; pos < M.b_end[j]; pos++) {
                        int i = (int)M.b_i[pos];
                        if (M.pinv[i] >= 0) continue;
                        M.w_idx[w] = i; M.w_val[w] = M.b_x[pos]; w++;
                    }
                }
                M.lend[j] = w;
                M.ckey[j] = mkkey(cnz, j);
                M.acols[nact + exa] = j;
            } else {
                M.lbeg[j] = M.lend[j] = M.lcap[j] = 0;
                M.ckey[j] = KEY_INF;
            }
        }
        put += tot; nact += tota;
    }
    bsync<NT>();
    /* row lines: pattern = row of the row-wise copy restricted to live bump columns
     * (ascending column index, setup_bump.rs:217-224) */
    for (int i = tid; i < m; i += NT) {
        int rnz = 0;
        if (M.pinv[i] < 0) {
            for (int pos = M.bt_ptr[i]; pos < M.bt_ptr[i + 1]; pos++) {
                int j = M.bt_idx[pos];
                if (M.qinv[j] < 0 && ccnt[j] > 0) rnz++;
            }
        }
        rcnt[i] = rnz;
    }
    bsync<NT>();
    for (int base = 0; base < m; base += NT) {
        int i = base + tid;
        int active = i < m && M.pinv[i] < 0;
        int rnz = active ? rcnt[i] : 0;
        int sz = active ? rnz + slack_of(M.prm, rnz) : 0;
        int tot, ex = block_excl_scan<NT>(sz, &tot, S.iscr);
        if (i < m) {
            if (active) {
                int b = put + ex;
                M.lbeg[m + i] = b; M.lcap[m + i] = b + sz;
                int w = b;
                for (int pos = M.bt_ptr[i]; pos < M.bt_ptr[i + 1]; pos++) {
                    int j = M.bt_idx[pos];
                    if (M.qinv[j] < 0 && ccnt[j] > 0) M.w_idx[w++] = j;
                }
                M.lend[m + i] = w;
                M.rkey[i] = mkkey(rnz, i);
            } else {
                M.lbeg[m + i] = M.lend[m + i] = M.lcap[m + i] = 0;
                M.rkey[i] = KEY_INF;
            }
        }
        put += tot;
    }
    if (tid == 0) {
        BLU_CHECK(S, put <= M.w_mem);
        S.w_half = 0; S.w_used = put; S.w_limit = M.w_mem;
        S.cstamp = m; S.rstamp = m;
        S.nact = nact; S.ndead = 0; S.rankdef = 0;
        M.info->bump_nz = bump_nz;
        M.info->bump_size = m - rank;
    }
    bsync<NT>();
}

#endif
