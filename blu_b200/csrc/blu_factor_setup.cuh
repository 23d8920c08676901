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

/* ------------------------------------------------------------------ */
/* singletons.rs:287-393 (columns) and 398-503 (rows): the peel as a      */
/* LEVEL-SYNCHRONOUS frontier.  The reference pops a FIFO queue; a line  */
/* enters the queue when its count drops to 1, so the queue is a         */
/* sequence of levels (the singletons present at the start, then those   */
/* created by processing the previous level) and the rank of a pivot is  */
/* its position in that sequence.  One level is processed in parallel:    */
/*  A. every queued line with count 1 finds its cross line and pivot; a  */
/*     pivot of 0 or below abstol is passed over (singletons.rs:353-355,  */
/*     463-465).  Two queued lines with the same cross line: the first    */
/*     in queue order wins, the other becomes empty (its count drops to  */
/*     0 when the winner is processed) -- atomicMin on the cross line.    */
/*  B. winners take consecutive ranks in queue order (block scan); they   */
/*     cannot disturb one another (a winner's only active entry is in    */
/*     its own cross line), so each scans its cross line independently:   */
/*     U row / L column entries in storage order at offsets from a scan   */
/*     of the per-winner counts, iset ^= , count -= 1 by atomics.         */
/*  C. a line whose count reaches 1 joins the next level.  Its place in   */
/*     the queue is that of the decrement that made the count 1, i.e.     */
/*     the LAST decrement it saw (a line that goes on to 0 is popped and  */
/*     skipped, singletons.rs:337-339, so its place is irrelevant): the   */
/*     key (rank of the winner << 32 | position in the winner's cross     */
/*     line) is accumulated with atomicMax and the next level is sorted   */
/*     by it (bitonic, in place).                                         */
/* Scratch: iset = iwork1[0..m), queue = iwork1[m..2m); win = tmpi[0..m), */
/* next = tmpi[m..2m), key = (u64*)(tmpi+2m)[m]; per queue entry: rank    */
/* (pstack), cross line (acols), count / offset (prank); pivots and the   */
/* sort buffers in gwork.                                              */
/* ------------------------------------------------------------------ */

/* ascending bitonic sort of (key, val) pairs, n padded to a power of two with KEY_INF; whole block */
template <int NT> __device__ void block_sort_pairs(u64 *key, int *val, int n) {
    int P = 1;
    while (P < n) P <<= 1;
    for (int t = n + (int)threadIdx.x; t < P; t += NT) { key[t] = KEY_INF; val[t] = -1; }
    bsync<NT>();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < P; t += NT) {
                const int u = t ^ j;
                if (u > t) {
                    const bool up = (t & k) == 0;
                    const u64 a = key[t], b = key[u];
                    if ((a > b) == up) { key[t] = b; key[u] = a; const int x = val[t]; val[t] = val[u]; val[u] = x; }
                }
            }
            bsync<NT>();
        }
    }
}

/* COLS = true: singleton columns (cross line = the row-wise copy, U rows are produced);
 * COLS = false: singleton rows (cross line = the column in the caller's storage, L columns are produced). */
template <int NT, bool COLS> __device__ void singleton_peel(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    int *iset = M.iwork1, *queue = M.iwork1 + m;
    int *win = M.tmpi, *next = M.tmpi + m;
    u64 *key = (u64 *)(M.tmpi + 2 * m);
    int *qrank = M.pstack, *qcnt = M.acols, *qoff = M.prank;      /* per queue entry: rank or -1, cross line or -1, count then offset */
    int *cnt_own = COLS ? M.qinv : M.pinv;      /* -(count) - 1 for active lines, rank >= 0 once pivotal */
    int *cnt_cross = COLS ? M.pinv : M.qinv;
    const double abstol = M.prm.abstol;
    u64 *skey = (u64 *)(M.gwork + m);      /* (slice 0 of gwork is the sparse solves' all-zero solution vector: left alone) */
    int *sval = (int *)(skey + 2 * (size_t)m);
    double *qpiv = (double *)(skey + 3 * (size_t)m);      /* (work0 / work1 stay untouched: the solves rely on work0 being zero) */

    /* level 0: the lines with exactly one entry, ascending index (singletons.rs:316-329, 427-440) */
    int ncur = 0;
    for (int base = 0; base < m; base += NT) {
        const int e = base + tid;
        int is1 = 0;
        if (e < m) {
            win[e] = 0x7fffffff; key[e] = 0;
            if (cnt_own[e] < 0) {
                int x = 0, n;
                if (COLS) { const i64 bb = M.b_begin[e], be = M.b_end[e]; for (i64 pos = bb; pos < be; pos++) x ^= (int)M.b_i[pos]; n = (int)(be - bb); }
                else { const int rb = M.bt_ptr[e], re = M.bt_ptr[e + 1]; for (int pos = rb; pos < re; pos++) x ^= M.bt_idx[pos]; n = re - rb; }
                iset[e] = x;
                cnt_own[e] = -n - 1;
                is1 = n == 1;
            }
        }
        int tot, ex = block_excl_scan<NT>(is1, &tot, S.iscr);
        if (is1) queue[ncur + ex] = e;
        ncur += tot;
    }
    bsync<NT>();
    int rank = S.rank;
    const int rk0 = rank;
    int put = COLS ? M.u_begin[rank] : M.l_begin_p[rank];      /* fill pointer of U (columns pass) / L (rows pass) */

    while (ncur > 0) {
        /* A. pivots and winners */
        for (int q = tid; q < ncur; q += NT) {
            const int e = queue[q];
            double piv = 0.0; int x = -1;
            if (cnt_own[e] == -2) {      /* still exactly one entry */
                x = iset[e];
                if (COLS) { for (i64 pos = M.b_begin[e]; pos < M.b_end[e]; pos++) if ((int)M.b_i[pos] == x) { piv = M.b_x[pos]; break; } }
                else { for (int pos = M.bt_ptr[e]; pos < M.bt_ptr[e + 1]; pos++) if (M.bt_idx[pos] == x) { piv = M.bt_val[pos]; break; } }
                if (piv == 0.0 || fabs(piv) < abstol) x = -1;      /* left to the bump */
                else atomicMin(&win[x], q);
            }
            qcnt[q] = x; qpiv[q] = piv;
        }
        bsync<NT>();
        /* B1. ranks in queue order; mark the pivots */
        int nwin = 0;
        for (int base = 0; base < ncur; base += NT) {
            const int q = base + tid;
            const int x = q < ncur ? qcnt[q] : -1;
            const int w = x >= 0 && win[x] == q;
            int tot, ex = block_excl_scan<NT>(w, &tot, S.iscr);
            if (q < ncur) {
                qrank[q] = w ? rank + nwin + ex : -1;
                if (w) { const int e = queue[q]; cnt_own[e] = rank + nwin + ex; cnt_cross[x] = rank + nwin + ex; M.colpiv[COLS ? e : x] = qpiv[q]; }
            }
            nwin += tot;
        }
        bsync<NT>();
        /* B2. entries each winner will emit (active entries of its cross line; + the terminator of an L column) */
        for (int q = wid; q < ncur; q += NW) {
            if (qrank[q] < 0) { if (lane == 0) qoff[q] = 0; continue; }
            const int x = qcnt[q];
            int n = 0;
            if (COLS) { for (int pos = M.bt_ptr[x] + lane; pos < M.bt_ptr[x + 1]; pos += 32) n += M.qinv[M.bt_idx[pos]] < 0; }
            else { for (i64 pos = M.b_begin[x] + lane; pos < M.b_end[x]; pos += 32) n += M.pinv[(int)M.b_i[pos]] < 0; }
            n = warp_sum(n);
            if (lane == 0) qoff[q] = n + (COLS ? 0 : 1);
        }
        bsync<NT>();
        int produced = 0;
        for (int base = 0; base < ncur; base += NT) {
            const int q = base + tid;
            const int n = q < ncur ? qoff[q] : 0;
            int tot, ex = block_excl_scan<NT>(n, &tot, S.iscr);
            if (q < ncur && qrank[q] >= 0) {
                const int off = put + produced + ex, rk = qrank[q];
                qoff[q] = off;      /* where this winner writes */
                if (COLS) M.u_begin[rk + 1] = off + n;
                else { M.l_begin_p[rk + 1] = off + n; M.l_idx[off + n - 1] = -1; }
            }
            produced += tot;
        }
        bsync<NT>();
        /* B3. scan the cross lines: emit, decrement, collect the next level */
        if (tid == 0) S.ncand = 0;      /* entries of `next` */
        bsync<NT>();
        for (int q = wid; q < ncur; q += NW) {
            const int rk = qrank[q];
            if (rk < 0) continue;
            const int x = qcnt[q];
            const double piv = qpiv[q];
            int off = qoff[q];
            const i64 lb = COLS ? (i64)M.bt_ptr[x] : M.b_begin[x], le = COLS ? (i64)M.bt_ptr[x + 1] : M.b_end[x];
            for (i64 base = lb; base < le; base += 32) {
                const i64 pos = base + lane;
                int o = -1; double v = 0.0; int act = 0;
                if (pos < le) {
                    if (COLS) { o = M.bt_idx[pos]; v = M.bt_val[pos]; act = M.qinv[o] < 0; }
                    else { o = (int)M.b_i[pos]; v = M.b_x[pos]; act = M.pinv[o] < 0; }
                }
                const unsigned am = __ballot_sync(FULLMASK, act);
                if (act) {
                    const int dst = off + __popc(am & lanemask_lt());
                    if (COLS) { M.u_idx[dst] = o; M.u_val[dst] = v; }
                    else { M.l_idx[dst] = o; M.l_val[dst] = __ddiv_rn(v, piv); }
                    atomicXor(&iset[o], x);
                    atomicMax((unsigned long long *)&key[o], ((unsigned long long)(unsigned)rk << 32) | (unsigned long long)(unsigned)(pos - lb));
                    const int old = atomicAdd(&cnt_own[o], 1);
                    if (old == -3) next[atomicAdd(&S.ncand, 1)] = o;      /* count 2 -> 1: a new singleton */
                }
                off += __popc(am);
            }
        }
        bsync<NT>();
        /* restore the cross-line scratch of this level, advance */
        for (int q = tid; q < ncur; q += NT) if (qcnt[q] >= 0) win[qcnt[q]] = 0x7fffffff;
        const int nnext = S.ncand;
        rank += nwin; put += produced;
        bsync<NT>();
        /* C. the next level in queue order */
        for (int t = tid; t < nnext; t += NT) { skey[t] = key[next[t]]; sval[t] = next[t]; }
        bsync<NT>();
        if (nnext > 1) block_sort_pairs<NT>(skey, sval, nnext);
        for (int t = tid; t < nnext; t += NT) queue[t] = sval[t];
        ncur = nnext;
        bsync<NT>();
    }
    /* the lines of the other factor are empty for these pivots: singletons.rs:385-391 / 495-500 */
    if (COLS) {
        const int lpos = M.l_begin_p[rk0];
        for (int rk = rk0 + tid; rk < rank; rk += NT) { M.l_idx[lpos + (rk - rk0)] = -1; M.l_begin_p[rk + 1] = lpos + (rk - rk0) + 1; }
    } else {
        const int upos = M.u_begin[rk0];
        for (int rk = rk0 + tid; rk < rank; rk += NT) M.u_begin[rk + 1] = upos;
    }
    /* gwork is the per-warp scatter space of the pivot steps and must be all zero when they start */
    for (size_t t = tid; t < 4 * (size_t)m; t += NT) M.gwork[m + t] = 0.0;
    if (tid == 0) S.rank = rank;
    bsync<NT>();
}
template <int NT> __device__ void singleton_cols(Shm &S) { singleton_peel<NT, true>(S); }
template <int NT> __device__ void singleton_rows(Shm &S) { singleton_peel<NT, false>(S); }

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
