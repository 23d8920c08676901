/* blu_factor_dense.cuh -- the dense tail of the bump factorization.
 *
 * The reference has no dense-tail switch (src/lu/factorize_bump.rs:12-49 runs the sparse pivot() to the
 * end), so this representation must reproduce what markowitz.rs:34-193 and pivot.rs:114-833 would do on
 * the line file, bit for bit.  Once the active submatrix has at most `kd` rows it is held as
 *
 *   dv[t*kd + c]       the value of entry (row slot t, column slot c), 0.0 when absent   (row-major)
 *   dn_key[t*kd + c]   .c = storage-order key of the entry inside its column (the column file of file.rs),
 *                      .r = storage-order key inside its row (the row file); only order matters
 *   rb, cb             presence bitmaps by row and by column
 *
 * RES = true: dv, rb and cb live in shared memory (one CTA per SM, kd <= 160 on B200: 200 KB of values);
 * the keys stay in HBM/L2 and are written fire-and-forget by the update and read O(kd) times per step.
 * RES = false: dv, rb, cb are the HBM arrays dn_val, dn_rbits, dn_cbits (any kd).
 *
 * What the line file encodes by position is carried by the keys:
 *   - markowitz.rs:96-112 takes, per candidate column, the first entry in storage order among those of
 *     minimum Markowitz cost  ==  argmin over (cost, key.c);
 *   - pivot.rs:228-304 rebuilds column j of the pivot row as [entries outside the pivot column, with the
 *     first of them and the pivot-row entry exchanged and the pivot-row entry dropped] ++ [one entry per
 *     row of the pivot column, in the pivot column's order]  ==  the first kept entry inherits the key of
 *     the pivot-row entry, every updated entry gets (epoch + its position in the pivot column);
 *   - pivot.rs:335-398 rebuilds row i as [entries outside the pivot row] ++ [the pivot row's columns in
 *     order]  ==  every updated entry gets (epoch + its position in the pivot row);
 *   - pivot_small (pivot.rs:646-662, 748-755) drops |x| <= droptol: the presence bit is cleared.
 * Arithmetic is the same __ddiv_rn / __dmul_rn / __dsub_rn per entry as the sparse step, so values are
 * bit-identical.  One elimination step then costs one coalesced read-modify-write per entry of the rank-1
 * update and O(kd) bookkeeping, instead of one ballot/prefix compaction per line.
 *
 * Only the general step (pivot_any / pivot_small) and the rank-deficiency step run here; when the search
 * picks a singleton row/column or a doubleton column (pivot.rs:835-1331), or a column has to be emptied
 * (pivot.rs:96-106, 1333-1381), the lines are written back to the W file in key order (dense_exit) and the
 * sparse code continues.
 */
#ifndef BLU_FACTOR_DENSE_CUH
#define BLU_FACTOR_DENSE_CUH
/* included by blu_factor_bump.cuh after its helpers (finish_step, warp_squeeze) */

/* Pointers into the dynamic shared memory, computed locally from its base in every function so that the
 * compiler knows their address space (LDS/STS instead of generic loads, and no aliasing with the global
 * stores of the update loop). */
struct DenseSm {
    u64 *skeyc;            /* Markowitz key of column slot c (count<<40 | stamp), KEY_INF when gone */
    u64 *scm;              /* colmax of column slot c as the bits of a non-negative double */
    double *cvalp;         /* values of the pivot column in its storage order */
    int *drow, *dcol;      /* slot -> row / column index */
    unsigned *keyc, *keyr; /* sort keys of the pivot column / row */
    u64 *sdrop;            /* pivot_small: cancellation mask per column of the pivot row (overlays keyc|keyr) */
    unsigned *candk;       /* key stash of the search candidates */
    BluKey2 *rowk;         /* the pivot row's keys, fetched by one bulk copy (16-byte aligned: kd is a multiple of 32) */
    unsigned *cmask, *rmask, *rfull;
    unsigned short *clist, *rlist, *posr, *rnz, *cnz, *tmps, *tmpr;
    unsigned *rb_s, *cb_s; double *dv_s;   /* resident bitmaps and values (RES only) */
};
__device__ __forceinline__ void dense_view(DenseSm &d, unsigned char *dyn, int KD, int KW) {
    u64 *p8 = (u64 *)dyn;
    d.skeyc = p8; p8 += KD;
    d.scm = p8; p8 += KD;
    d.cvalp = (double *)p8; p8 += KD;
    int *p4 = (int *)p8;
    d.drow = p4; p4 += KD;
    d.dcol = p4; p4 += KD;
    d.keyc = (unsigned *)p4; d.sdrop = (u64 *)p4; p4 += KD;
    d.keyr = (unsigned *)p4; p4 += KD;
    d.rowk = (BluKey2 *)p4; p4 += KD;
    d.candk = (unsigned *)p4; p4 += DENSE_STASH * KD;
    d.cmask = (unsigned *)p4; p4 += KW;
    d.rmask = (unsigned *)p4; p4 += KW;
    d.rfull = (unsigned *)p4; p4 += KW;
    unsigned short *p2 = (unsigned short *)p4;
    d.clist = p2; p2 += KD;
    d.rlist = p2; p2 += KD;
    d.posr = p2; p2 += KD;
    d.rnz = p2; p2 += KD;
    d.cnz = p2; p2 += KD;
    d.tmps = p2; p2 += KD;
    d.tmpr = p2; p2 += KD;
    unsigned *b = (unsigned *)(dyn + blu_dense_smem_bytes(KD));
    d.rb_s = b; d.cb_s = b + KD * KW;
    d.dv_s = (double *)(b + 2 * KD * KW);
}
#define DENSE_VIEW(RES) \
    BLU_DYN_SMEM(dyn_); DenseSm d; dense_view(d, dyn_, S.kd, S.kw); \
    double *const dv = RES ? d.dv_s : S.M.dn_val; \
    unsigned *const rbm = RES ? d.rb_s : S.M.dn_rbits; \
    unsigned *const cbm = RES ? d.cb_s : S.M.dn_cbits; \
    BluKey2 *const dkey = S.M.dn_key

__device__ __forceinline__ int bit_test(const unsigned *row, int t) { return (row[t >> 5] >> (t & 31)) & 1; }
/* number of set bits below position t */
__device__ __forceinline__ int bits_rank(const unsigned *row, int t) {
    const int w = t >> 5;
    int r = 0;
    for (int q = 0; q < w; q++) r += __popc(row[q]);
    return r + __popc(row[w] & ((1u << (t & 31)) - 1u));
}

/* ------------------------------------------------------------------ */
/* line file -> dense arrays                                           */
/* ------------------------------------------------------------------ */
template <int NT, bool RES> __device__ __noinline__ void dense_enter(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int KD = S.kd, KW = S.kw;
    DENSE_VIEW(RES);
    int *rowslot = M.iwork1, *colslot = M.iwork1 + m;
    /* slots in ascending index order */
    int nr = 0, nc = 0;
    for (int base = 0; base < m; base += NT) {
        const int i = base + tid;
        const int ra = i < m && M.pinv[i] < 0;
        const int ca = i < m && M.ckey[i] != KEY_INF;
        int totr, exr = block_excl_scan<NT>(ra, &totr, S.iscr);
        int totc, exc = block_excl_scan<NT>(ca, &totc, S.iscr);
        if (i < m) {
            rowslot[i] = ra ? nr + exr : -1;
            colslot[i] = ca ? nc + exc : -1;
            if (ra && nr + exr < KD) d.drow[nr + exr] = i;
            if (ca && nc + exc < KD) d.dcol[nc + exc] = i;
        }
        nr += totr; nc += totc;
    }
    if (nr > KD || nc > KD) { if (tid == 0) BLU_CHECK(S, 0); bsync<NT>(); return; }
    for (size_t q = tid; q < (size_t)nr * KD; q += NT) dv[q] = 0.0;
    for (int q = tid; q < KD * KW; q += NT) { rbm[q] = 0; cbm[q] = 0; }
    for (int c = tid; c < KD; c += NT) { d.skeyc[c] = KEY_INF; d.cnz[c] = 0; d.rnz[c] = 0; }
    bsync<NT>();
    for (int c = wid; c < nc; c += NW) {
        const int j = d.dcol[c];
        const int b = M.lbeg[j], e = M.lend[j];
        for (int pos = b + lane; pos < e; pos += 32) {
            const int t = rowslot[M.w_idx[pos]];
            const size_t off = (size_t)t * KD + c;
            dv[off] = M.w_val[pos];
            dkey[off].c = (unsigned short)(pos - b);
            atomicOr(&rbm[t * KW + (c >> 5)], 1u << (c & 31));
            atomicOr(&cbm[c * KW + (t >> 5)], 1u << (t & 31));
        }
        if (lane == 0) { d.cnz[c] = (unsigned short)(e - b); d.skeyc[c] = M.ckey[j]; d.scm[c] = (u64)__double_as_longlong(M.colpiv[j]); }
    }
    for (int t = wid; t < nr; t += NW) {
        const int i = d.drow[t];
        const int b = M.lbeg[m + i], e = M.lend[m + i];
        for (int pos = b + lane; pos < e; pos += 32) {
            const int c = colslot[M.w_idx[pos]];
            dkey[(size_t)t * KD + c].r = (unsigned short)(pos - b);
        }
        if (lane == 0) d.rnz[t] = (unsigned short)(e - b);
    }
    if (tid == 0) {
        S.dense = 1; S.nrs = nr; S.ncs = nc;
        S.epoch = 1;          /* the position keys handed out here are epoch 0 */
        S.dense_entries++;
        S.n_kind[6]++;
    }
    bsync<NT>();
}

/* ------------------------------------------------------------------ */
/* dense arrays -> line file (key order == storage order)              */
/* ------------------------------------------------------------------ */
template <int NT, bool RES> __device__ __noinline__ void dense_exit(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int KD = S.kd, KW = S.kw, nr = S.nrs, nc = S.ncs;
    DENSE_VIEW(RES);
    const int base = S.w_half * M.w_mem;
    int put = 0, nact = 0;
    for (int b0 = 0; b0 < nc; b0 += NT) {
        const int c = b0 + tid;
        const int alive = c < nc && d.skeyc[c] != KEY_INF;
        const int n = alive ? d.cnz[c] : 0;
        const int sz = n > 0 ? n + slack_of(M.prm, n) : 0;
        int tot, ex = block_excl_scan<NT>(sz, &tot, S.iscr);
        int tota, exa = block_excl_scan<NT>(alive, &tota, S.iscr);
        if (alive) {
            const int j = d.dcol[c], b = base + put + ex;
            M.lbeg[j] = b; M.lend[j] = b + n; M.lcap[j] = b + sz;
            M.ckey[j] = d.skeyc[c];
            M.colpiv[j] = __longlong_as_double((long long)d.scm[c]);
            M.acols[nact + exa] = j;
        }
        put += tot; nact += tota;
    }
    for (int b0 = 0; b0 < nr; b0 += NT) {
        const int t = b0 + tid;
        const int alive = t < nr && M.rkey[d.drow[t]] != KEY_INF;
        const int n = alive ? d.rnz[t] : 0;
        const int sz = alive ? n + slack_of(M.prm, n) : 0;
        int tot, ex = block_excl_scan<NT>(sz, &tot, S.iscr);
        if (alive) {
            const int i = d.drow[t], b = base + put + ex;
            M.lbeg[m + i] = b; M.lend[m + i] = b + n; M.lcap[m + i] = b + sz;
        }
        put += tot;
    }
    if (put > M.w_mem) {
        if (tid == 0) { M.info->addmem_w = put - M.w_mem; S.status = BLU_REALLOCATE; }
        bsync<NT>();
        return;
    }
    bsync<NT>();
    for (int c = wid; c < nc; c += NW) {
        if (d.skeyc[c] == KEY_INF) continue;
        const int b = M.lbeg[d.dcol[c]];
        const unsigned *cb = cbm + c * KW;
        for (int t = lane; t < nr; t += 32) {
            if (!bit_test(cb, t)) continue;
            const size_t off = (size_t)t * KD + c;
            const unsigned my = dkey[off].c;
            int r = 0;
            for (int w = 0; w < KW; w++) {
                unsigned word = cb[w];
                while (word) { const int t2 = w * 32 + __ffs((int)word) - 1; word &= word - 1; r += dkey[(size_t)t2 * KD + c].c < my; }
            }
            M.w_idx[b + r] = d.drow[t]; M.w_val[b + r] = dv[off];
        }
    }
    for (int t = wid; t < nr; t += NW) {
        const int i = d.drow[t];
        if (M.rkey[i] == KEY_INF) continue;
        const int b = M.lbeg[m + i];
        const unsigned *rb = rbm + t * KW;
        for (int c = lane; c < nc; c += 32) {
            if (!bit_test(rb, c)) continue;
            const unsigned my = dkey[(size_t)t * KD + c].r;
            int r = 0;
            for (int w = 0; w < KW; w++) {
                unsigned word = rb[w];
                while (word) { const int c2 = w * 32 + __ffs((int)word) - 1; word &= word - 1; r += dkey[(size_t)t * KD + c2].r < my; }
            }
            M.w_idx[b + r] = d.dcol[c];
        }
    }
    bsync<NT>();
    /* the dense-step arrays overlay the row/column marks of the sparse fast path */
    if (S.smarks) for (int i = tid; i < m; i += NT) { S.rm[i] = 0; S.cm[i] = 0; }
    if (tid == 0) {
        S.dense = 0;
        S.w_used = base + put; S.w_limit = base + M.w_mem;
        S.nact = nact; S.ndead = 0;
        S.dense_block_rank = S.rank + 8;
        S.wc = -1; S.wr = -1;
    }
    bsync<NT>();
}

/* ------------------------------------------------------------------ */
/* markowitz.rs:34-123 on the dense arrays (search_rows == 0)          */
/* ------------------------------------------------------------------ */
template <int NT, bool RES> __device__ __noinline__ void dense_search(Shm &S) {
    Mat &M = S.M;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int KD = S.kd, KW = S.kw, nr = S.nrs, nc = S.ncs;
    DENSE_VIEW(RES);
    int maxsearch = M.prm.maxsearch;
    if (maxsearch < 1) maxsearch = 1;
    if (maxsearch > MAXCAND) maxsearch = MAXCAND;
    /* the first `maxsearch` live columns in ascending (count, stamp) order: one warp, no block barriers */
    if (wid == 0) {
        int ncand = 0;
        u64 prev = 0; int have_prev = 0;
        while (ncand < maxsearch) {
            u64 k0 = KEY_INF, k1 = KEY_INF, k2 = KEY_INF;
            int j0 = -1, j1 = -1, j2 = -1;
            for (int c = lane; c < nc; c += 32) {
                const u64 k = d.skeyc[c];
                if (k >= KEY_PARK || (have_prev && k <= prev)) continue;
                if (k < k2) {
                    if (k < k1) {
                        k2 = k1; j2 = j1;
                        if (k < k0) { k1 = k0; j1 = j0; k0 = k; j0 = c; }
                        else { k1 = k; j1 = c; }
                    } else { k2 = k; j2 = c; }
                }
            }
            int got = 0;
            for (int r = 0; r < 3 && ncand < maxsearch; r++) {
                const u64 best = warp_min64(k0);
                if (best == KEY_INF) break;
                if (k0 == best) {
                    S.cand_col[ncand] = j0;
                    k0 = k1; j0 = j1; k1 = k2; j1 = j2; k2 = KEY_INF; j2 = -1;
                }
                prev = best; have_prev = 1;
                ncand++; got++;
            }
            if (got < 3) break;
        }
        if (lane == 0) S.ncand = ncand;
    }
    bsync<NT>();
    const int ncand = S.ncand;
    if (ncand == 0) { if (tid == 0) { BLU_CHECK(S, 0); } bsync<NT>(); return; }
    if (key_cnt(d.skeyc[S.cand_col[0]]) == 0) {      /* markowitz.rs:73-78 */
        if (tid == 0) { S.dpc = S.cand_col[0]; S.pivot_col = d.dcol[S.cand_col[0]]; S.pivot_row = -1; }
        bsync<NT>();
        return;
    }
    const double abstol = M.prm.abstol, reltol = M.prm.reltol;
    for (int cc = wid; cc < ncand; cc += NW) {
        const int c = S.cand_col[cc];
        const i64 nz1 = d.cnz[c];
        const double cmx = __longlong_as_double((long long)d.scm[c]);
        const double tol = fmax(abstol, reltol * cmx);
        const unsigned *cb = cbm + c * KW;
        u64 best = KEY_INF; int bt = -1;
        if (cc < DENSE_STASH) {
            /* all key loads of the column in flight together; the pivot step reuses them */
            unsigned *stash = d.candk + cc * KD;
            {   /* kd <= 256: at most eight rows per lane, their key loads issued back to back (one round trip) */
                unsigned kq[8];
                #pragma unroll
                for (int u = 0; u < 8; u++) { const int t = lane + 32 * u; kq[u] = (t < nr && bit_test(cb, t)) ? dkey[(size_t)t * KD + c].c : 0xffffffffu; }
                #pragma unroll
                for (int u = 0; u < 8; u++) { const int t = lane + 32 * u; if (t < nr) stash[t] = kq[u]; }
            }
            __syncwarp();
            for (int t = lane; t < nr; t += 32) {
                const unsigned kq = stash[t];
                if (kq == 0xffffffffu) continue;
                const double x = fabs(dv[(size_t)t * KD + c]);
                if (x == 0.0 || x < tol) continue;
                const u64 mc = (u64)((nz1 - 1) * (i64)(d.rnz[t] - 1));
                const u64 key = (mc << 32) | (u64)kq;       /* ties: first in storage order, markowitz.rs:105 */
                if (key < best) { best = key; bt = t; }
            }
        } else {
            for (int t = lane; t < nr; t += 32) {
                if (!bit_test(cb, t)) continue;
                const size_t off = (size_t)t * KD + c;
                const double x = fabs(dv[off]);
                if (x == 0.0 || x < tol) continue;
                const u64 mc = (u64)((nz1 - 1) * (i64)(d.rnz[t] - 1));
                const u64 key = (mc << 32) | (u64)dkey[off].c;
                if (key < best) { best = key; bt = t; }
            }
        }
        const u64 wb = warp_min64(best);
        const unsigned own = __ballot_sync(FULLMASK, best == wb && wb != KEY_INF);
        int tt = -1;
        if (own) tt = __shfl_sync(FULLMASK, bt, __ffs((int)own) - 1);
        if (lane == 0) {
            S.cand_mc[cc] = wb == KEY_INF ? -1 : (i64)(wb >> 32);
            S.cand_row[cc] = tt;
        }
    }
    bsync<NT>();
    if (tid == 0) {
        i64 mc64 = (i64)M.m * (i64)M.m;
        int bt = -1, bc = -1, bcc = -1;
        for (int cc = 0; cc < ncand; cc++)
            if (S.cand_mc[cc] >= 0 && S.cand_mc[cc] < mc64) { mc64 = S.cand_mc[cc]; bt = S.cand_row[cc]; bc = S.cand_col[cc]; bcc = cc; }
        BLU_CHECK(S, bc >= 0);
        if (bc >= 0) { S.dpt = bt; S.dpc = bc; S.pivot_row = d.drow[bt]; S.pivot_col = d.dcol[bc]; S.dpcand = bcc < DENSE_STASH ? bcc : -1; }
        S.nsearch += ncand;
    }
    bsync<NT>();
}

/* ------------------------------------------------------------------ */
/* pivot_any (pivot.rs:114-458) / pivot_small (pivot.rs:460-833)       */
/* ------------------------------------------------------------------ */
template <int NT, bool RES, bool SMALL> __device__ __noinline__ void dense_pivot(Shm &S) {
    Mat &M = S.M;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int KD = S.kd, KW = S.kw, nr = S.nrs, nc = S.ncs;
    DENSE_VIEW(RES);
    const int tp = S.dpt, cp = S.dpc, rank = S.rank;
    const double droptol = M.prm.droptol, abstol = M.prm.abstol;
    const unsigned *cbp = cbm + cp * KW, *rbp = rbm + tp * KW;
    const int n = d.cnz[cp], k = d.rnz[tp];
    const int cnz1 = n - 1, rnz1 = k - 1;

    i64 tq = clock64();
    /* the keys of the pivot row are one contiguous segment of kd * 4 bytes in HBM/L2: one thread hands the copy
     * to the bulk-copy engine (cp.async.bulk + mbarrier) and the block gathers the pivot column meanwhile */
    const unsigned phase = S.mbar_phase;
    if (tid == 0) bulk_copy_g2s(d.rowk, dkey + (size_t)tp * KD, (unsigned)(KD * sizeof(BluKey2)), &S.mbar);
    /* 1. the pivot column and the pivot row with their storage-order keys */
    {
        const int sc = S.dpcand;
        for (int t = tid; t < nr; t += NT)
            if (bit_test(cbp, t)) {
                const int e = bits_rank(cbp, t);
                d.tmps[e] = (unsigned short)t;
                d.keyc[e] = sc >= 0 ? d.candk[sc * KD + t] : dkey[(size_t)t * KD + cp].c;
            }
    }
#ifdef BLU_EMU
    bsync<NT>();      /* (the emulator's copy ran on thread 0) */
#endif
    mbar_wait(&S.mbar, phase);
    for (int c = tid; c < nc; c += NT)
        if (bit_test(rbp, c)) { const int e = bits_rank(rbp, c); d.tmpr[e] = (unsigned short)c; d.keyr[e] = d.rowk[c].r; }
    for (int w = tid; w < KW; w += NT) {
        d.cmask[w] = cbp[w] & ~(w == (tp >> 5) ? 1u << (tp & 31) : 0u);
        d.rfull[w] = rbp[w];
        d.rmask[w] = rbp[w] & ~(w == (cp >> 5) ? 1u << (cp & 31) : 0u);
    }
    bsync<NT>();
    /* 2. storage order = ascending key (rank by counting) */
    for (int e = tid; e < n; e += NT) {
        const unsigned my = d.keyc[e];
        int r = 0;
        for (int f = 0; f < n; f++) r += d.keyc[f] < my;
        const int t = d.tmps[e];
        d.clist[r] = (unsigned short)t;
        if (t == tp) S.wc = r;
    }
    for (int e = tid; e < k; e += NT) {
        const unsigned my = d.keyr[e];
        int r = 0;
        for (int f = 0; f < k; f++) r += d.keyr[f] < my;
        const int c = d.tmpr[e];
        d.rlist[r] = (unsigned short)c;
        if (c == cp) S.wr = r;
    }
    bsync<NT>();
    if (tid == 0) {
        /* pivot to the front of its column and row, pivot.rs:142-154 */
        unsigned short x = d.clist[0]; d.clist[0] = d.clist[S.wc]; d.clist[S.wc] = x;
        x = d.rlist[0]; d.rlist[0] = d.rlist[S.wr]; d.rlist[S.wr] = x;
        S.flag_a = 0; S.flag_b = 0;
    }
    bsync<NT>();
    for (int p = tid; p < n; p += NT) d.cvalp[p] = dv[(size_t)d.clist[p] * KD + cp];
    for (int kk = tid; kk < k; kk += NT) {
        d.posr[d.rlist[kk]] = (unsigned short)kk;
        if (SMALL) d.sdrop[kk] = 0;      /* (keyc / keyr are free again: the lists are ranked) */
    }
    bsync<NT>();
    const double pivot = d.cvalp[0];
    const int ubase = S.uput, lbase = S.lput;
    const i64 cbase = S.cstamp, rbase = S.rstamp;
    const unsigned ekey = S.epoch << 8;

    /* 3. per column of the pivot row: the entries outside the pivot column (pivot.rs:231-262).  Their
     * maximum seeds colmax; the first of them in storage order takes the place of the pivot-row entry. */
    for (int kk = 1 + tid; kk <= rnz1; kk += NT) {
        const int c = d.rlist[kk];
        const unsigned *cb = cbm + c * KW;
        double cmx = 0.0; unsigned kmin = 0xffffffffu; int tmin = -1;
        {   /* a column whose only entry outside the pivot column is the pivot-row entry (the usual case once the
             * tail is full) has nothing to exchange and no maximum to seed: no key is needed */
            int others = 0;
            for (int w = 0; w < KW; w++) others += __popc(cb[w] & ~d.cmask[w]);
            if (others == 1) { d.scm[c] = 0; continue; }
        }
        const unsigned short kpr = dkey[(size_t)tp * KD + c].c;
        /* the keys live in HBM/L2: fetch them eight at a time so that one round trip serves eight entries */
        int w = 0; unsigned tb = cb[0] & ~d.cmask[0];
        for (;;) {
            int tt[8]; unsigned kq[8];
            #pragma unroll
            for (int u = 0; u < 8; u++) {
                while (tb == 0 && w + 1 < KW) { w++; tb = cb[w] & ~d.cmask[w]; }
                if (tb) { tt[u] = w * 32 + __ffs((int)tb) - 1; tb &= tb - 1; } else tt[u] = -1;
            }
            if (tt[0] < 0) break;
            #pragma unroll
            for (int u = 0; u < 8; u++) kq[u] = tt[u] >= 0 ? dkey[(size_t)tt[u] * KD + c].c : 0xffffffffu;
            #pragma unroll
            for (int u = 0; u < 8; u++) {
                if (tt[u] < 0) continue;
                if (kq[u] < kmin) { kmin = kq[u]; tmin = tt[u]; }
                if (tt[u] != tp) { const double ax = fabs(dv[(size_t)tt[u] * KD + c]); if (ax > cmx) cmx = ax; }
            }
            if (tt[7] < 0) break;
        }
        if (tmin < 0) { BLU_CHECK(S, 0); }
        else if (tmin != tp) dkey[(size_t)tmin * KD + c].c = kpr;
        d.scm[c] = (u64)__double_as_longlong(cmx);
    }
    bsync<NT>();
    if (tid == 0) { const i64 now = clock64(); S.t_phase[14] += now - tq; tq = now; }

    /* 4. the rank-1 update, pivot.rs:285-304 / 638-662: a warp owns 32 column slots (and every RS-th row
     * of the pivot column when there are more warps than column words) */
    {
        const int RS = NW >= KW ? NW / KW : 1;
        for (int u0 = wid; u0 < KW * RS; u0 += NW) {
            const int cw = u0 % KW, rs = u0 / KW;
            const unsigned rf = d.rfull[cw], rm = d.rmask[cw];
            if (rf == 0) continue;
            const int c = cw * 32 + lane;
            const bool inR = (rm >> lane) & 1u;
            double a = 0.0; unsigned rkv = 0; int kk = 0;
            if (inR) { kk = d.posr[c]; a = __ddiv_rn(dv[(size_t)tp * KD + c], pivot); rkv = ekey + (unsigned)kk; }
            double cmx = 0.0;
            u64 mydrop = 0;
            for (int p0 = 1 + rs; p0 <= cnz1; p0 += 4 * RS) {
                int tt[4]; double xv[4], cv[4];
                #pragma unroll
                for (int u = 0; u < 4; u++) {      /* the four loads are in flight together */
                    const int p = p0 + u * RS;
                    tt[u] = p <= cnz1 ? (int)d.clist[p] : -1;
                    cv[u] = p <= cnz1 ? d.cvalp[p] : 0.0;
                    xv[u] = (inR && tt[u] >= 0) ? dv[(size_t)tt[u] * KD + c] : 0.0;
                }
                #pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (tt[u] < 0) continue;       /* uniform: p does not depend on the lane */
                    const int p = p0 + u * RS;
                    const size_t off = (size_t)tt[u] * KD + c;
                    int keep = 0;
                    if (inR) {
                        const double x = __dsub_rn(xv[u], __dmul_rn(a, cv[u]));
                        const double ax = fabs(x);
                        keep = SMALL ? ax > droptol : 1;
                        if (keep) {
                            dv[off] = x;
                            BluKey2 kv; kv.c = (unsigned short)(ekey + (unsigned)p); kv.r = (unsigned short)rkv;
                            dkey[off] = kv;
                            if (ax > cmx) cmx = ax;
                        } else {
                            dv[off] = 0.0;
                            mydrop |= 1ull << (p - 1);
                        }
                    }
                    if (SMALL) {
                        const unsigned km = __ballot_sync(FULLMASK, keep);
                        if (lane == 0) { unsigned *wp = rbm + tt[u] * KW + cw; *wp = (*wp & ~rf) | km; }
                    }
                }
            }
            if (inR && cmx > 0.0) atomicMax((unsigned long long *)&d.scm[c], (unsigned long long)__double_as_longlong(cmx));
            if (SMALL && mydrop) atomicOr((unsigned long long *)&d.sdrop[kk], (unsigned long long)mydrop);
        }
    }
    if (!SMALL) {
        /* pivot_any keeps every updated entry: row i = (old row minus the pivot row's columns) + those columns */
        for (int q = tid; q < cnz1 * KW; q += NT) {
            const int p = 1 + q / KW, w = q % KW;
            unsigned *wp = rbm + (int)d.clist[p] * KW + w;
            *wp = (*wp & ~d.rfull[w]) | d.rmask[w];
        }
    }
    bsync<NT>();
    if (tid == 0) { const i64 now = clock64(); S.t_phase[15] += now - tq; tq = now; }

    /* 5. counts, Markowitz keys, U row (pivot.rs:306-328), L column (:403-415) */
    double acc = 0.0;
    for (int kk = 1 + tid; kk <= rnz1; kk += NT) {
        const int c = d.rlist[kk], j = d.dcol[c];
        unsigned *cb = cbm + c * KW;
        for (int w = 0; w < KW; w++) {
            unsigned word = cb[w] | d.cmask[w];
            if (w == (tp >> 5)) word &= ~(1u << (tp & 31));
            cb[w] = word;
        }
        if (SMALL) {
            u64 drop = d.sdrop[kk];
            while (drop) {
                const int b = __ffsll((long long)drop) - 1;
                drop &= drop - 1;
                const int t = d.clist[b + 1];
                cb[t >> 5] &= ~(1u << (t & 31));
            }
        }
        int cnt = 0;
        for (int w = 0; w < KW; w++) cnt += __popc(cb[w]);
        const int oldnz = d.cnz[c];
        d.cnz[c] = (unsigned short)cnt;
        const double cmx = __longlong_as_double((long long)d.scm[c]);
        M.colpiv[j] = cmx;
        d.skeyc[c] = mkckey(cnt, cbase + kk, cmx, abstol);
        const double xr = dv[(size_t)tp * KD + c];
        if (fabs(xr) > droptol) { M.u_idx[ubase + kk - 1] = j; M.u_val[ubase + kk - 1] = xr; }
        else { M.u_idx[ubase + kk - 1] = -2; M.u_val[ubase + kk - 1] = 0.0; S.flag_a = 1; }
        if (cmx == 0.0 || cmx < abstol) S.need_remove = 1;
        acc += 12.0 * (oldnz + cnt);
    }
    for (int p = 1 + tid; p <= cnz1; p += NT) {
        const int t = d.clist[p], i = d.drow[t];
        const unsigned *rb = rbm + t * KW;
        int cnt = 0;
        for (int w = 0; w < KW; w++) cnt += __popc(rb[w]);
        const int oldnz = d.rnz[t];
        d.rnz[t] = (unsigned short)cnt;
        M.rkey[i] = mkkey(cnt, rbase + p);
        const double x = __ddiv_rn(d.cvalp[p], pivot);
        if (fabs(x) > droptol) { M.l_idx[lbase + p - 1] = i; M.l_val[lbase + p - 1] = x; }
        else { M.l_idx[lbase + p - 1] = -2; M.l_val[lbase + p - 1] = 0.0; S.flag_b = 1; }
        acc += 4.0 * (oldnz + cnt);
    }
    acc = warp_sumd(acc);      /* (sums of small integers: exact in any order) */
    if (lane == 0 && acc != 0.0) atomicAdd(&S.elim_bytes, acc);
    bsync<NT>();

    /* 6. the pivot row and column leave the active submatrix */
    for (int w = tid; w < KW; w += NT) { cbm[cp * KW + w] = 0; rbm[tp * KW + w] = 0; }
    if (wid == 0) {
        int ln = cnz1, un = rnz1;
        if (S.flag_b) ln = warp_squeeze(M.l_idx, M.l_val, lbase, cnz1);
        if (S.flag_a) un = warp_squeeze(M.u_idx, M.u_val, ubase, rnz1);
        if (lane == 0) {
            M.l_idx[lbase + ln] = -1;
            finish_step(S, rank, lbase + ln + 1, ubase + un, pivot, cnz1 + 1, rnz1 + 1);
            S.cstamp = cbase + rnz1 + 1;
            S.rstamp = rbase + cnz1 + 1;
            S.epoch++;
            d.skeyc[cp] = KEY_INF; d.cnz[cp] = 0; d.rnz[tp] = 0;
            S.n_kind[5]++;
            S.n_kind[7] += clock64() - tq;
            S.mbar_phase = phase ^ 1u;
        }
    }
    bsync<NT>();
}

#endif
