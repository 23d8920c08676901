/* blu_factor_dense.cuh -- the dense tail of the bump factorization.
 *
 * The reference has no dense-tail switch (src/lu/factorize_bump.rs:12-49 runs the sparse pivot() to the
 * end), so this representation must reproduce what markowitz.rs:34-193 and pivot.rs:114-833 would do on
 * the line file, bit for bit.  Once the active submatrix has at most `kd` rows it is held as
 *
 *   dn_val[t*kd + c]   the value of entry (row slot t, column slot c), 0.0 when absent   (row-major)
 *   dn_key[t*kd + c]   .c = storage-order key of the entry inside its column (the column file of file.rs),
 *                      .r = storage-order key inside its row (the row file); only order matters
 *   dn_rbits, dn_cbits presence bitmaps by row and by column
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

static inline size_t blu_dense_smem_bytes(int kd) {
    return (size_t)kd * (3 * 8 + 4 * 4 + 7 * 2) + (size_t)(kd / 32) * 3 * 4;
}

__device__ __forceinline__ void dense_carve(Shm &S, unsigned char *dyn) {
    const int KD = S.kd, KW = S.kw;
    u64 *p8 = (u64 *)dyn;
    S.skeyc = p8; p8 += KD;
    S.scm = p8; p8 += KD;
    S.cvalp = (double *)p8; p8 += KD;
    int *p4 = (int *)p8;
    S.drow = p4; p4 += KD;
    S.dcol = p4; p4 += KD;
    S.keyc = (unsigned *)p4; p4 += KD;
    S.keyr = (unsigned *)p4; p4 += KD;
    S.cmask = (unsigned *)p4; p4 += KW;
    S.rmask = (unsigned *)p4; p4 += KW;
    S.rfull = (unsigned *)p4; p4 += KW;
    unsigned short *p2 = (unsigned short *)p4;
    S.clist = p2; p2 += KD;
    S.rlist = p2; p2 += KD;
    S.posr = p2; p2 += KD;
    S.rnz = p2; p2 += KD;
    S.cnz = p2; p2 += KD;
    S.tmps = p2; p2 += KD;
    S.tmpr = p2; p2 += KD;
}

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
template <int NT> __device__ void dense_enter(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int KD = S.kd, KW = S.kw;
    int *rowslot = M.iwork1, *colslot = M.iwork1 + m;
    if (tid == 0) dense_carve(S, S.dyn);
    bsync<NT>();
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
            if (ra && nr + exr < KD) S.drow[nr + exr] = i;
            if (ca && nc + exc < KD) S.dcol[nc + exc] = i;
        }
        nr += totr; nc += totc;
    }
    if (nr > KD || nc > KD) { if (tid == 0) BLU_CHECK(S, 0); bsync<NT>(); return; }
    for (size_t q = tid; q < (size_t)nr * KD; q += NT) M.dn_val[q] = 0.0;
    for (int q = tid; q < KD * KW; q += NT) { M.dn_rbits[q] = 0; M.dn_cbits[q] = 0; }
    for (int c = tid; c < KD; c += NT) { S.skeyc[c] = KEY_INF; S.cnz[c] = 0; S.rnz[c] = 0; }
    bsync<NT>();
    for (int c = wid; c < nc; c += NW) {
        const int j = S.dcol[c];
        const int b = M.lbeg[j], e = M.lend[j];
        for (int pos = b + lane; pos < e; pos += 32) {
            const int t = rowslot[M.w_idx[pos]];
            const size_t off = (size_t)t * KD + c;
            M.dn_val[off] = M.w_val[pos];
            M.dn_key[off].c = (unsigned)(pos - b);
            atomicOr(&M.dn_rbits[t * KW + (c >> 5)], 1u << (c & 31));
            atomicOr(&M.dn_cbits[c * KW + (t >> 5)], 1u << (t & 31));
        }
        if (lane == 0) { S.cnz[c] = (unsigned short)(e - b); S.skeyc[c] = M.ckey[j]; }
    }
    for (int t = wid; t < nr; t += NW) {
        const int i = S.drow[t];
        const int b = M.lbeg[m + i], e = M.lend[m + i];
        for (int pos = b + lane; pos < e; pos += 32) {
            const int c = colslot[M.w_idx[pos]];
            M.dn_key[(size_t)t * KD + c].r = (unsigned)(pos - b);
        }
        if (lane == 0) S.rnz[t] = (unsigned short)(e - b);
    }
    if (tid == 0) {
        S.dense = 1; S.nrs = nr; S.ncs = nc;
        S.ekc = (unsigned)KD; S.ekr = (unsigned)KD;       /* above every position key handed out here */
        S.dense_entries++;
        S.n_kind[6]++;
    }
    bsync<NT>();
}

/* ------------------------------------------------------------------ */
/* dense arrays -> line file (key order == storage order)              */
/* ------------------------------------------------------------------ */
template <int NT> __device__ void dense_exit(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int KD = S.kd, KW = S.kw, nr = S.nrs, nc = S.ncs;
    const int base = S.w_half * M.w_mem;
    int put = 0, nact = 0;
    for (int b0 = 0; b0 < nc; b0 += NT) {
        const int c = b0 + tid;
        const int alive = c < nc && S.skeyc[c] != KEY_INF;
        const int n = alive ? S.cnz[c] : 0;
        const int sz = n > 0 ? n + slack_of(M.prm, n) : 0;
        int tot, ex = block_excl_scan<NT>(sz, &tot, S.iscr);
        int tota, exa = block_excl_scan<NT>(alive, &tota, S.iscr);
        if (alive) {
            const int j = S.dcol[c], b = base + put + ex;
            M.lbeg[j] = b; M.lend[j] = b + n; M.lcap[j] = b + sz;
            M.ckey[j] = S.skeyc[c];
            M.acols[nact + exa] = j;
        }
        put += tot; nact += tota;
    }
    for (int b0 = 0; b0 < nr; b0 += NT) {
        const int t = b0 + tid;
        const int alive = t < nr && M.rkey[S.drow[t]] != KEY_INF;
        const int n = alive ? S.rnz[t] : 0;
        const int sz = alive ? n + slack_of(M.prm, n) : 0;
        int tot, ex = block_excl_scan<NT>(sz, &tot, S.iscr);
        if (alive) {
            const int i = S.drow[t], b = base + put + ex;
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
        if (S.skeyc[c] == KEY_INF) continue;
        const int b = M.lbeg[S.dcol[c]];
        const unsigned *cb = M.dn_cbits + c * KW;
        for (int t = lane; t < nr; t += 32) {
            if (!bit_test(cb, t)) continue;
            const size_t off = (size_t)t * KD + c;
            const unsigned my = M.dn_key[off].c;
            int r = 0;
            for (int w = 0; w < KW; w++) {
                unsigned word = cb[w];
                while (word) { const int t2 = w * 32 + __ffs((int)word) - 1; word &= word - 1; r += M.dn_key[(size_t)t2 * KD + c].c < my; }
            }
            M.w_idx[b + r] = S.drow[t]; M.w_val[b + r] = M.dn_val[off];
        }
    }
    for (int t = wid; t < nr; t += NW) {
        const int i = S.drow[t];
        if (M.rkey[i] == KEY_INF) continue;
        const int b = M.lbeg[m + i];
        const unsigned *rb = M.dn_rbits + t * KW;
        for (int c = lane; c < nc; c += 32) {
            if (!bit_test(rb, c)) continue;
            const unsigned my = M.dn_key[(size_t)t * KD + c].r;
            int r = 0;
            for (int w = 0; w < KW; w++) {
                unsigned word = rb[w];
                while (word) { const int c2 = w * 32 + __ffs((int)word) - 1; word &= word - 1; r += M.dn_key[(size_t)t * KD + c2].r < my; }
            }
            M.w_idx[b + r] = S.dcol[c];
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
template <int NT> __device__ void dense_search(Shm &S) {
    Mat &M = S.M;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int KD = S.kd, KW = S.kw, nr = S.nrs, nc = S.ncs;
    int maxsearch = M.prm.maxsearch;
    if (maxsearch < 1) maxsearch = 1;
    if (maxsearch > MAXCAND) maxsearch = MAXCAND;
    int ncand = 0;
    u64 prev = 0; int have_prev = 0;
    while (ncand < maxsearch) {
        u64 k0 = KEY_INF, k1 = KEY_INF, k2 = KEY_INF;
        int j0 = -1, j1 = -1, j2 = -1;
        for (int c = tid; c < nc; c += NT) {
            const u64 k = S.skeyc[c];
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
            const u64 best = block_min64<NT>(k0, S.kscr);
            if (best == KEY_INF) break;
            if (k0 == best) {
                S.cand_col[ncand] = j0;
                k0 = k1; j0 = j1; k1 = k2; j1 = j2; k2 = KEY_INF; j2 = -1;
            }
            prev = best; have_prev = 1;
            ncand++; got++;
        }
        bsync<NT>();
        if (got < 3) break;
    }
    if (ncand == 0) { if (tid == 0) { BLU_CHECK(S, 0); } bsync<NT>(); return; }
    if (key_cnt(S.skeyc[S.cand_col[0]]) == 0) {      /* markowitz.rs:73-78 */
        if (tid == 0) { S.dpc = S.cand_col[0]; S.pivot_col = S.dcol[S.cand_col[0]]; S.pivot_row = -1; }
        bsync<NT>();
        return;
    }
    const double abstol = M.prm.abstol, reltol = M.prm.reltol;
    for (int cc = wid; cc < ncand; cc += NW) {
        const int c = S.cand_col[cc];
        const i64 nz1 = S.cnz[c];
        const double cmx = M.colpiv[S.dcol[c]];
        const double tol = fmax(abstol, reltol * cmx);
        const unsigned *cb = M.dn_cbits + c * KW;
        u64 best = KEY_INF; int bt = -1;
        for (int t = lane; t < nr; t += 32) {
            if (!bit_test(cb, t)) continue;
            const size_t off = (size_t)t * KD + c;
            const double x = fabs(M.dn_val[off]);
            if (x == 0.0 || x < tol) continue;
            const u64 mc = (u64)((nz1 - 1) * (i64)(S.rnz[t] - 1));
            const u64 key = (mc << 32) | (u64)M.dn_key[off].c;       /* ties: first in storage order, markowitz.rs:105 */
            if (key < best) { best = key; bt = t; }
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
        int bt = -1, bc = -1;
        for (int cc = 0; cc < ncand; cc++)
            if (S.cand_mc[cc] >= 0 && S.cand_mc[cc] < mc64) { mc64 = S.cand_mc[cc]; bt = S.cand_row[cc]; bc = S.cand_col[cc]; }
        BLU_CHECK(S, bc >= 0);
        if (bc >= 0) { S.dpt = bt; S.dpc = bc; S.pivot_row = S.drow[bt]; S.pivot_col = S.dcol[bc]; }
        S.nsearch += ncand;
    }
    bsync<NT>();
}

/* ------------------------------------------------------------------ */
/* pivot_any (pivot.rs:114-458) / pivot_small (pivot.rs:460-833)       */
/* ------------------------------------------------------------------ */
template <int NT> __device__ void dense_pivot(Shm &S, const bool small) {
    Mat &M = S.M;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int KD = S.kd, KW = S.kw, nr = S.nrs, nc = S.ncs;
    const int tp = S.dpt, cp = S.dpc, rank = S.rank;
    const double droptol = M.prm.droptol, abstol = M.prm.abstol;
    const unsigned *cbp = M.dn_cbits + cp * KW, *rbp = M.dn_rbits + tp * KW;
    const int n = S.cnz[cp], k = S.rnz[tp];
    const int cnz1 = n - 1, rnz1 = k - 1;

    /* 1. the pivot column and the pivot row with their storage-order keys */
    for (int t = tid; t < nr; t += NT)
        if (bit_test(cbp, t)) { const int e = bits_rank(cbp, t); S.tmps[e] = (unsigned short)t; S.keyc[e] = M.dn_key[(size_t)t * KD + cp].c; }
    for (int c = tid; c < nc; c += NT)
        if (bit_test(rbp, c)) { const int e = bits_rank(rbp, c); S.tmpr[e] = (unsigned short)c; S.keyr[e] = M.dn_key[(size_t)tp * KD + c].r; }
    for (int w = tid; w < KW; w += NT) {
        S.cmask[w] = cbp[w] & ~(w == (tp >> 5) ? 1u << (tp & 31) : 0u);
        S.rfull[w] = rbp[w];
        S.rmask[w] = rbp[w] & ~(w == (cp >> 5) ? 1u << (cp & 31) : 0u);
    }
    bsync<NT>();
    /* 2. storage order = ascending key (rank by counting) */
    for (int e = tid; e < n; e += NT) {
        const unsigned my = S.keyc[e];
        int r = 0;
        for (int f = 0; f < n; f++) r += S.keyc[f] < my;
        const int t = S.tmps[e];
        S.clist[r] = (unsigned short)t;
        if (t == tp) S.wc = r;
    }
    for (int e = tid; e < k; e += NT) {
        const unsigned my = S.keyr[e];
        int r = 0;
        for (int f = 0; f < k; f++) r += S.keyr[f] < my;
        const int c = S.tmpr[e];
        S.rlist[r] = (unsigned short)c;
        if (c == cp) S.wr = r;
    }
    bsync<NT>();
    if (tid == 0) {
        /* pivot to the front of its column and row, pivot.rs:142-154 */
        unsigned short x = S.clist[0]; S.clist[0] = S.clist[S.wc]; S.clist[S.wc] = x;
        x = S.rlist[0]; S.rlist[0] = S.rlist[S.wr]; S.rlist[S.wr] = x;
        S.flag_a = 0; S.flag_b = 0;
    }
    bsync<NT>();
    for (int p = tid; p < n; p += NT) S.cvalp[p] = M.dn_val[(size_t)S.clist[p] * KD + cp];
    for (int kk = tid; kk < k; kk += NT) {
        const int c = S.rlist[kk];
        S.posr[c] = (unsigned short)kk;
        S.scm[c] = 0;
        if (small && kk > 0) M.cancelled[kk - 1] = 0;
    }
    bsync<NT>();
    const double pivot = S.cvalp[0];
    const int ubase = M.u_begin[rank], lbase = M.l_begin_p[rank];
    const i64 cbase = S.cstamp, rbase = S.rstamp;
    const unsigned ekc = S.ekc, ekr = S.ekr;

    /* 3. per column of the pivot row: the entries outside the pivot column (pivot.rs:231-262).  Their
     * maximum seeds colmax; the first of them in storage order takes the place of the pivot-row entry. */
    for (int kk = 1 + tid; kk <= rnz1; kk += NT) {
        const int c = S.rlist[kk];
        const unsigned *cb = M.dn_cbits + c * KW;
        double cmx = 0.0; unsigned kmin = 0xffffffffu; int tmin = -1;
        for (int w = 0; w < KW; w++) {
            unsigned tb = cb[w] & ~S.cmask[w];
            while (tb) {
                const int t = w * 32 + __ffs((int)tb) - 1;
                tb &= tb - 1;
                const size_t off = (size_t)t * KD + c;
                const unsigned key = M.dn_key[off].c;
                if (key < kmin) { kmin = key; tmin = t; }
                if (t != tp) { const double ax = fabs(M.dn_val[off]); if (ax > cmx) cmx = ax; }
            }
        }
        if (tmin < 0) { BLU_CHECK(S, 0); }
        else if (tmin != tp) M.dn_key[(size_t)tmin * KD + c].c = M.dn_key[(size_t)tp * KD + c].c;
        S.scm[c] = (u64)__double_as_longlong(cmx);
    }
    bsync<NT>();

    /* 4. the rank-1 update, pivot.rs:285-304 / 638-662: a warp owns 32 column slots (and every RS-th row
     * of the pivot column when there are more warps than column words) */
    {
        const int RS = NW >= KW ? NW / KW : 1;
        for (int u = wid; u < KW * RS; u += NW) {
            const int cw = u % KW, rs = u / KW;
            const unsigned rf = S.rfull[cw], rm = S.rmask[cw];
            if (rf == 0) continue;
            const int c = cw * 32 + lane;
            const bool inR = (rm >> lane) & 1u;
            double a = 0.0; unsigned rkv = 0; int kk = 0;
            if (inR) { kk = S.posr[c]; a = __ddiv_rn(M.dn_val[(size_t)tp * KD + c], pivot); rkv = ekr + (unsigned)kk; }
            double cmx = 0.0;
            for (int p = 1 + rs; p <= cnz1; p += RS) {
                const int t = S.clist[p];
                const size_t off = (size_t)t * KD + c;
                int keep = 0;
                if (inR) {
                    const double x = __dsub_rn(M.dn_val[off], __dmul_rn(a, S.cvalp[p]));
                    const double ax = fabs(x);
                    keep = small ? ax > droptol : 1;
                    if (keep) {
                        M.dn_val[off] = x;
                        BluKey2 kv; kv.c = ekc + (unsigned)p; kv.r = rkv;
                        M.dn_key[off] = kv;
                        if (ax > cmx) cmx = ax;
                    } else {
                        M.dn_val[off] = 0.0;
                        atomicOr((unsigned long long *)&M.cancelled[kk - 1], 1ull << (p - 1));
                    }
                }
                const unsigned km = small ? __ballot_sync(FULLMASK, keep) : rm;
                if (lane == 0) { unsigned *wp = M.dn_rbits + t * KW + cw; *wp = (*wp & ~rf) | km; }
            }
            if (inR && cmx > 0.0) atomicMax((unsigned long long *)&S.scm[c], (unsigned long long)__double_as_longlong(cmx));
        }
    }
    bsync<NT>();

    /* 5. counts, Markowitz keys, U row (pivot.rs:306-328), L column (:403-415) */
    double acc = 0.0;
    for (int kk = 1 + tid; kk <= rnz1; kk += NT) {
        const int c = S.rlist[kk], j = S.dcol[c];
        unsigned *cb = M.dn_cbits + c * KW;
        for (int w = 0; w < KW; w++) {
            unsigned word = cb[w] | S.cmask[w];
            if (w == (tp >> 5)) word &= ~(1u << (tp & 31));
            cb[w] = word;
        }
        if (small) {
            u64 drop = M.cancelled[kk - 1];
            while (drop) {
                const int b = __ffsll((long long)drop) - 1;
                drop &= drop - 1;
                const int t = S.clist[b + 1];
                cb[t >> 5] &= ~(1u << (t & 31));
            }
        }
        int cnt = 0;
        for (int w = 0; w < KW; w++) cnt += __popc(cb[w]);
        const int oldnz = S.cnz[c];
        S.cnz[c] = (unsigned short)cnt;
        const double cmx = __longlong_as_double((long long)S.scm[c]);
        M.colpiv[j] = cmx;
        S.skeyc[c] = mkckey(cnt, cbase + kk, cmx, abstol);
        const double xr = M.dn_val[(size_t)tp * KD + c];
        if (fabs(xr) > droptol) { M.u_idx[ubase + kk - 1] = j; M.u_val[ubase + kk - 1] = xr; }
        else { M.u_idx[ubase + kk - 1] = -2; M.u_val[ubase + kk - 1] = 0.0; S.flag_a = 1; }
        if (cmx == 0.0 || cmx < abstol) S.need_remove = 1;
        acc += 12.0 * (oldnz + cnt);
    }
    for (int p = 1 + tid; p <= cnz1; p += NT) {
        const int t = S.clist[p], i = S.drow[t];
        const unsigned *rb = M.dn_rbits + t * KW;
        int cnt = 0;
        for (int w = 0; w < KW; w++) cnt += __popc(rb[w]);
        const int oldnz = S.rnz[t];
        S.rnz[t] = (unsigned short)cnt;
        M.rkey[i] = mkkey(cnt, rbase + p);
        const double x = __ddiv_rn(S.cvalp[p], pivot);
        if (fabs(x) > droptol) { M.l_idx[lbase + p - 1] = i; M.l_val[lbase + p - 1] = x; }
        else { M.l_idx[lbase + p - 1] = -2; M.l_val[lbase + p - 1] = 0.0; S.flag_b = 1; }
        acc += 4.0 * (oldnz + cnt);
    }
    if (acc != 0.0) atomicAdd(&S.elim_bytes, acc);
    bsync<NT>();

    /* 6. the pivot row and column leave the active submatrix */
    for (int w = tid; w < KW; w += NT) { M.dn_cbits[cp * KW + w] = 0; M.dn_rbits[tp * KW + w] = 0; }
    if (wid == 0) {
        int ln = cnz1, un = rnz1;
        if (S.flag_b) ln = warp_squeeze(M.l_idx, M.l_val, lbase, cnz1);
        if (S.flag_a) un = warp_squeeze(M.u_idx, M.u_val, ubase, rnz1);
        if (lane == 0) {
            M.l_idx[lbase + ln] = -1;
            finish_step(S, rank, lbase + ln + 1, ubase + un, pivot, cnz1 + 1, rnz1 + 1);
            S.cstamp = cbase + rnz1 + 1;
            S.rstamp = rbase + cnz1 + 1;
            S.ekc = ekc + (unsigned)cnz1 + 1u; S.ekr = ekr + (unsigned)rnz1 + 1u;
            S.skeyc[cp] = KEY_INF; S.cnz[cp] = 0; S.rnz[tp] = 0;
            S.n_kind[5]++;
        }
    }
    bsync<NT>();
}

#endif
