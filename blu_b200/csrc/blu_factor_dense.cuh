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
 * RES = 1: dv, rb and cb live in shared memory (one CTA per SM, kd <= 160 on B200: 200 KB of values);
 * the keys stay in HBM/L2 and are written fire-and-forget by the update and read O(kd) times per step.
 * RES = 0: dv, rb, cb are the HBM arrays dn_val, dn_rbits, dn_cbits (any kd).
 * RES = 2: the bitmaps in shared memory, the values in HBM/L2: the first stage of a two-stage tail, which starts at
 * order kd_big (256) in a launch with one CTA per SM and moves the active submatrix into shared memory
 * (dense_restage) when it has shrunk to kd_small.
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
#ifndef RING_LAG
#define RING_LAG 2           /* rows between a buffer's store and its refill (stage-1 row ring) */
#endif

/* Pointers into the dynamic shared memory, computed locally from its base in every function so that the
 * compiler knows their address space (LDS/STS instead of generic loads, and no aliasing with the global
 * stores of the update loop). */
struct DenseSm {
    u64 *skeyc;            /* Markowitz key of column slot c (count<<40 | stamp), KEY_INF when gone */
    unsigned *skey32;      /* the same order in 32 bits for the candidate scan: count<<23 | (stamp - dstamp_base), or the rank of the
                            * stamp among the columns present at dense_enter; 0xffffffff when gone or hidden (KEY_PARK) */
    u64 *scm;              /* colmax of column slot c as the bits of a non-negative double */
    double *cvalp;         /* values of the pivot column in its storage order */
    double *arow;          /* pivot row divided by the pivot, by column slot (pivot.rs:285: the multiplier of each column) */
    int *drow, *dcol;      /* slot -> row / column index */
    unsigned *keyc, *keyr; /* sort keys of the pivot column / row */
    u64 *sdrop;            /* pivot_small: cancellation mask per column of the pivot row (overlays keyc|keyr) */
    unsigned *candk;       /* key stash of the search candidates */
    BluKey2 *rowk;         /* the pivot row's keys, fetched by one bulk copy (16-byte aligned: kd is a multiple of 32) */
    unsigned *cmask, *rmask, *rfull;
    unsigned short *clist, *rlist, *posr, *rnz, *cnz, *tmps, *tmpr;
    unsigned *kminp;       /* per column of the pivot row: (key << 8 | row slot) of its first entry outside the pivot column and row (overlays tmps|tmpr) */
    unsigned *rb_s, *cb_s; double *dv_s;   /* resident bitmaps and values (RES only) */
};
__device__ __forceinline__ void dense_view(DenseSm &d, unsigned char *dyn, int KD, int KW) {
    u64 *p8 = (u64 *)dyn;
    d.skeyc = p8; p8 += KD;
    d.scm = p8; p8 += KD;
    d.cvalp = (double *)p8; p8 += KD;
    d.arow = (double *)p8; p8 += KD;
    int *p4 = (int *)p8;
    d.drow = p4; p4 += KD;
    d.dcol = p4; p4 += KD;
    d.keyc = (unsigned *)p4; d.sdrop = (u64 *)p4; p4 += KD;
    d.keyr = (unsigned *)p4; p4 += KD;
    d.rowk = (BluKey2 *)p4; p4 += KD;
    d.candk = (unsigned *)p4; p4 += DENSE_STASH * KD;
    d.skey32 = (unsigned *)p4; p4 += KD;
    d.cmask = (unsigned *)p4; p4 += KW;
    d.rmask = (unsigned *)p4; p4 += KW;
    d.rfull = (unsigned *)p4; p4 += KW;
    unsigned short *p2 = (unsigned short *)p4;
    d.clist = p2; p2 += KD;
    d.rlist = p2; p2 += KD;
    d.posr = p2; p2 += KD;
    d.rnz = p2; p2 += KD;
    d.cnz = p2; p2 += KD;
    d.tmps = p2; d.kminp = (unsigned *)p2; p2 += KD;
    d.tmpr = p2; p2 += KD;
    unsigned *b = (unsigned *)(dyn + blu_dense_smem_bytes(KD));
    d.rb_s = b; d.cb_s = b + KD * KW;
    d.dv_s = (double *)(b + 2 * KD * KW);
}
#define DENSE_VIEW(RES) \
    BLU_DYN_SMEM(dyn_); DenseSm d; dense_view(d, dyn_, S.kd, S.kw); \
    double *const dv = RES == 1 ? d.dv_s : S.M.dn_val; \
    unsigned *const rbm = RES ? d.rb_s : S.M.dn_rbits; \
    unsigned *const cbm = RES ? d.cb_s : S.M.dn_cbits; \
    BluKey2 *const dkey = S.M.dn_key
/* a value of the dense array (past L1 in the first stage of a two-stage tail, whose rows bulk stores rewrite) */
#define DV(off) (RES == 2 ? ld_l2(dv + (off)) : dv[off])

/* byte offset of the row ring behind the per-step arrays and the bitmaps of order kd (128-byte aligned) */
__device__ __forceinline__ size_t dense_ring_offset(int kd) {
    return ((blu_dense_smem_bytes(kd) + (size_t)2 * kd * (kd / 32) * 4 + 127) & ~(size_t)127) + RING_BARS * sizeof(u64);
}
/* the ring's completion barriers live right in front of it; they exist from dense_enter to the end of the stage */
__device__ __forceinline__ u64 *dense_ring_bars(unsigned char *dyn, int kd) { return (u64 *)(dyn + dense_ring_offset(kd)) - RING_BARS; }
template <int NT> __device__ __forceinline__ void dense_ring_retire(Shm &S) {
    if (S.ring_nbuf) {
        BLU_DYN_SMEM(dynq_);
        u64 *bars = dense_ring_bars(dynq_, S.kd);
        for (int q = threadIdx.x; q < (NT / 32) * S.ring_nbuf; q += NT) mbar_inval(&bars[q]);
    }
    bsync<NT>();
    if (threadIdx.x == 0) S.ring_nbuf = 0;
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
template <int NT, int RES> __device__ __noinline__ void dense_enter(Shm &S) {
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
    {   /* no entry: key 0xffff/0xffff (dense_step ranks the keys of a row without looking at the bitmaps) */
        unsigned *k32 = (unsigned *)dkey;
        for (size_t q = tid; q < (size_t)KD * KD; q += NT) k32[q] = 0xffffffffu;
    }
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
        S.ring_nbuf = 0;
        if (RES == 2) {      /* what the launch has beyond this stage's arrays becomes the row ring of the update */
            const size_t base = dense_ring_offset(KD);
            const size_t have = blu_dense_smem_bytes_resident(S.kd_small);
            const size_t level = (size_t)NW * KD * sizeof(double);
            int nb = have > base ? (int)((have - base) / level) : 0;
            if (nb > RING_BARS / NW) nb = RING_BARS / NW;
            if (nb > 8) nb = 8;
            S.ring_nbuf = nb > RING_LAG ? nb : 0;
        }
        S.dense = 1; S.nrs = nr; S.ncs = nc;
        S.epoch = 1;          /* the position keys handed out here are epoch 0 */
        S.dense_entries++;
        S.n_kind[6]++;
    }
    bsync<NT>();
    /* 32-bit search keys: the stamps of the columns present now are replaced by their ranks (< nc); the stamps handed
     * out from here on count up from nc */
    for (int c = tid; c < KD; c += NT) {
        const u64 key = d.skeyc[c];
        unsigned k32 = 0xffffffffu;
        if (key < KEY_PARK) {
            const u64 smask = ((u64)1 << STAMP_BITS) - 1;
            const u64 st = key & smask;
            unsigned r = 0;
            for (int c2 = 0; c2 < nc; c2++) { const u64 k2 = d.skeyc[c2]; r += k2 != KEY_INF && (k2 & smask) < st; }
            k32 = ((unsigned)key_cnt(key) << 23) | r;
        }
        d.skey32[c] = k32;
    }
    if (tid == 0) S.dstamp_base = S.cstamp - nc;
    bsync<NT>();
    if (RES == 2 && S.ring_nbuf) {
        u64 *bars = dense_ring_bars(dyn_, KD);
        for (int q = tid; q < NW * S.ring_nbuf; q += NT) mbar_init(&bars[q], 1);
        if (tid < 32) S.ring_phase[tid] = 0;
        bsync<NT>();
    }
}

/* ------------------------------------------------------------------ */
/* dense arrays -> line file (key order == storage order)              */
/* ------------------------------------------------------------------ */
template <int NT, int RES> __device__ __noinline__ void dense_exit(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int KD = S.kd, KW = S.kw, nr = S.nrs, nc = S.ncs;
    DENSE_VIEW(RES);
    if (RES == 2) dense_ring_retire<NT>(S);
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
            M.w_idx[b + r] = d.drow[t]; M.w_val[b + r] = DV(off);
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
/* the dense pivot loop: markowitz.rs:34-123 + pivot_any / pivot_small   */
/* ------------------------------------------------------------------ */
/* One elimination step on the dense arrays is five block barriers long:
 *   A  warp 0 picks the candidate columns (the first `maxsearch` live columns in (count, stamp) order) while the
 *      finisher warp still writes the previous step's L/U pointers                                   -- barrier 1
 *   B  one warp per candidate column evaluates it (markowitz.rs:82-123) and leaves the column's storage-order keys
 *      in shared memory; the warp that finishes last makes the choice, checks the room in L and U
 *      (pivot.rs:69-81) and starts the bulk copy of the pivot row's keys                             -- barrier 2
 *   C  one thread per entry of the pivot column and of the pivot row ranks its key by counting: this is the
 *      storage order, with the pivot moved to the front (pivot.rs:142-154)                           -- barrier 3
 *   D  one thread per column of the pivot row looks at the entries outside the pivot column (pivot.rs:231-262);
 *   E  the rank-1 update (pivot.rs:285-304 / 638-662), a warp per 32 column slots                    -- barrier 4
 *   F  counts, Markowitz keys, the U row and the L column (pivot.rs:306-328, 403-415)                -- barrier 5
 * Invariant used by C: an entry that is not in the active submatrix carries the key 0xffff/0xffff in dn_key
 * (dense_enter fills the array with it, a dropped entry and the entries of a pivot column get it in E), so that
 * the keys of a row can be ranked as 32-bit words (.r in the upper half) without looking at the bitmaps. */
#define DRUN_DONE 0          /* nothing left to do here: all pivots found, an error, or Reallocate (S.status) */
#define DRUN_SPARSE_PIVOT 1  /* S.pivot_row/col chosen, but the step belongs to the sparse code (singleton row/column,
                              * doubleton column, epoch counter exhausted): dense_exit, then pivot */
#define DRUN_REMOVE 2        /* a step is done and a column has to be emptied (pivot.rs:96-106): dense_exit + post_remove_cols */
#define DRUN_RESTAGE 3       /* the first stage has shrunk to kd_small rows: dense_restage, then on in shared memory */

template <int NT, int RES, bool SMALL> __device__ __forceinline__ void dense_step(Shm &S, const int rank) {
    Mat &M = S.M;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int KD = S.kd, KW = S.kw, nr = S.nrs, nc = S.ncs;
    DENSE_VIEW(RES);
    const int tp = S.dpt, cp = S.dpc;
    const double droptol = M.prm.droptol, abstol = M.prm.abstol;
    const unsigned *cbp = cbm + cp * KW, *rbp = rbm + tp * KW;
    const int n = d.cnz[cp], k = d.rnz[tp];
    const int cnz1 = n - 1, rnz1 = k - 1;
    const unsigned phase = S.mbar_phase;
    const unsigned *rowk32 = (const unsigned *)d.rowk;      /* (r << 16) | c of the pivot row's entries */
    i64 tq = clock64();

    /* C. storage order of the pivot column and the pivot row = ascending key, pivot first */
    for (int w = NT - 1 - tid; w < KW; w += NT) {
        d.cmask[w] = cbp[w] & ~(w == (tp >> 5) ? 1u << (tp & 31) : 0u);
        d.rfull[w] = rbp[w];
        d.rmask[w] = rbp[w] & ~(w == (cp >> 5) ? 1u << (cp & 31) : 0u);
    }
    if (tid == NT - 1) { S.flag_a = 0; S.flag_b = 0; }
    {
        const unsigned *stash = d.candk + S.dpcand * KD;      /* the column's keys, 0xffffffff where there is no entry */
        const uint4 *s4 = (const uint4 *)stash;
        const int n4 = (nr + 3) >> 2;
        const unsigned pk = stash[tp];
        for (int t0 = wid * 32; t0 < nr; t0 += NT) {      /* (whole warps: the rank of the pivot is a warp's work) */
            const int t = t0 + lane;
            const unsigned my = t < nr ? stash[t] : 0xffffffffu;
            int r = 0;
            if (my != 0xffffffffu)
                for (int f = 0; f < n4; f++) {
                    const uint4 q = s4[f];
                    r += (q.x < my) + (q.y < my) + (q.z < my) + (q.w < my);
                }
            /* the first entry and the pivot exchange places: only the first entry needs the pivot's rank */
            int wc = 0;
            if (__ballot_sync(FULLMASK, my != 0xffffffffu && r == 0 && t != tp)) {
                for (int f = lane; f < 4 * n4; f += 32) wc += stash[f] < pk;
                wc = warp_sum(wc);
            }
            if (my != 0xffffffffu) {
                const int p = t == tp ? 0 : (r == 0 ? wc : r);
                d.clist[p] = (unsigned short)t;
                d.cvalp[p] = DV((size_t)t * KD + cp);
            }
        }
    }
    mbar_wait(&S.mbar, phase);
    {
        const uint4 *s4 = (const uint4 *)rowk32;
        const int n4 = (nc + 3) >> 2;
        const unsigned pk = rowk32[cp];
        for (int it0 = wid * 32; it0 < KD + nc; it0 += NT) {
            if (it0 < KD) continue;      /* (the warps that ranked the column are busy) */
            const int c = it0 - KD + lane;
            const bool on = c < nc && bit_test(rbp, c);
            /* the multiplier of the column, once per step (the loads are in flight during the ranking) */
            const double xrow = on ? DV((size_t)tp * KD + c) : 0.0, xpiv = DV((size_t)tp * KD + cp);
            const unsigned my = c < nc ? rowk32[c] : 0xffffffffu;
#ifdef BLU_EMU
            if (c < nc) BLU_CHECK(S, ((my >> 16) == 0xffffu) == !on);
#endif
            int r = 0;
            if (on)
                for (int f = 0; f < n4; f++) {
                    const uint4 q = s4[f];
                    r += (q.x < my) + (q.y < my) + (q.z < my) + (q.w < my);
                }
            int wr = 0;
            if (__ballot_sync(FULLMASK, on && r == 0 && c != cp)) {
                for (int f = lane; f < 4 * n4; f += 32) wr += rowk32[f] < pk;
                wr = warp_sum(wr);
            }
            if (on) {
                const int p = c == cp ? 0 : (r == 0 ? wr : r);
                d.rlist[p] = (unsigned short)c;
                d.posr[c] = (unsigned short)p;
                if (c != cp) { d.scm[c] = 0; d.kminp[c] = 0xffffffffu; d.arow[c] = __ddiv_rn(xrow, xpiv); }
                if (SMALL) d.sdrop[p] = 0;
            }
        }
    }
    bsync<NT>();
    const double pivot = d.cvalp[0];
    const int ubase = S.uput, lbase = S.lput;
    const i64 cbase = S.cstamp, rbase = S.rstamp, sbase = S.dstamp_base;
    const unsigned ekey = S.epoch << 8;
    if (tid == 0) { const i64 now = clock64(); S.t_phase[14] += now - tq; tq = now; }

    /* D. per column of the pivot row: the entries outside the pivot column and the pivot row (pivot.rs:231-262).
     * Their maximum seeds colmax, and the first of them in storage order -- if it comes before the pivot-row entry
     * -- takes that entry's place (its key, written in F).  G threads share a column (each a few words of its
     * bitmap), so that a thread has at most eight entries and all their key (HBM/L2) and value loads are in flight
     * together: one round trip per step.  A column with no such entry (the usual case once the tail is full) costs
     * no memory access at all. */
    {
        int G = 1;
        while (G < 8 && G * 2 <= KW && rnz1 * G * 2 <= NT) G *= 2;
        for (int it = tid; it < rnz1 * G; it += NT) {
            const int kk = 1 + it / G, g = it % G;
            const int c = d.rlist[kk];
            const unsigned *cb = cbm + c * KW;
            double cmx = 0.0; unsigned kmin = 0xffffffffu;
            int w = g; unsigned tb = 0;
            if (w < KW) { tb = cb[w] & ~d.cmask[w]; if (w == (tp >> 5)) tb &= ~(1u << (tp & 31)); }
            for (;;) {
                int tt[8]; unsigned kq[8]; double xq[8];
                #pragma unroll
                for (int u = 0; u < 8; u++) {
                    while (tb == 0 && w + G < KW) { w += G; tb = cb[w] & ~d.cmask[w]; if (w == (tp >> 5)) tb &= ~(1u << (tp & 31)); }
                    if (tb) { tt[u] = w * 32 + __ffs((int)tb) - 1; tb &= tb - 1; } else tt[u] = -1;
                }
                if (tt[0] < 0) break;
                #pragma unroll
                for (int u = 0; u < 8; u++) {
                    kq[u] = tt[u] >= 0 ? dkey[(size_t)tt[u] * KD + c].c : 0xffffu;
                    xq[u] = tt[u] >= 0 ? DV((size_t)tt[u] * KD + c) : 0.0;
                }
                #pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (tt[u] < 0) continue;
                    const unsigned pk = (kq[u] << 8) | (unsigned)tt[u];
                    if (pk < kmin) kmin = pk;
                    const double ax = fabs(xq[u]);
                    if (ax > cmx) cmx = ax;
                }
                if (tt[7] < 0) break;
            }
            if (kmin != 0xffffffffu) atomicMin(&d.kminp[c], kmin);
            if (cmx > 0.0) atomicMax((unsigned long long *)&d.scm[c], (unsigned long long)__double_as_longlong(cmx));
        }
    }

    /* E. the rank-1 update.  First stage of a two-stage tail (values in HBM/L2), pivot_any: every warp streams its
     * share of the pivot column's rows through a private ring of shared-memory buffers -- a bulk copy brings a
     * whole row (kd * 8 bytes, contiguous) in, the warp updates it in place, a bulk copy takes it back -- so
     * that nbuf rows per warp are in flight whatever the register budget (cp.async.bulk + mbarrier, SASS
     * UBLKCP); with plain loads the step is bound by the latency of eight loads per thread. */
    if (RES == 2 && !SMALL && S.ring_nbuf >= 2) {
        const int NB = S.ring_nbuf;
        BLU_DYN_SMEM(dynr_);
        double *const mybuf = (double *)(dynr_ + dense_ring_offset(KD)) + (size_t)wid * NB * KD;
        u64 *const mybar = dense_ring_bars(dynr_, KD) + wid * NB;
        unsigned ph = S.ring_phase[wid];
        const unsigned rowbytes = (unsigned)(KD * sizeof(double));
        double a[8], cmx[8]; unsigned rkv[8]; unsigned inRm = 0; int pw = -1;
        #pragma unroll
        for (int cw = 0; cw < 8; cw++) {
            a[cw] = 0.0; cmx[cw] = 0.0; rkv[cw] = 0;
            if (cw < KW) {
                const int c = cw * 32 + lane;
                if ((d.rmask[cw] >> lane) & 1u) {
                    inRm |= 1u << cw;
                    a[cw] = d.arow[c];
                    rkv[cw] = ekey + (unsigned)d.posr[c];
                }
                if (c == cp) pw = cw;
            }
        }
        const int nmine = cnz1 > wid ? (cnz1 - 1 - wid) / NW + 1 : 0;      /* rows p = 1 + wid + i * NW */
        BluKey2 gone; gone.c = 0xffffu; gone.r = 0xffffu;
        if (lane == 0) {
            bulk_fence();
            for (int i = 0; i < NB && i < nmine; i++)
                bulk_copy_g2s(mybuf + (size_t)i * KD, dv + (size_t)d.clist[1 + wid + i * NW] * KD, rowbytes, &mybar[i]);
        }
        for (int i = 0; i < nmine; i++) {
            const int b = i % NB, p = 1 + wid + i * NW;
            const int t = d.clist[p];
            const double cv = d.cvalp[p];
            double *const buf = mybuf + (size_t)b * KD;
            mbar_wait(&mybar[b], (ph >> b) & 1u);
            ph ^= 1u << b;
            #pragma unroll
            for (int cw = 0; cw < 8; cw++) {      /* (no bit of inRm beyond the last word; straight-line code: selects and predicated stores, no divergent branches) */
                if (cw * 32 >= KD) continue;
                const int c = cw * 32 + lane;
                const bool in = (inRm >> cw) & 1u;
                const double x0 = buf[c];
                const double x1 = __dsub_rn(x0, __dmul_rn(a[cw], cv));
                const double x = in ? x1 : x0;
                if (in) buf[c] = x;
                BluKey2 kv;
                kv.c = in ? (unsigned short)(ekey + (unsigned)p) : (unsigned short)0xffffu;
                kv.r = in ? (unsigned short)rkv[cw] : (unsigned short)0xffffu;
                if (in || cw == pw) dkey[(size_t)t * KD + c] = kv;      /* (cw == pw: the pivot column leaves the active submatrix) */
                const double ax = in ? fabs(x) : 0.0;
                cmx[cw] = ax > cmx[cw] ? ax : cmx[cw];
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_copy_s2g(dv + (size_t)t * KD, buf, rowbytes);
                bulk_commit();
                if (i >= RING_LAG && i - RING_LAG + NB < nmine) {      /* a buffer is free once its store has read it: the store of RING_LAG rows ago has had time to */
                    bulk_wait_read<RING_LAG>();
                    const int bp = (i - RING_LAG) % NB;
                    bulk_copy_g2s(mybuf + (size_t)bp * KD, dv + (size_t)d.clist[1 + wid + (i - RING_LAG + NB) * NW] * KD, rowbytes, &mybar[bp]);
                }
            }
        }
        if (lane == 0) { bulk_wait<0>(); bulk_fence(); S.ring_phase[wid] = ph; }
        #pragma unroll
        for (int cw = 0; cw < 8; cw++)
            if (((inRm >> cw) & 1u) && cmx[cw] > 0.0)
                atomicMax((unsigned long long *)&d.scm[cw * 32 + lane], (unsigned long long)__double_as_longlong(cmx[cw]));
    } else {
        const int RS = NW >= KW ? NW / KW : 1;
        const int chunk = (cnz1 + RS - 1) / RS;
        BluKey2 gone; gone.c = 0xffffu; gone.r = 0xffffu;
        for (int u0 = wid; u0 < KW * RS; u0 += NW) {
            const int cw = u0 % KW, rs = u0 / KW;
            const unsigned rf = d.rfull[cw], rm = d.rmask[cw];
            if (rf == 0) continue;
            const int c = cw * 32 + lane;
            const bool inR = (rm >> lane) & 1u;
            const bool isP = c == cp;
            double a = 0.0; unsigned rkv = 0; int kk = 0;
            if (inR) { kk = d.posr[c]; a = d.arow[c]; rkv = ekey + (unsigned)kk; }
            double cmx = 0.0;
            u64 mydrop = 0;
            const int pend = (rs + 1) * chunk < cnz1 ? (rs + 1) * chunk : cnz1;
            constexpr int UR = RES == 1 ? 4 : 8;      /* (values in HBM/L2: more loads in flight) */
            for (int p0 = 1 + rs * chunk; p0 <= pend; p0 += UR) {
                int tt[UR]; double xv[UR], cv[UR];
                #pragma unroll
                for (int u = 0; u < UR; u++) {      /* the loads are in flight together */
                    const int p = p0 + u;
                    tt[u] = p <= pend ? (int)d.clist[p] : -1;
                    cv[u] = p <= pend ? d.cvalp[p] : 0.0;
                    xv[u] = (inR && tt[u] >= 0) ? DV((size_t)tt[u] * KD + c) : 0.0;
                }
                #pragma unroll
                for (int u = 0; u < UR; u++) {
                    if (tt[u] < 0) continue;       /* uniform: p does not depend on the lane */
                    const int p = p0 + u;
                    const size_t off = (size_t)tt[u] * KD + c;
                    /* straight-line code (selects and predicated stores): the lanes of a word differ in inR */
                    const double x = __dsub_rn(xv[u], __dmul_rn(a, cv[u]));
                    const double ax = fabs(x);
                    const int keep = inR && (SMALL ? ax > droptol : 1);
                    if (inR) dv[off] = keep ? x : 0.0;
                    BluKey2 kv;
                    kv.c = keep ? (unsigned short)(ekey + (unsigned)p) : (unsigned short)0xffffu;
                    kv.r = keep ? (unsigned short)rkv : (unsigned short)0xffffu;
                    if (inR || isP) dkey[off] = kv;      /* (dropped entries and the pivot column leave the active submatrix) */
                    cmx = (keep && ax > cmx) ? ax : cmx;
                    if (SMALL && inR && !keep) mydrop |= 1ull << (p - 1);
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

    /* F. counts, Markowitz keys, U row (pivot.rs:306-328), L column (:403-415); the pivot row and column leave */
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
        {
            const u64 k64 = mkckey(cnt, cbase + kk, cmx, abstol);
            d.skeyc[c] = k64;
            d.skey32[c] = k64 >= KEY_PARK ? 0xffffffffu : ((unsigned)cnt << 23) | (unsigned)(cbase + kk - sbase);
        }
        {   /* pivot.rs:261-262: the first entry outside the pivot column moves to the place of the pivot-row entry */
            const unsigned pk = d.kminp[c], kpr = rowk32[c] & 0xffffu;
            if ((pk >> 8) < kpr) dkey[(size_t)(pk & 0xffu) * KD + c].c = (unsigned short)kpr;
        }
        const double xr = DV((size_t)tp * KD + c);
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
    for (int w = NT - 1 - tid; w < KW; w += NT) { cbm[cp * KW + w] = 0; rbm[tp * KW + w] = 0; }
    if (tid == NT - 1) { d.skeyc[cp] = KEY_INF; d.skey32[cp] = 0xffffffffu; d.cnz[cp] = 0; d.rnz[tp] = 0; }
    bsync<NT>();

    /* the finisher warp closes the step while warp 0 already looks for the next candidates */
    if (wid == (NW > 1 ? 1 : 0)) {
        int ln = cnz1, un = rnz1;
        if (S.flag_b) ln = warp_squeeze(M.l_idx, M.l_val, lbase, cnz1);
        if (S.flag_a) un = warp_squeeze(M.u_idx, M.u_val, ubase, rnz1);
        if (lane == 0) {
            M.l_idx[lbase + ln] = -1;
            finish_step(S, rank, lbase + ln + 1, ubase + un, pivot, cnz1 + 1, rnz1 + 1);
            S.cstamp = cbase + rnz1 + 1;
            S.rstamp = rbase + cnz1 + 1;
            S.epoch++;
            S.n_kind[5]++; S.n_kind[SMALL ? 3 : 4]++;
            S.mbar_phase = phase ^ 1u;
        }
    }
    if (tid == 0) S.n_kind[7] += clock64() - tq;
}

template <int NT, int RES> __device__ __noinline__ int dense_run(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int KD = S.kd, KW = S.kw, nr = S.nrs, nc = S.ncs;
    DENSE_VIEW(RES);
    int maxsearch = M.prm.maxsearch;
    if (maxsearch < 1) maxsearch = 1;
    if (maxsearch > MAXCAND) maxsearch = MAXCAND;
    const double abstol = M.prm.abstol, reltol = M.prm.reltol;
    int rank = S.rank, rankdef = S.rankdef;      /* (the finisher warp updates S.rank behind the barrier) */
    int code = DRUN_DONE;
    for (;;) {
        if (rank + rankdef >= m) break;
        if (RES == 2 && m - rank <= S.kd_small) { code = DRUN_RESTAGE; break; }
        i64 t0 = clock64();
        /* A. the first `maxsearch` live columns in ascending (count, stamp) order: one warp, no block barriers */
        if (wid == 0) {
            int ncand = 0;
            unsigned prev = 0; int have_prev = 0;
            while (ncand < maxsearch) {
                unsigned k0 = 0xffffffffu, k1 = 0xffffffffu, k2 = 0xffffffffu;
                int j0 = -1, j1 = -1, j2 = -1;
                for (int c = lane; c < nc; c += 32) {
                    const unsigned kq = d.skey32[c];
                    if (kq == 0xffffffffu || (have_prev && kq <= prev)) continue;
                    if (kq < k2) {
                        if (kq < k1) {
                            k2 = k1; j2 = j1;
                            if (kq < k0) { k1 = k0; j1 = j0; k0 = kq; j0 = c; }
                            else { k1 = kq; j1 = c; }
                        } else { k2 = kq; j2 = c; }
                    }
                }
                int got = 0;
                for (int r = 0; r < 3 && ncand < maxsearch; r++) {
                    const unsigned best = warp_min_u32(k0);
                    if (best == 0xffffffffu) break;
                    if (k0 == best) {
                        S.cand_col[ncand] = j0;
                        k0 = k1; j0 = j1; k1 = k2; j1 = j2; k2 = 0xffffffffu; j2 = -1;
                    }
                    prev = best; have_prev = 1;
                    ncand++; got++;
                }
                if (got < 3) break;
            }
            if (lane == 0) { S.ncand = ncand; S.cand_done = 0; }
        }
        bsync<NT>();
        const int ncand = S.ncand;
        if (ncand == 0) { if (tid == 0) { BLU_CHECK(S, 0); } break; }
        if (key_cnt(d.skeyc[S.cand_col[0]]) == 0) {
            /* markowitz.rs:73-78 + factorize_bump.rs:23-31: an empty column is dropped without a pivot */
            bsync<NT>();
            if (tid == 0) {
                const int c0 = S.cand_col[0], pc = d.dcol[c0];
                S.dpc = c0; S.pivot_col = pc; S.pivot_row = -1;
                M.ckey[pc] = KEY_INF; d.skeyc[c0] = KEY_INF; d.skey32[c0] = 0xffffffffu; S.ndead++; S.rankdef++;
                S.t_phase[13] += clock64() - t0;
            }
            rankdef++;
            bsync<NT>();
            continue;
        }
        /* B. one warp per candidate column */
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
                {   /* kd <= 256: at most eight rows per lane; their key and value loads are issued back to back (one
                     * round trip to HBM/L2 for the keys, and for the values too in the first stage) */
                    unsigned kq[8]; double xq[8];
                    #pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int t = lane + 32 * u;
                        const bool on = t < nr && bit_test(cb, t);
                        kq[u] = on ? dkey[(size_t)t * KD + c].c : 0xffffffffu;
                        xq[u] = on ? DV((size_t)t * KD + c) : 0.0;
                    }
                    #pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int t = lane + 32 * u;
                        if (t < KD) stash[t] = kq[u];
                        if (kq[u] == 0xffffffffu) continue;
                        const double x = fabs(xq[u]);
                        if (x == 0.0 || x < tol) continue;
                        const u64 mc = (u64)((nz1 - 1) * (i64)(d.rnz[t] - 1));
                        const u64 key = (mc << 32) | (u64)kq[u];       /* ties: first in storage order, markowitz.rs:105 */
                        if (key < best) { best = key; bt = t; }
                    }
                }
            } else {
                for (int t = lane; t < nr; t += 32) {
                    if (!bit_test(cb, t)) continue;
                    const size_t off = (size_t)t * KD + c;
                    const double x = fabs(DV(off));
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
            int last = 0;
            if (lane == 0) {
                S.cand_mc[cc] = wb == KEY_INF ? -1 : (i64)(wb >> 32);
                S.cand_row[cc] = tt;
                __threadfence_block();
                last = atomicAdd(&S.cand_done, 1) == ncand - 1;
            }
            last = __shfl_sync(FULLMASK, last, 0);
            if (!last) continue;
            /* every candidate is evaluated: the choice (markowitz.rs:96-112, the first of minimum cost) */
            __threadfence_block();
            i64 mc64 = (i64)M.m * (i64)M.m;
            int pt = -1, pc = -1, pcc = -1;
            for (int q = 0; q < ncand; q++) {
                const i64 v = S.cand_mc[q];
                if (v >= 0 && v < mc64) { mc64 = v; pt = S.cand_row[q]; pc = S.cand_col[q]; pcc = q; }
            }
            if (pc < 0) { if (lane == 0) { BLU_CHECK(S, 0); } continue; }
            if (pcc >= DENSE_STASH) {      /* (maxsearch > DENSE_STASH) the chosen column's keys were not kept: row 0 is free now */
                const unsigned *cb2 = cbm + pc * KW;
                for (int t = lane; t < KD; t += 32) d.candk[t] = (t < nr && bit_test(cb2, t)) ? dkey[(size_t)t * KD + pc].c : 0xffffffffu;
                pcc = 0;
            }
            if (lane == 0) {
                S.dpt = pt; S.dpc = pc; S.pivot_row = d.drow[pt]; S.pivot_col = d.dcol[pc]; S.dpcand = pcc;
                S.nsearch += ncand;
                const int nz_col = d.cnz[pc], nz_row = d.rnz[pt];
                /* room in L and U, pivot.rs:69-81 */
                int st = BLU_OK;
                int room = M.l_mem - S.lput;
                if (room < nz_col) { M.info->addmem_l = nz_col - room; st = BLU_REALLOCATE; }
                room = M.u_mem - S.uput;
                if (room < nz_row - 1) { M.info->addmem_u = nz_row - 1 - room; st = BLU_REALLOCATE; }
                int general = 0;
                if (st != BLU_OK) S.status = st;
                else if (nz_row > 1 && nz_col > 2 && S.epoch < 255) {      /* (eight bits of epoch in the keys) */
                    general = 1;
                    /* the keys of the pivot row are one contiguous segment of kd * 4 bytes in HBM/L2: handed to the
                     * bulk-copy engine (cp.async.bulk + mbarrier), it arrives while the block ranks the pivot column */
                    bulk_fence();
                    bulk_copy_g2s(d.rowk, dkey + (size_t)pt * KD, (unsigned)(KD * sizeof(BluKey2)), &S.mbar);
                }
                S.dgeneral = general;
            }
        }
        bsync<NT>();
        if (tid == 0) S.t_phase[13] += clock64() - t0;
        if (S.status != BLU_OK) break;
        if (!S.dgeneral) { code = DRUN_SPARSE_PIVOT; break; }
        t0 = clock64();
        if ((int)d.cnz[S.dpc] - 1 <= MAXROW_SMALL) dense_step<NT, RES, true>(S, rank);
        else dense_step<NT, RES, false>(S, rank);
        rank++;
        if (tid == 0) S.t_phase[12] += clock64() - t0;
        if (S.status != BLU_OK) break;
        if (S.need_remove) { code = DRUN_REMOVE; break; }
    }
    bsync<NT>();      /* (the finisher warp is through) */
    return code;
}

/* ------------------------------------------------------------------ */
/* first stage (order kd_big, values in HBM/L2) -> second stage (order kd_small, everything in shared memory)  */
/* ------------------------------------------------------------------ */
/* The live rows and columns keep their relative slot order; values, bitmaps and per-slot arrays are compacted into
 * the layout of order kd_small, the keys into dn_key2 with their values unchanged (the epoch counter goes on: the
 * two stages together take at most kd_big - DENSE_MIN_ROWS < 255 steps).  Needs a thread per old slot. */
template <int NT> __device__ __noinline__ void dense_restage(Shm &S) {
    Mat &M = S.M;
    const int tid = threadIdx.x;
    const int KDb = S.kd, KWb = S.kw, KDs = S.kd_small, KWs = KDs / 32;
    const int nrb = S.nrs, ncb = S.ncs;
    BLU_DYN_SMEM(dyn_);
    DenseSm o, n;
    dense_ring_retire<NT>(S);
    dense_view(o, dyn_, KDb, KWb);      /* the two layouts overlap: everything of the old one goes through registers */
    dense_view(n, dyn_, KDs, KWs);
    /* a. per-slot state of the old layout */
    const int t = tid;
    const int ra = t < nrb && M.rkey[o.drow[t < nrb ? t : 0]] != KEY_INF;
    const int ca = t < ncb && o.skeyc[t < ncb ? t : 0] != KEY_INF;
    const int r_drow = ra ? o.drow[t] : 0, r_rnz = ra ? (int)o.rnz[t] : 0;
    const int c_dcol = ca ? o.dcol[t] : 0, c_cnz = ca ? (int)o.cnz[t] : 0;
    const u64 c_key = ca ? o.skeyc[t] : KEY_INF, c_cm = ca ? o.scm[t] : 0;
    const unsigned c_k32 = ca ? o.skey32[t] : 0xffffffffu;
    int nr2, nc2;
    const int rpos = block_excl_scan<NT>(ra, &nr2, S.iscr);
    const int cpos = block_excl_scan<NT>(ca, &nc2, S.iscr);
    if (nr2 > KDs || nc2 > KDs) { if (tid == 0) BLU_CHECK(S, 0); bsync<NT>(); return; }
    /* b. per-slot state of the new layout; tmps / tmpr map the new slots to the old ones */
    for (int q = tid; q < KDs; q += NT) { n.skeyc[q] = KEY_INF; n.skey32[q] = 0xffffffffu; n.scm[q] = 0; n.cnz[q] = 0; n.rnz[q] = 0; }
    bsync<NT>();
    if (ra) { n.drow[rpos] = r_drow; n.rnz[rpos] = (unsigned short)r_rnz; n.tmps[rpos] = (unsigned short)t; }
    if (ca) { n.dcol[cpos] = c_dcol; n.cnz[cpos] = (unsigned short)c_cnz; n.skeyc[cpos] = c_key; n.skey32[cpos] = c_k32; n.scm[cpos] = c_cm; n.tmpr[cpos] = (unsigned short)t; }
    bsync<NT>();
    /* c. bitmaps: the old ones sit behind the old per-slot arrays, clear of the new per-slot arrays but not of the
     * new bitmaps, so the new words wait in registers (at most 8 per thread: NT >= kd_big) until all are computed */
    {
        unsigned rwv[8], cwv[8];
        #pragma unroll
        for (int i = 0; i < 8; i++) {
            const int q = tid + i * NT;
            unsigned rw = 0, cw = 0;
            if (q < KDs * KWs) {
                const int a = q / KWs, w = q % KWs;
                if (a < nr2) {
                    const unsigned *orow = o.rb_s + (int)n.tmps[a] * KWb;
                    for (int j = 0; j < 32; j++) { const int c2 = w * 32 + j; if (c2 < nc2 && bit_test(orow, n.tmpr[c2])) rw |= 1u << j; }
                }
                if (a < nc2) {
                    const unsigned *ocol = o.cb_s + (int)n.tmpr[a] * KWb;
                    for (int j = 0; j < 32; j++) { const int t2 = w * 32 + j; if (t2 < nr2 && bit_test(ocol, n.tmps[t2])) cw |= 1u << j; }
                }
            }
            rwv[i] = rw; cwv[i] = cw;
        }
        bsync<NT>();
        #pragma unroll
        for (int i = 0; i < 8; i++) {
            const int q = tid + i * NT;
            if (q < KDs * KWs) { n.rb_s[q] = rwv[i]; n.cb_s[q] = cwv[i]; }
        }
    }
    bsync<NT>();
    /* d. values and keys (the values overwrite the old bitmaps) */
    {
        const double *dvb = M.dn_val;
        const unsigned *kb = (const unsigned *)M.dn_key;
        unsigned *ks = (unsigned *)M.dn_key2;
        for (int q = tid; q < KDs * KDs; q += NT) {
            const int a = q / KDs, c2 = q % KDs;
            double v = 0.0; unsigned key = 0xffffffffu;
            if (a < nr2 && c2 < nc2 && bit_test(n.rb_s + a * KWs, c2)) {
                const size_t off = (size_t)n.tmps[a] * KDb + n.tmpr[c2];
                v = ld_l2(dvb + off); key = kb[off];
            }
            n.dv_s[q] = v; ks[q] = key;
        }
    }
    bsync<NT>();
    if (tid == 0) {
        S.kd = KDs; S.kw = KWs; S.dv_smem = 1; S.nrs = nr2; S.ncs = nc2;
        M.dn_key = M.dn_key2;
        S.n_kind[6]++;
    }
    bsync<NT>();
}

#endif
