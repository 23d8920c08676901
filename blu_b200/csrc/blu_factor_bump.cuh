/* blu_factor_bump.cuh -- phase 3: Markowitz search + one elimination step per pivot,
 * looped inside the kernel (reference: src/lu/factorize_bump.rs:12-49,
 * src/lu/markowitz.rs:34-219, src/lu/pivot.rs:48-1381).
 *
 * Parallel formulation that keeps the reference's pivot sequence bit for bit:
 *  - the count buckets of list.rs are FIFO by last insertion; a column/row carries the
 *    key (count << 40 | stamp) with a monotone stamp handed out in the order the
 *    reference would call list_add, so "walk bucket nz from the head" == ascending key;
 *  - lines (columns with values, rows pattern-only) are updated by one warp each with
 *    ballot/prefix compaction, which reproduces the in-line storage order of the
 *    sequential code;
 *  - a*b and the subtraction are separate roundings (__dmul_rn/__dsub_rn) like Rust.
 * Where a line lives in memory is free (file.rs keeps only in-line order), so lines
 * that outgrow their slack move to the end of the live W half (atomic bump pointer)
 * and garbage collection copies all live lines to the other half.
 */
#ifndef BLU_FACTOR_BUMP_CUH
#define BLU_FACTOR_BUMP_CUH
#ifndef PF_DIST
#define PF_DIST 2   /* how many of its own lines ahead a warp prefetches into L2 */
#endif
#ifndef SEARCH_MERGE
#define SEARCH_MERGE 0   /* measured on B200: the every-thread merge costs 3 % more than three block-wide minima */
#endif
#ifndef REGE
#define REGE 8   /* line entries per lane held in registers by the fast paths (lines <= 32*REGE) */
#endif
#include "blu_dev_common.cuh"

/* ------------------------------------------------------------------ */
/* garbage collection (role of file.rs:92 file_compress)               */
/* ------------------------------------------------------------------ */
template <int NT> __device__ __noinline__ void w_compact(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int nbase = (1 - S.w_half) * M.w_mem;
    int *newbeg = M.tmpi;       /* 2m */
    int *newcap = M.tmpi + 2 * m;
    /* try with slack first; if that does not fit fall back to exact sizes */
    for (int attempt = 0; attempt < 2; attempt++) {
        int put = 0;
        for (int base = 0; base < 2 * m; base += NT) {
            int l = base + tid;
            int sz = 0, nz = 0;
            if (l < 2 * m) {
                int live = l < m ? (M.ckey[l] != KEY_INF) : (M.rkey[l - m] != KEY_INF);
                nz = live ? M.lend[l] - M.lbeg[l] : 0;
                sz = (live && nz > 0) ? nz + (attempt == 0 ? slack_of(M.prm, nz) : 0) : 0;
            }
            int tot, ex = block_excl_scan<NT>(sz, &tot, S.iscr);
            if (l < 2 * m) { newbeg[l] = nbase + put + ex; newcap[l] = nbase + put + ex + sz; }
            put += tot;
        }
        bsync<NT>();
        if (put <= M.w_mem) { if (tid == 0) S.w_used = nbase + put; break; }
        if (attempt == 1 && tid == 0) BLU_CHECK(S, 0);
    }
    bsync<NT>();
    for (int l = wid; l < 2 * m; l += NW) {
        int ob = M.lbeg[l], oe = M.lend[l];
        int live = l < m ? (M.ckey[l] != KEY_INF) : (M.rkey[l - m] != KEY_INF);
        int nb = newbeg[l];
        if (live && oe > ob) {
            for (int t = lane; t < oe - ob; t += 32) {
                M.w_idx[nb + t] = M.w_idx[ob + t];
                if (l < m) M.w_val[nb + t] = M.w_val[ob + t];
            }
        }
        __syncwarp();
        if (lane == 0) {
            int nz = (live && oe > ob) ? oe - ob : 0;
            M.lbeg[l] = nb; M.lend[l] = nb + nz; M.lcap[l] = newcap[l];
        }
    }
    if (tid == 0) {
        S.w_half = 1 - S.w_half;
        S.w_limit = nbase + M.w_mem;
        S.ngarbage++;
    }
    bsync<NT>();
}

/* Make sure `grow` more slots can be handed out; may compact.  Uniform result. */
template <int NT> __device__ bool w_reserve(Shm &S, i64 grow) {
    if (grow > (i64)(S.w_limit - S.w_used)) {
        w_compact<NT>(S);
        if (grow > (i64)(S.w_limit - S.w_used)) {
            if (threadIdx.x == 0) { S.M.info->addmem_w = grow - (S.w_limit - S.w_used); S.status = BLU_REALLOCATE; }
            bsync<NT>();
            return false;
        }
    }
    return true;
}

/* ------------------------------------------------------------------ */
/* min-tree over the column keys                                       */
/* The count buckets of list.rs:54-104 give the reference "first       */
/* column of the smallest non-empty bucket" in O(1); with keys         */
/* (count << 40 | stamp) the same column is the minimum key, and the   */
/* next ones in the bucket walk (markowitz.rs:82-123) are the next     */
/* smallest keys.  For bumps of 10^4..10^6 columns a scan of the       */
/* active list per pivot is O(bump); this tree (fanout 32, level 0 =   */
/* ckey itself) answers the first few minima in O(levels) and is       */
/* repaired in O(levels) per column whose key changed.                 */
/* ------------------------------------------------------------------ */
__device__ __forceinline__ const u64 *ctree_level(const Shm &S, int l) { return l == 0 ? S.M.ckey : S.M.ctree + S.tree_off[l]; }

template <int NT> __device__ __noinline__ void ctree_build(Shm &S) {
    Mat &M = S.M;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    if (tid == 0) {
        int n = M.m, off = 0, l = 0;
        S.tree_n[0] = n; S.tree_off[0] = 0;
        while (n > 32 && l < 7) { n = (n + 31) / 32; l++; S.tree_n[l] = n; S.tree_off[l] = off; off += n; }
        S.tree_levels = l;
    }
    bsync<NT>();
    for (int l = 1; l <= S.tree_levels; l++) {
        const u64 *src = ctree_level(S, l - 1);
        u64 *dst = M.ctree + S.tree_off[l];
        const int nsrc = S.tree_n[l - 1];
        for (int q = wid; q < S.tree_n[l]; q += NW) {
            const int c = q * 32 + lane;
            const u64 k = warp_min64(c < nsrc ? src[c] : KEY_INF);
            if (lane == 0) dst[q] = k;
        }
        bsync<NT>();
    }
}

/* the keys of columns cols[0..ncols) (and of column `extra` if >= 0) changed: repair their ancestors */
template <int NT> __device__ __noinline__ void ctree_update(Shm &S, const int *cols, int ncols, int extra) {
    Mat &M = S.M;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int tot = ncols + (extra >= 0 ? 1 : 0);
    for (int l = 1; l <= S.tree_levels; l++) {
        const u64 *src = ctree_level(S, l - 1);
        u64 *dst = M.ctree + S.tree_off[l];
        const int nsrc = S.tree_n[l - 1];
        for (int q = wid; q < tot; q += NW) {
            const int j = q < ncols ? cols[q] : extra;
            const int node = j >> (5 * l);
            const int c = node * 32 + lane;
            const u64 k = warp_min64(c < nsrc ? src[c] : KEY_INF);
            if (lane == 0) dst[node] = k;
        }
        bsync<NT>();
    }
}

/* the first `maxsearch` live columns in ascending key order -> S.cand_col / S.ncand.  One warp; best-first
 * walk from the root: the frontier holds rows of 32 sibling keys in shared memory (fk, fi), every lane keeps
 * the minimum of its own column of the frontier in registers. */
static __device__ __noinline__ void ctree_topk(Shm &S, int maxsearch, u64 *fk, int *fi, int *flev) {
    const int lane = threadIdx.x & 31;
    int nrows = 0, ncand = 0;
    u64 lk = KEY_INF; int lr = -1;
    int level = S.tree_levels, base = 0;      /* the row to add next */
    for (;;) {
        {   /* add the 32 children [base, base+32) of `level` */
            const int r = nrows++;
            const int c = base + lane;
            const u64 k = c < S.tree_n[level] ? ctree_level(S, level)[c] : KEY_INF;
            fk[r * 32 + lane] = k; fi[r * 32 + lane] = c;
            if (lane == 0) flev[r] = level;
            if (k < lk) { lk = k; lr = r; }
            __syncwarp();
        }
        for (;;) {
            const u64 best = warp_min64(lk);
            if (best >= KEY_PARK) { if (lane == 0) S.ncand = ncand; __syncwarp(); return; }
            const int owner = __ffs((int)__ballot_sync(FULLMASK, lk == best)) - 1;
            const int r = __shfl_sync(FULLMASK, lr, owner);
            const int lev = flev[r], id = fi[r * 32 + owner];
            if (lane == owner) {      /* consume the entry; this lane's minimum is recomputed over its column */
                fk[r * 32 + lane] = KEY_INF;
                lk = KEY_INF; lr = -1;
                for (int q = 0; q < nrows; q++) { const u64 k = fk[q * 32 + lane]; if (k < lk) { lk = k; lr = q; } }
            }
            __syncwarp();
            if (lev == 0) {
                if (lane == 0) S.cand_col[ncand] = id;
                ncand++;
                if (ncand >= maxsearch) { if (lane == 0) S.ncand = ncand; __syncwarp(); return; }
            } else { level = lev - 1; base = id * 32; break; }
        }
    }
}

/* ------------------------------------------------------------------ */
/* Markowitz search, markowitz.rs:34-193 (search_rows == 0 path)       */
/* ------------------------------------------------------------------ */
template <int NT> __device__ void markowitz_search(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    int maxsearch = M.prm.maxsearch;
    if (maxsearch < 1) maxsearch = 1;
    if (maxsearch > MAXCAND) maxsearch = MAXCAND;

    /* shrink the active-column list when more than half of it is dead */
    if (!S.use_tree && S.ndead * 2 > S.nact && S.nact > 2 * NT) {
        int *tmp = M.tmpi;
        int put = 0;
        for (int base = 0; base < S.nact; base += NT) {
            int t = base + tid;
            int j = t < S.nact ? M.acols[t] : -1;
            int live = j >= 0 && M.ckey[j] != KEY_INF;
            int tot, ex = block_excl_scan<NT>(live, &tot, S.iscr);
            if (live) tmp[put + ex] = j;
            put += tot;
        }
        bsync<NT>();
        for (int t = tid; t < put; t += NT) M.acols[t] = tmp[t];
        if (tid == 0) { S.nact = put; S.ndead = 0; }
        bsync<NT>();
    }
    const int nact = S.nact;

    /* the first `maxsearch` live columns in ascending (count, stamp) order */
    int ncand = 0;
    u64 prev = 0; int have_prev = 0;
    if (S.use_tree) {
        if (wid == 0) {
            /* frontier: 1 + maxsearch * levels rows of 32 (key, node) pairs in the (idle) pivot-step buffers */
            u64 *fk = (u64 *)S.cval;
            const int maxrows = 2 + maxsearch * (S.tree_levels + 1);
            int *fi = (int *)(fk + maxrows * 32);
            ctree_topk(S, maxsearch, fk, fi, fi + maxrows * 32);
        }
        bsync<NT>();
        ncand = S.ncand;
    } else
    while (ncand < maxsearch) {
        u64 k0 = KEY_INF, k1 = KEY_INF, k2 = KEY_INF;
        int j0 = -1, j1 = -1, j2 = -1;
        /* four independent (slot -> column -> key) chains in flight per thread: the loop is a
         * pure latency chain otherwise (measured: 17k cycles per search before) */
        for (int t0 = tid; t0 < nact; t0 += 4 * NT) {
            int jj[4]; u64 kk[4];
            #pragma unroll
            for (int u = 0; u < 4; u++) { const int t = t0 + u * NT; jj[u] = t < nact ? M.acols[t] : -1; }
            #pragma unroll
            for (int u = 0; u < 4; u++) kk[u] = jj[u] >= 0 ? M.ckey[jj[u]] : KEY_INF;
            #pragma unroll
            for (int u = 0; u < 4; u++) {
                const u64 k = kk[u]; const int j = jj[u];
                if (k >= KEY_PARK || (have_prev && k <= prev)) continue;      /* dead (KEY_INF) or passed over (mkckey) */
                if (k < k2) {
                    if (k < k1) {
                        k2 = k1; j2 = j1;
                        if (k < k0) { k1 = k0; j1 = j0; k0 = k; j0 = j; }
                        else { k1 = k; j1 = j; }
                    } else { k2 = k; j2 = j; }
                }
            }
        }
        int got = 0;
        if (SEARCH_MERGE && 3 * NW <= 40) {
            /* three smallest of the warp by shuffles, then ONE barrier and a 3-of-(3*NW) merge that every
             * thread runs for itself (keys are unique): 2 barriers per pass instead of 7.  The per-warp
             * results sit in the block-reduction scratch (40 entries), hence the bound on NW. */
            #pragma unroll
            for (int r = 0; r < 3; r++) {
                const u64 best = warp_min64(k0);
                const unsigned own = __ballot_sync(FULLMASK, k0 == best && best != KEY_INF);
                int jb = -1;
                if (own) jb = __shfl_sync(FULLMASK, j0, __ffs((int)own) - 1);
                if (lane == 0) { S.kscr[wid * 3 + r] = best; S.iscr[wid * 3 + r] = jb; }
                if (k0 == best && best != KEY_INF) { k0 = k1; j0 = j1; k1 = k2; j1 = j2; k2 = KEY_INF; j2 = -1; }
            }
            bsync<NT>();
            u64 b0 = KEY_INF, b1 = KEY_INF, b2 = KEY_INF;
            int c0 = -1, c1 = -1, c2 = -1;
            for (int q = 0; q < 3 * NW; q++) {
                const u64 k = S.kscr[q]; const int j = S.iscr[q];
                if (k < b2) {
                    if (k < b1) {
                        b2 = b1; c2 = c1;
                        if (k < b0) { b1 = b0; c1 = c0; b0 = k; c0 = j; }
                        else { b1 = k; c1 = j; }
                    } else { b2 = k; c2 = j; }
                }
            }
            if (b0 != KEY_INF && ncand < maxsearch) { if (tid == 0) S.cand_col[ncand] = c0; prev = b0; have_prev = 1; ncand++; got++; }
            if (b1 != KEY_INF && ncand < maxsearch) { if (tid == 0) S.cand_col[ncand] = c1; prev = b1; ncand++; got++; }
            if (b2 != KEY_INF && ncand < maxsearch) { if (tid == 0) S.cand_col[ncand] = c2; prev = b2; ncand++; got++; }
        } else {
            for (int r = 0; r < 3 && ncand < maxsearch; r++) {
                u64 best = block_min64<NT>(k0, S.kscr);
                if (best == KEY_INF) break;
                if (k0 == best) {
                    S.cand_col[ncand] = j0;
                    k0 = k1; j0 = j1; k1 = k2; j1 = j2; k2 = KEY_INF; j2 = -1;
                }
                prev = best; have_prev = 1;
                ncand++; got++;
            }
        }
        bsync<NT>();
        if (got < 3) break;     /* fewer live columns than asked for */
    }
    if (ncand == 0) { if (tid == 0) { BLU_CHECK(S, 0); } bsync<NT>(); return; }

    /* empty column => rank-deficiency step (markowitz.rs:73-78): bucket 0 sorts first */
    if (key_cnt(M.ckey[S.cand_col[0]]) == 0) {
        if (tid == 0) { S.pivot_col = S.cand_col[0]; S.pivot_row = -1; }
        bsync<NT>();
        return;
    }

    /* evaluate the candidates, one warp per column (markowitz.rs:82-123) */
    const double abstol = M.prm.abstol, reltol = M.prm.reltol;
    for (int c = wid; c < ncand; c += NW) {
        const int j = S.cand_col[c];
        const int beg = M.lbeg[j], end = M.lend[j];
        const i64 nz1 = end - beg;
        const double cmx = M.colpiv[j];
        const double tol = fmax(abstol, reltol * cmx);
        u64 bestmc = KEY_INF; int bestpos = 0x7fffffff;
        for (int pos = beg + lane; pos < end; pos += 32) {
            double x = fabs(M.w_val[pos]);
            if (x == 0.0 || x < tol) continue;
            int i = M.w_idx[pos];
            i64 nz2 = M.lend[m + i] - M.lbeg[m + i];
            u64 mc = (u64)((nz1 - 1) * (nz2 - 1));
            if (mc < bestmc) { bestmc = mc; bestpos = pos; }
        }
        u64 wmc = warp_min64(bestmc);
        int p = (bestmc == wmc && wmc != KEY_INF) ? bestpos : 0x7fffffff;
        p = warp_min(p);
        if (lane == 0) {
            S.cand_mc[c] = wmc == KEY_INF ? -1 : (i64)wmc;
            S.cand_row[c] = wmc == KEY_INF ? -1 : M.w_idx[p];
        }
    }
    bsync<NT>();
    if (tid == 0) {
        i64 mc64 = (i64)m * (i64)m;
        int pr = -1, pc = -1;
        for (int c = 0; c < ncand; c++) {
            if (S.cand_mc[c] >= 0 && S.cand_mc[c] < mc64) { mc64 = S.cand_mc[c]; pr = S.cand_row[c]; pc = S.cand_col[c]; }
        }
        BLU_CHECK(S, pc >= 0);
        S.pivot_row = pr; S.pivot_col = pc;
        S.nsearch += ncand;
    }
    bsync<NT>();
}

/* ------------------------------------------------------------------ */
/* Markowitz search with search_rows != 0, markowitz.rs:34-219          */
/* Columns and rows are visited in the reference's order -- count by    */
/* count, the columns of a count before its rows, FIFO inside a bucket  */
/* (ascending key) -- and every early exit (markowitz.rs:109, 171) and  */
/* the parking of rows whose cheap entries are all unstable (:178-179,  */
/* bucket m+1: key count m+1 until an update re-stamps the row,         */
/* pivot.rs:387-397) is reproduced.  The walk is sequential by nature   */
/* (each item is judged against the best cost so far), so one item at a */
/* time: the block finds the next column / row by key, warp 0 judges it.*/
/* ------------------------------------------------------------------ */
template <int NT> __device__ __noinline__ void markowitz_search_rows(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double abstol = M.prm.abstol, reltol = M.prm.reltol;
    int maxsearch = M.prm.maxsearch;
    if (maxsearch < 1) maxsearch = 1;
    u64 prevc = 0, prevr = 0; int havec = 0, haver = 0;
    if (tid == 0) { S.pivot_row = -1; S.pivot_col = -1; S.cand_mc[0] = (i64)m * (i64)m; S.ncand = 0; S.flag_a = 0; }
    bsync<NT>();
    for (;;) {
        /* next column and next row in bucket order */
        u64 kc = KEY_INF, kr = KEY_INF; int jc = -1, ir = -1;
        for (int t = tid; t < S.nact; t += NT) {
            const int j = M.acols[t];
            const u64 k = M.ckey[j];
            if (k >= KEY_PARK || (havec && k <= prevc)) continue;
            if (k < kc) { kc = k; jc = j; }
        }
        for (int i = tid; i < m; i += NT) {
            const u64 k = M.rkey[i];
            if (k == KEY_INF || key_cnt(k) > m || key_cnt(k) == 0 || (haver && k <= prevr)) continue;      /* gone, parked in bucket m+1, or empty (the walk starts at count 1, markowitz.rs:80) */
            if (k < kr) { kr = k; ir = i; }
        }
        const u64 bc = block_min64<NT>(kc, S.kscr);
        const u64 br = block_min64<NT>(kr, S.kscr);
        if (bc == KEY_INF && br == KEY_INF) break;
        if (kc == bc && bc != KEY_INF) S.iscr[38] = jc;
        if (kr == br && br != KEY_INF) S.iscr[39] = ir;
        bsync<NT>();
        const bool take_col = bc != KEY_INF && (br == KEY_INF || key_cnt(bc) <= key_cnt(br));
        if (take_col && key_cnt(bc) == 0 && !havec && !haver) {      /* markowitz.rs:73-78: bucket 0 is looked at first (a local test: tid 0 writes S.pivot_col below) */
            if (tid == 0) { S.pivot_col = S.iscr[38]; S.pivot_row = -1; }
            bsync<NT>();
            return;
        }
        if (wid == 0) {
            i64 mc64 = S.cand_mc[0];
            int nsearch = S.ncand, stop = 0;
            if (take_col) {
                const int j = S.iscr[38];
                const i64 nz1 = key_cnt(bc);
                const double cmx = M.colpiv[j];
                const int beg = M.lbeg[j], end = M.lend[j];
                const double tol = fmax(abstol, reltol * cmx);
                const i64 thr = (nz1 - 1) * (nz1 - 1);
                /* markowitz.rs:94-112: strict improvements in storage order; the first improving entry within
                 * the early-exit bound ends the search */
                u64 cmin = KEY_INF; int cminpos = -1;
                for (int base = beg; base < end && !stop; base += 32) {
                    const int pos = base + lane;
                    i64 mc = -1; int i = -1;
                    if (pos < end) {
                        const double x = fabs(M.w_val[pos]);
                        if (!(x == 0.0 || x < tol)) { i = M.w_idx[pos]; mc = (nz1 - 1) * (i64)(M.lend[m + i] - M.lbeg[m + i] - 1); }
                    }
                    const unsigned hit = __ballot_sync(FULLMASK, mc >= 0 && mc < mc64 && mc <= thr);
                    if (hit) {
                        const int l = __ffs((int)hit) - 1;
                        /* entries before it in this chunk may improve without ending the search, but the exit entry wins */
                        const int ii = __shfl_sync(FULLMASK, i, l);
                        const i64 mcl = __shfl_sync(FULLMASK, mc, l);
                        if (lane == 0) { S.pivot_row = ii; S.pivot_col = j; S.cand_mc[0] = mcl; }
                        stop = 1;
                        break;
                    }
                    const u64 key = mc >= 0 ? (((u64)mc << 32) | (u64)(unsigned)(pos - beg)) : KEY_INF;
                    const u64 wmin = warp_min64(key);
                    if (wmin < cmin) { cmin = wmin; cminpos = beg + (int)(wmin & 0xffffffffu); }
                }
                if (!stop) {
                    if (cmin != KEY_INF && (i64)(cmin >> 32) < mc64) {
                        if (lane == 0) { S.pivot_row = M.w_idx[cminpos]; S.pivot_col = j; S.cand_mc[0] = (i64)(cmin >> 32); }
                    }
                    nsearch++;
                    if (nsearch >= maxsearch) stop = 1;
                }
            } else {
                const int i = S.iscr[39];
                const i64 nz1 = key_cnt(br);
                const int rb = M.lbeg[m + i], re = M.lend[m + i];
                int cheap = 0, found = 0;
                for (int rpos = rb; rpos < re && !stop; rpos++) {      /* markowitz.rs:142-176 */
                    const int j = M.w_idx[rpos];
                    const int cb = M.lbeg[j], ce = M.lend[j];
                    const i64 mc = (nz1 - 1) * (i64)(ce - cb - 1);
                    if (mc >= mc64) continue;
                    cheap = 1;
                    const double cmx = M.colpiv[j];
                    if (cmx == 0.0 || cmx < abstol) continue;
                    double x = 0.0;
                    for (int base = cb; base < ce; base += 32) {
                        const int pos = base + lane;
                        const int h = pos < ce && M.w_idx[pos] == i;
                        const unsigned hm = __ballot_sync(FULLMASK, h);
                        if (hm) { const double v = h ? M.w_val[pos] : 0.0; x = fabs(__shfl_sync(FULLMASK, v, __ffs((int)hm) - 1)); break; }
                    }
                    if (x >= abstol && x >= reltol * cmx) {
                        found = 1;
                        mc64 = mc;
                        if (lane == 0) { S.pivot_row = i; S.pivot_col = j; S.cand_mc[0] = mc; }
                        if (mc64 <= nz1 * (nz1 - 1)) stop = 1;
                    }
                }
                if (!stop) {
                    if (cheap && !found) {
                        if (lane == 0) { M.rkey[i] = mkkey(m + 1, S.rstamp); S.rstamp++; }      /* parked, markowitz.rs:178-179 */
                    } else {
                        nsearch++;
                        if (nsearch >= maxsearch) stop = 1;
                    }
                }
            }
            if (lane == 0) { S.ncand = nsearch; S.flag_a = stop; }
        }
        bsync<NT>();
        if (S.flag_a) break;
        if (take_col) { prevc = bc; havec = 1; } else { prevr = br; haver = 1; }
    }
    if (tid == 0) {
        BLU_CHECK(S, S.pivot_col >= 0);
        S.nsearch += S.ncand;      /* (flag_a is lowered by the next user after a barrier of its own, not here: other threads may still be reading it) */
    }
    bsync<NT>();
}

/* ------------------------------------------------------------------ */
/* helpers shared by the elimination variants                          */
/* ------------------------------------------------------------------ */

/* squeeze out entries marked idx == -2 from [base, base+n) keeping order; one warp.
 * Returns the new count (uniform). */
__device__ __forceinline__ int warp_squeeze(int *idx, double *val, int base, int n) {
    const int lane = threadIdx.x & 31;
    int put = base;
    for (int b = 0; b < n; b += 32) {
        int pos = base + b + lane;
        int ok = (b + lane) < n;
        int i = ok ? idx[pos] : -2; double v = ok ? val[pos] : 0.0;
        int keep = ok && i != -2;
        unsigned km = __ballot_sync(FULLMASK, keep);
        __syncwarp();
        if (keep) { int d = put + __popc(km & lanemask_lt()); idx[d] = i; val[d] = v; }
        put += __popc(km);
        __syncwarp();
    }
    return put - base;
}

/* pivot.rs:1333-1381: empty column j whose max dropped below abstol.  One warp. */
__device__ __forceinline__ void warp_remove_col(Shm &S, int j) {
    Mat &M = S.M;
    const int m = M.m, lane = threadIdx.x & 31;
    const int cbeg = M.lbeg[j], cend = M.lend[j];
    for (int pos = cbeg; pos < cend; pos++) {
        const int i = M.w_idx[pos];
        const int rb = M.lbeg[m + i], re = M.lend[m + i];
        int where = -1;
        for (int b = rb; b < re; b += 32) {
            int q = b + lane;
            int hit = q < re && M.w_idx[q] == j;
            unsigned hm = __ballot_sync(FULLMASK, hit);
            if (hm) { where = b + __ffs((int)hm) - 1; break; }
        }
        if (lane == 0) {
            if (where < 0) { BLU_CHECK(S, 0); }
            else {
                M.w_idx[where] = M.w_idx[re - 1];
                M.lend[m + i] = re - 1;
                M.rkey[i] = mkkey(re - 1 - rb, S.rstamp++);
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        M.colpiv[j] = 0.0;
        M.lend[j] = cbeg;
        M.ckey[j] = mkkey(0, S.cstamp++);
    }
    __syncwarp();
}

/* pivot.rs:96-106: after a step, empty every touched column whose max is 0 or < abstol */
template <int NT> __device__ void post_remove_cols(Shm &S, int rank) {
    Mat &M = S.M;
    /* the flag was raised before the pivot step's closing barrier, so every thread reads the same value
     * here; it is lowered only after a barrier of its own, when every thread has read it */
    if (S.need_remove) {
        if ((threadIdx.x >> 5) == 0) {
            const double abstol = M.prm.abstol;
            for (int pos = M.u_begin[rank]; pos < M.u_begin[rank + 1]; pos++) {
                int j = M.u_idx[pos];
                double c = M.colpiv[j];
                if (c == 0.0 || c < abstol) warp_remove_col(S, j);
            }
        }
        bsync<NT>();
        if (threadIdx.x == 0) S.need_remove = 0;
    }
}

/* common tail of every variant: pointers, pivot value, unlink pivot row/column */
__device__ __forceinline__ void finish_step(Shm &S, int rank, int lput, int uput, double pivot,
                                            int nz_col, int nz_row) {
    Mat &M = S.M;
    const int m = M.m, pc = S.pivot_col, pr = S.pivot_row;
    M.l_begin_p[rank + 1] = lput;
    M.u_begin[rank + 1] = uput;
    S.lput = lput; S.uput = uput;
    M.colpiv[pc] = pivot;
    if (!S.dense) {      /* (the dense tail rebuilds the line table when it ends) */
        M.lend[pc] = M.lbeg[pc];
        M.lend[m + pr] = M.lbeg[m + pr];
    }
    M.ckey[pc] = KEY_INF;
    M.rkey[pr] = KEY_INF;
    S.ndead++;
    /* factorize_bump.rs:39-43; published by the barrier that ends every pivot variant */
    M.pinv[pr] = rank; M.qinv[pc] = rank;
    S.rank = rank + 1;
    S.factor_flops += (i64)(nz_col - 1) * (i64)(nz_row - 1);
    S.elim_bytes += 12.0 * nz_col + 4.0 * nz_row + 12.0 * (nz_col - 1) + 12.0 * (nz_row - 1);
    S.nelim_div += nz_col - 1;
}

/* ------------------------------------------------------------------ */
/* pivot_any (pivot.rs:114-458) and pivot_small (pivot.rs:460-833)     */
/* One body, two homes for the per-step state.                          */
/*  FAST (m <= SMARK_MAX and pivot column and row fit the cache): the    */
/*    pivot column and row, the line headers (begin, end, capacity) of   */
/*    every line the step touches and the row/column marks are staged in */
/*    shared memory by ONE parallel pass over the pivot column and row   */
/*    (the pass that also bounds the growth, pivot.rs:156-208), so a     */
/*    column update is  load line -> compute -> store  instead of        */
/*    header -> line -> mark lookup -> compute -> store; the line two    */
/*    ahead of each warp is pulled into L2 meanwhile.                    */
/*  general: marks and headers in global memory, the pivot column / row  */
/*    cached in shared memory when they fit, lines of any length.        */
/* Arithmetic and storage order are identical in both.                   */
/* ------------------------------------------------------------------ */
template <int NT, bool FAST> __device__ void pivot_general_t(Shm &S, const bool small) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int pc = S.pivot_col, pr = S.pivot_row, rank = S.rank;
    const double droptol = M.prm.droptol, abstol = M.prm.abstol;
    int cbeg = M.lbeg[pc], rbeg = M.lbeg[m + pr];
    const int cnz1 = M.lend[pc] - cbeg - 1, rnz1 = M.lend[m + pr] - rbeg - 1;
    const int *cidx, *ridx; const double *cval; double *work;
    double pivot;

    if (FAST) {
        /* one pass: stage column, row and headers; bound the growth; find the pivot */
        i64 grow = 0;
        for (int p = tid; p <= cnz1; p += NT) {
            const int i = M.w_idx[cbeg + p];
            S.cidx[p] = i; S.cval[p] = M.w_val[cbeg + p];
            if (i == pr) S.wc = p;
            else {
                const int b = M.lbeg[m + i], e = M.lend[m + i];
                S.rhb[p] = b; S.rhe[p] = e; S.rhc[p] = M.lcap[m + i];
                const int nz = e - b;
                grow += nz + rnz1 + slack_of(M.prm, nz + rnz1);
            }
        }
        for (int k = tid; k <= rnz1; k += NT) {
            const int j = M.w_idx[rbeg + k];
            S.ridx[k] = j;
            if (j == pc) S.wr = k;
            else {
                const int b = M.lbeg[j], e = M.lend[j];
                S.chb[k] = b; S.che[k] = e; S.chc[k] = M.lcap[j];
                const int nz = e - b;
                grow += nz + cnz1 + slack_of(M.prm, nz + cnz1);
            }
        }
        grow = block_sum64<NT>(grow, S.kscr);
        const int wc = S.wc, wr = S.wr;
        if (wc < 0 || wr < 0) { if (tid == 0) BLU_CHECK(S, 0); bsync<NT>(); return; }
        /* (block_sum64 ended with a barrier: the staged arrays are complete and visible) */
        if (tid == 0) {
            /* pivot to the front of its column and row (pivot.rs:142-154), headers travel along */
            int ti = S.cidx[0]; S.cidx[0] = S.cidx[wc]; S.cidx[wc] = ti;
            double tv = S.cval[0]; S.cval[0] = S.cval[wc]; S.cval[wc] = tv;
            S.rhb[wc] = S.rhb[0]; S.rhe[wc] = S.rhe[0]; S.rhc[wc] = S.rhc[0];
            ti = S.ridx[0]; S.ridx[0] = S.ridx[wr]; S.ridx[wr] = ti;
            S.chb[wr] = S.chb[0]; S.che[wr] = S.che[0]; S.chc[wr] = S.chc[0];
            S.flag_a = 0; S.flag_b = 0;
        }
        bsync<NT>();
        {
            const int ng0 = S.ngarbage;
            if (!w_reserve<NT>(S, grow)) return;
            if (S.ngarbage != ng0) {      /* the lines moved: reload the headers */
                for (int p = 1 + tid; p <= cnz1; p += NT) { const int i = S.cidx[p]; S.rhb[p] = M.lbeg[m + i]; S.rhe[p] = M.lend[m + i]; S.rhc[p] = M.lcap[m + i]; }
                for (int k = 1 + tid; k <= rnz1; k += NT) { const int j = S.ridx[k]; S.chb[k] = M.lbeg[j]; S.che[k] = M.lend[j]; S.chc[k] = M.lcap[j]; }
            }
        }
        pivot = S.cval[0];
        cidx = S.cidx; ridx = S.ridx; cval = S.cval;
        for (int p = 1 + tid; p <= cnz1; p += NT) S.rm[cidx[p]] = (unsigned short)p;
        for (int k = tid; k <= rnz1; k += NT) S.cm[ridx[k]] = 1;
        work = S.work + (size_t)wid * S.cap;
        for (int p = lane; p <= cnz1; p += 32) work[p] = 0.0;
        bsync<NT>();
    } else {
        /* prologue, pivot.rs:142-208: find the pivot in its column and row, bound the growth */
        int cend = M.lend[pc], rend = M.lend[m + pr];
        i64 grow = 0; int wc = -1, wr = -1;
        for (int pos = cbeg + tid; pos < cend; pos += NT) {
            int i = M.w_idx[pos];
            if (i == pr) wc = pos;
            else { int nz = M.lend[m + i] - M.lbeg[m + i]; grow += nz + rnz1 + slack_of(M.prm, nz + rnz1); }
        }
        for (int pos = rbeg + tid; pos < rend; pos += NT) {
            int j = M.w_idx[pos];
            if (j == pc) wr = pos;
            else { int nz = M.lend[j] - M.lbeg[j]; grow += nz + cnz1 + slack_of(M.prm, nz + cnz1); }
        }
        grow = block_sum64<NT>(grow, S.kscr);
        wc = block_max<NT>(wc, S.iscr);
        wr = block_max<NT>(wr, S.iscr);
        if (wc < 0 || wr < 0) { if (tid == 0) BLU_CHECK(S, 0); bsync<NT>(); return; }
        if (tid == 0) {
            int ti = M.w_idx[cbeg]; M.w_idx[cbeg] = M.w_idx[wc]; M.w_idx[wc] = ti;
            double tv = M.w_val[cbeg]; M.w_val[cbeg] = M.w_val[wc]; M.w_val[wc] = tv;
            ti = M.w_idx[rbeg]; M.w_idx[rbeg] = M.w_idx[wr]; M.w_idx[wr] = ti;
        }
        bsync<NT>();
        if (!w_reserve<NT>(S, grow)) return;
        cbeg = M.lbeg[pc]; rbeg = M.lbeg[m + pr];
        pivot = M.w_val[cbeg];
        /* stage the pivot column / row in shared memory when they fit */
        const bool ccached = cnz1 + 1 <= S.cap, rcached = rnz1 + 1 <= S.cap;
        cidx = ccached ? S.cidx : M.w_idx + cbeg;
        cval = ccached ? S.cval : M.w_val + cbeg;
        ridx = rcached ? S.ridx : M.w_idx + rbeg;
        if (ccached) for (int p = tid; p <= cnz1; p += NT) { S.cidx[p] = M.w_idx[cbeg + p]; S.cval[p] = M.w_val[cbeg + p]; }
        if (rcached) for (int k = tid; k <= rnz1; k += NT) S.ridx[k] = M.w_idx[rbeg + k];
        for (int p = 1 + tid; p <= cnz1; p += NT) M.rowmark[M.w_idx[cbeg + p]] = p;
        for (int k = tid; k <= rnz1; k += NT) M.colmark[M.w_idx[rbeg + k]] = 1;
        work = ccached ? S.work + (size_t)wid * S.cap : M.gwork + (size_t)wid * m;
        if (ccached) for (int p = lane; p <= cnz1; p += 32) work[p] = 0.0;
        if (tid == 0) { S.flag_a = 0; S.flag_b = 0; }
        bsync<NT>();
    }
    /* where a row / column mark and a line header come from */
    auto rowmark_of = [&](int i) -> int { return FAST ? (int)S.rm[i] : M.rowmark[i]; };
    auto colmark_of = [&](int j) -> int { return FAST ? (int)S.cm[j] : M.colmark[j]; };

    const int ubase = M.u_begin[rank];
    const i64 cbase = S.cstamp, rbase = S.rstamp;
    double acc_bytes = 0.0;

    /* column file update, pivot.rs:219-331 / 569-693: one warp per column of the pivot row */
    for (int k = 1 + wid; k <= rnz1; k += NW) {
        if (FAST && k + PF_DIST * NW <= rnz1) {      /* a later line of this warp: start pulling it in now */
            const int nb = S.chb[k + PF_DIST * NW], nn = S.che[k + PF_DIST * NW] - nb;
            warp_prefetch_l2(M.w_idx + nb, nn * 4);
            warp_prefetch_l2(M.w_val + nb, nn * 8);
        }
        const int j = ridx[k];
        int beg = FAST ? S.chb[k] : M.lbeg[j], end = FAST ? S.che[k] : M.lend[j], cap = FAST ? S.chc[k] : M.lcap[j];
        const int oldnz = end - beg;
        int put, where = -1;
        double cmx = 0.0, xrj;
        int nT;
        if (oldnz <= 32 * REGE) {
            /* register-resident path: every load of the line is issued before the first
             * dependent use, so one column costs ~2 memory round trips instead of 2 per chunk */
            int ei[REGE], emk[REGE], toff[REGE]; double ev[REGE];
            #pragma unroll
            for (int e = 0; e < REGE; e++) {
                int pos = beg + e * 32 + lane;
                bool valid = pos < end;
                ei[e] = valid ? M.w_idx[pos] : -1;
                ev[e] = valid ? M.w_val[pos] : 0.0;
            }
            #pragma unroll
            for (int e = 0; e < REGE; e++) emk[e] = (e * 32 < oldnz && ei[e] >= 0) ? rowmark_of(ei[e]) : -1;
            int tcount = 0; double myx = 0.0; int mine = 0;
            #pragma unroll
            for (int e = 0; e < REGE; e++) {
                if (e * 32 < oldnz) {
                    int isT = emk[e] == 0;
                    unsigned tm = __ballot_sync(FULLMASK, isT);
                    toff[e] = tcount + __popc(tm & lanemask_lt());
                    if (isT) { if (ei[e] == pr) { where = toff[e]; myx = ev[e]; mine = 1; } else { double a = fabs(ev[e]); if (a > cmx) cmx = a; } }
                    if (emk[e] > 0) work[emk[e]] = ev[e];
                    tcount += __popc(tm);
                } else toff[e] = 0;
            }
            unsigned hm = __ballot_sync(FULLMASK, mine);
            if (hm == 0) { if (lane == 0) BLU_CHECK(S, 0); continue; }
            const int hl = __ffs((int)hm) - 1;
            where = __shfl_sync(FULLMASK, where, hl);
            xrj = __shfl_sync(FULLMASK, myx, hl);
            nT = tcount;
            int dstb = beg + 1;
            if (cap - (beg + nT) < cnz1) {      /* the line moves to the end of the file */
                int room = cnz1 + slack_of(M.prm, nT + cnz1);
                int np = 0;
                if (lane == 0) { np = atomicAdd(&S.w_used, nT - 1 + room); atomicAdd(&S.nexpand, 1); }
                np = __shfl_sync(FULLMASK, np, 0);
                dstb = np; cap = np + nT - 1 + room;
            }
            /* T with its first element and the pivot-row element exchanged, first dropped */
            #pragma unroll
            for (int e = 0; e < REGE; e++) {
                if (emk[e] == 0) {
                    int t = toff[e];
                    if (t != where) {
                        int slot = t == 0 ? where - 1 : t - 1;
                        M.w_idx[dstb + slot] = ei[e]; M.w_val[dstb + slot] = ev[e];
                    }
                }
            }
            beg = dstb; put = dstb + nT - 1;
            __syncwarp();
        } else {
            put = beg;
            for (int base = beg; base < end; base += 32) {
                int pos = base + lane;
                int valid = pos < end;
                int i = valid ? M.w_idx[pos] : 0;
                double x = valid ? M.w_val[pos] : 0.0;
                int mk = valid ? rowmark_of(i) : 0;
                int isT = valid && mk == 0;
                if (valid && mk > 0) work[mk] = x;
                unsigned tm = __ballot_sync(FULLMASK, isT);
                int dst = put + __popc(tm & lanemask_lt());
                if (isT) { if (i == pr) where = dst; else { double a = fabs(x); if (a > cmx) cmx = a; } }
                __syncwarp();
                if (isT) { M.w_idx[dst] = i; M.w_val[dst] = x; }
                put += __popc(tm);
            }
            where = warp_max(where);
            __syncwarp();
            if (where < 0) { if (lane == 0) BLU_CHECK(S, 0); continue; }
            xrj = M.w_val[where];
            __syncwarp();
            if (lane == 0 && where != beg) { M.w_idx[where] = M.w_idx[beg]; M.w_val[where] = M.w_val[beg]; }
            __syncwarp();
            nT = put - beg;
            beg += 1;                            /* the pivot-row entry leaves the line */
            if (cap - put < cnz1) {              /* move the line to the end of the file */
                int nz = put - beg;
                int room = cnz1 + slack_of(M.prm, nT + cnz1);
                int np = 0;
                if (lane == 0) { np = atomicAdd(&S.w_used, nz + room); atomicAdd(&S.nexpand, 1); }
                np = __shfl_sync(FULLMASK, np, 0);
                for (int t = lane; t < nz; t += 32) { M.w_idx[np + t] = M.w_idx[beg + t]; M.w_val[np + t] = M.w_val[beg + t]; }
                beg = np; put = np + nz; cap = np + nz + room;
                __syncwarp();
            }
        }
        const double a = __ddiv_rn(xrj, pivot);
        u64 cmask = 0;
        for (int base = 1; base <= cnz1; base += 32) {
            int p = base + lane;
            int valid = p <= cnz1;
            double x = 0.0;
            if (valid) { x = __dsub_rn(work[p], __dmul_rn(a, cval[p])); work[p] = 0.0; }
            if (!small) {
                if (valid) {
                    M.w_idx[put + p - 1] = cidx[p]; M.w_val[put + p - 1] = x;
                    double ax = fabs(x); if (ax > cmx) cmx = ax;
                }
            } else {
                int keep = valid && fabs(x) > droptol;
                unsigned km = __ballot_sync(FULLMASK, keep);
                unsigned dm = __ballot_sync(FULLMASK, valid && !keep);
                if (keep) {
                    int d = put + __popc(km & lanemask_lt());
                    M.w_idx[d] = cidx[p]; M.w_val[d] = x;
                    double ax = fabs(x); if (ax > cmx) cmx = ax;
                }
                cmask |= (u64)dm << (base - 1);
                put += __popc(km);
            }
        }
        if (!small) put += cnz1;
        cmx = warp_maxd(cmx);
        if (lane == 0) {
            M.lbeg[j] = beg; M.lend[j] = put; M.lcap[j] = cap;
            M.colpiv[j] = cmx;
            M.ckey[j] = mkckey(put - beg, cbase + k, cmx, abstol);
            if (small) M.cancelled[k - 1] = cmask;
            if (fabs(xrj) > droptol) { M.u_idx[ubase + k - 1] = j; M.u_val[ubase + k - 1] = xrj; }
            else { M.u_idx[ubase + k - 1] = -2; M.u_val[ubase + k - 1] = 0.0; S.flag_a = 1; }
            if (cmx == 0.0 || cmx < abstol) S.need_remove = 1;
            acc_bytes += 12.0 * (oldnz + put - beg);
        }
        __syncwarp();
    }
    if (small) bsync<NT>();      /* the row update needs every column's cancellation mask */

    /* row file update, pivot.rs:335-401 / 697-774: one warp per row of the pivot column */
    for (int p = 1 + wid; p <= cnz1; p += NW) {
        if (FAST && p + PF_DIST * NW <= cnz1) warp_prefetch_l2(M.w_idx + S.rhb[p + PF_DIST * NW], (S.rhe[p + PF_DIST * NW] - S.rhb[p + PF_DIST * NW]) * 4);
        const int i = cidx[p];
        const int line = m + i;
        int beg = FAST ? S.rhb[p] : M.lbeg[line], end = FAST ? S.rhe[p] : M.lend[line], cap = FAST ? S.rhc[p] : M.lcap[line];
        const int oldnz = end - beg;
        int put;
        if (oldnz <= 32 * REGE) {
            int rj[REGE], roff[REGE];
            #pragma unroll
            for (int e = 0; e < REGE; e++) {
                int pos = beg + e * 32 + lane;
                rj[e] = pos < end ? M.w_idx[pos] : -1;
            }
            int kcount = 0;
            #pragma unroll
            for (int e = 0; e < REGE; e++) {
                roff[e] = -1;
                if (e * 32 < oldnz) {
                    int keep = rj[e] >= 0 && colmark_of(rj[e]) == 0;
                    unsigned km = __ballot_sync(FULLMASK, keep);
                    if (keep) roff[e] = kcount + __popc(km & lanemask_lt());
                    kcount += __popc(km);
                }
            }
            int dstb = beg;
            if (cap - (beg + kcount) < rnz1) {
                int room = rnz1 + slack_of(M.prm, kcount + rnz1);
                int np = 0;
                if (lane == 0) { np = atomicAdd(&S.w_used, kcount + room); atomicAdd(&S.nexpand, 1); }
                np = __shfl_sync(FULLMASK, np, 0);
                dstb = np; cap = np + kcount + room;
            }
            #pragma unroll
            for (int e = 0; e < REGE; e++) if (roff[e] >= 0) M.w_idx[dstb + roff[e]] = rj[e];
            beg = dstb; put = dstb + kcount;
            __syncwarp();
        } else {
            put = beg;
            for (int base = beg; base < end; base += 32) {
                int pos = base + lane;
                int valid = pos < end;
                int j = valid ? M.w_idx[pos] : 0;
                int keep = valid && colmark_of(j) == 0;
                unsigned km = __ballot_sync(FULLMASK, keep);
                __syncwarp();
                if (keep) M.w_idx[put + __popc(km & lanemask_lt())] = j;
                put += __popc(km);
            }
            __syncwarp();
            if (cap - put < rnz1) {
                int nz = put - beg;
                int room = rnz1 + slack_of(M.prm, nz + rnz1);
                int np = 0;
                if (lane == 0) { np = atomicAdd(&S.w_used, nz + room); atomicAdd(&S.nexpand, 1); }
                np = __shfl_sync(FULLMASK, np, 0);
                for (int t = lane; t < nz; t += 32) M.w_idx[np + t] = M.w_idx[beg + t];
                beg = np; put = np + nz; cap = np + nz + room;
                __syncwarp();
            }
        }
        if (!small) {
            for (int k = 1 + lane; k <= rnz1; k += 32) M.w_idx[put + k - 1] = ridx[k];
            put += rnz1;
        } else {
            for (int base = 1; base <= rnz1; base += 32) {
                int k = base + lane;
                int keep = k <= rnz1 && ((M.cancelled[k - 1] >> (p - 1)) & 1ull) == 0;
                unsigned km = __ballot_sync(FULLMASK, keep);
                if (keep) M.w_idx[put + __popc(km & lanemask_lt())] = ridx[k];
                put += __popc(km);
            }
        }
        if (lane == 0) {
            M.lbeg[line] = beg; M.lend[line] = put; M.lcap[line] = cap;
            M.rkey[i] = mkkey(put - beg, rbase + p);
            acc_bytes += 4.0 * (oldnz + put - beg);
        }
        __syncwarp();
    }

    /* L column, pivot.rs:403-415 (tentative slots, squeezed if something was dropped) */
    const int lbase = M.l_begin_p[rank];
    for (int p = 1 + tid; p <= cnz1; p += NT) {
        double x = __ddiv_rn(cval[p], pivot);
        if (fabs(x) > droptol) { M.l_idx[lbase + p - 1] = cidx[p]; M.l_val[lbase + p - 1] = x; }
        else { M.l_idx[lbase + p - 1] = -2; M.l_val[lbase + p - 1] = 0.0; S.flag_b = 1; }
    }
    if (acc_bytes != 0.0) atomicAdd(&S.elim_bytes, acc_bytes);
    bsync<NT>();
    /* clear marks */
    if (FAST) {
        for (int p = 1 + tid; p <= cnz1; p += NT) S.rm[cidx[p]] = 0;
        for (int k = tid; k <= rnz1; k += NT) S.cm[ridx[k]] = 0;
    } else {
        for (int p = 1 + tid; p <= cnz1; p += NT) M.rowmark[cidx[p]] = 0;
        for (int k = tid; k <= rnz1; k += NT) M.colmark[ridx[k]] = 0;
    }
    if (wid == 0) {
        int ln = cnz1, un = rnz1;
        if (S.flag_b) ln = warp_squeeze(M.l_idx, M.l_val, lbase, cnz1);
        if (S.flag_a) un = warp_squeeze(M.u_idx, M.u_val, ubase, rnz1);
        if (lane == 0) {
            M.l_idx[lbase + ln] = -1;
            finish_step(S, rank, lbase + ln + 1, ubase + un, pivot, cnz1 + 1, rnz1 + 1);
            S.cstamp = cbase + rnz1 + 1;
            S.rstamp = rbase + cnz1 + 1;
            if (FAST) { S.wc = -1; S.wr = -1; }
        }
    }
    bsync<NT>();
}
template <int NT> __device__ __noinline__ void pivot_general(Shm &S, const bool small) { pivot_general_t<NT, false>(S, small); }      /* (the rare variant: kept out of line so that it does not weigh on the register allocation of the hot loop) */
template <int NT> __device__ __forceinline__ void pivot_general_fast(Shm &S, const bool small) { pivot_general_t<NT, true>(S, small); }

/* ------------------------------------------------------------------ */
/* pivot_singleton_row, pivot.rs:835-926                               */
/* ------------------------------------------------------------------ */
template <int NT> __device__ void pivot_singleton_row(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int pc = S.pivot_col, pr = S.pivot_row, rank = S.rank;
    const double droptol = M.prm.droptol;
    const int cbeg = M.lbeg[pc], cend = M.lend[pc];
    const int n = cend - cbeg;
    int wc = -1;
    for (int pos = cbeg + tid; pos < cend; pos += NT) if (M.w_idx[pos] == pr) wc = pos;
    wc = block_max<NT>(wc, S.iscr);
    if (wc < 0) { if (tid == 0) BLU_CHECK(S, 0); bsync<NT>(); return; }
    const double pivot = M.w_val[wc];
    const int lbase = M.l_begin_p[rank];
    const i64 rbase = S.rstamp;
    if (tid == 0) S.flag_b = 0;
    bsync<NT>();
    /* L in column order, skipping the pivot: slot = position minus one behind the pivot */
    for (int pos = cbeg + tid; pos < cend; pos += NT) {
        if (pos == wc) continue;
        int slot = lbase + (pos - cbeg) - (pos > wc ? 1 : 0);
        double x = __ddiv_rn(M.w_val[pos], pivot);
        if (fabs(x) > droptol) { M.l_idx[slot] = M.w_idx[pos]; M.l_val[slot] = x; }
        else { M.l_idx[slot] = -2; M.l_val[slot] = 0.0; S.flag_b = 1; }
    }
    /* each row of the column loses the pivot column: move-last-into-hole, re-stamp in column order */
    double acc_bytes = 0.0;
    for (int q = wid; q < n; q += NW) {
        const int pos = cbeg + q;
        const int i = M.w_idx[pos];
        if (i == pr) continue;
        const int rb = M.lbeg[m + i], re = M.lend[m + i];
        int where = -1;
        for (int b = rb; b < re; b += 32) {
            int t = b + lane;
            int hit = t < re && M.w_idx[t] == pc;
            unsigned hm = __ballot_sync(FULLMASK, hit);
            if (hm) { where = b + __ffs((int)hm) - 1; break; }
        }
        if (lane == 0) {
            if (where < 0) { BLU_CHECK(S, 0); }
            else {
                M.w_idx[where] = M.w_idx[re - 1];
                M.lend[m + i] = re - 1;
                M.rkey[i] = mkkey(re - 1 - rb, rbase + q);
                acc_bytes += 4.0 * (2 * (re - rb) - 1);
            }
        }
        __syncwarp();
    }
    if (acc_bytes != 0.0) atomicAdd(&S.elim_bytes, acc_bytes);
    bsync<NT>();
    if (wid == 0) {
        int ln = n - 1;
        if (S.flag_b) ln = warp_squeeze(M.l_idx, M.l_val, lbase, n - 1);
        if (lane == 0) {
            M.l_idx[lbase + ln] = -1;
            finish_step(S, rank, lbase + ln + 1, M.u_begin[rank], pivot, n, 1);
            S.rstamp = rbase + n;
        }
    }
    bsync<NT>();
}

/* ------------------------------------------------------------------ */
/* pivot_singleton_col, pivot.rs:928-1025                              */
/* ------------------------------------------------------------------ */
template <int NT> __device__ void pivot_singleton_col(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int pc = S.pivot_col, pr = S.pivot_row, rank = S.rank;
    const double droptol = M.prm.droptol, abstol = M.prm.abstol;
    const int cbeg = M.lbeg[pc];
    const int rbeg = M.lbeg[m + pr], rend = M.lend[m + pr];
    const int n = rend - rbeg;
    const double pivot = M.w_val[cbeg];
    int wr = -1;
    for (int pos = rbeg + tid; pos < rend; pos += NT) if (M.w_idx[pos] == pc) wr = pos;
    wr = block_max<NT>(wr, S.iscr);
    if (wr < 0) { if (tid == 0) BLU_CHECK(S, 0); bsync<NT>(); return; }
    const int ubase = M.u_begin[rank];
    const i64 cbase = S.cstamp;
    if (tid == 0) S.flag_a = 0;
    bsync<NT>();
    double acc_bytes = 0.0;
    for (int q = wid; q < n; q += NW) {
        const int rpos = rbeg + q;
        if (rpos == wr) continue;
        const int j = M.w_idx[rpos];
        const int beg = M.lbeg[j], end = M.lend[j];
        int where = -1; double cmx = 0.0, xrj = 0.0;
        for (int b = beg; b < end; b += 32) {
            int t = b + lane;
            if (t < end) {
                double v = M.w_val[t];
                if (M.w_idx[t] == pr) { where = t; xrj = v; }
                else { double a = fabs(v); if (a > cmx) cmx = a; }
            }
        }
        where = warp_max(where);
        cmx = warp_maxd(cmx);
        if (where < 0) { if (lane == 0) BLU_CHECK(S, 0); continue; }
        xrj = M.w_val[where];
        __syncwarp();
        if (lane == 0) {
            int slot = ubase + q - (rpos > wr ? 1 : 0);
            if (fabs(xrj) > droptol) { M.u_idx[slot] = j; M.u_val[slot] = xrj; }
            else { M.u_idx[slot] = -2; M.u_val[slot] = 0.0; S.flag_a = 1; }
            M.w_idx[where] = M.w_idx[end - 1];
            M.w_val[where] = M.w_val[end - 1];
            M.lend[j] = end - 1;
            M.ckey[j] = mkckey(end - 1 - beg, cbase + q, cmx, abstol);
            M.colpiv[j] = cmx;
            if (cmx == 0.0 || cmx < abstol) S.need_remove = 1;
            acc_bytes += 12.0 * (2 * (end - beg) - 1);
        }
        __syncwarp();
    }
    if (acc_bytes != 0.0) atomicAdd(&S.elim_bytes, acc_bytes);
    bsync<NT>();
    if (wid == 0) {
        int un = n - 1;
        if (S.flag_a) un = warp_squeeze(M.u_idx, M.u_val, ubase, n - 1);
        if (lane == 0) {
            int lbase = M.l_begin_p[rank];
            M.l_idx[lbase] = -1;
            finish_step(S, rank, lbase + 1, ubase + un, pivot, 1, n);
            S.cstamp = cbase + n;
        }
    }
    bsync<NT>();
}

/* ------------------------------------------------------------------ */
/* pivot_doubleton_col, pivot.rs:1027-1331                             */
/* ------------------------------------------------------------------ */
template <int NT> __device__ void pivot_doubleton_col(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const int pc = S.pivot_col, pr = S.pivot_row, rank = S.rank;
    const double droptol = M.prm.droptol, abstol = M.prm.abstol;
    int cbeg = M.lbeg[pc];
    int rbeg = M.lbeg[m + pr], rend = M.lend[m + pr];
    const int rnz1 = rend - rbeg - 1;

    /* pivot to the front of column and row, pivot.rs:1068-1082 */
    int wr = -1;
    for (int pos = rbeg + tid; pos < rend; pos += NT) if (M.w_idx[pos] == pc) wr = pos;
    wr = block_max<NT>(wr, S.iscr);
    if (wr < 0) { if (tid == 0) BLU_CHECK(S, 0); bsync<NT>(); return; }
    if (tid == 0) {
        if (M.w_idx[cbeg] != pr) {
            int ti = M.w_idx[cbeg]; M.w_idx[cbeg] = M.w_idx[cbeg + 1]; M.w_idx[cbeg + 1] = ti;
            double tv = M.w_val[cbeg]; M.w_val[cbeg] = M.w_val[cbeg + 1]; M.w_val[cbeg + 1] = tv;
        }
        int ti = M.w_idx[rbeg]; M.w_idx[rbeg] = M.w_idx[wr]; M.w_idx[wr] = ti;
        S.flag_a = 0; S.flag_b = 0;
    }
    bsync<NT>();
    /* room for the other row, pivot.rs:1087-1111 */
    {
        const int orow = M.w_idx[cbeg + 1];
        int nz = M.lend[m + orow] - M.lbeg[m + orow];
        i64 grow = nz + rnz1 + slack_of(M.prm, nz + rnz1);
        if (!w_reserve<NT>(S, grow)) return;
    }
    cbeg = M.lbeg[pc]; rbeg = M.lbeg[m + pr]; rend = M.lend[m + pr];
    const double pivot = M.w_val[cbeg];
    const int other_row = M.w_idx[cbeg + 1];
    const double other_value = M.w_val[cbeg + 1];
    const double q_op = __ddiv_rn(other_value, pivot);
    const int ubase = M.u_begin[rank];
    const i64 cbase = S.cstamp;
    int *kind = M.tmpi;          /* per column of the pivot row: 1 = fill-in, 2 = cancelled */
    double acc_bytes = 0.0;

    /* column file update, pivot.rs:1115-1222 */
    for (int k = 1 + wid; k <= rnz1; k += NW) {
        const int j = M.w_idx[rbeg + k];
        const int beg = M.lbeg[j];
        int end = M.lend[j];
        const int oldnz = end - beg;
        int wp = -1, wo = -1; double cmx = 0.0;
        for (int b = beg; b < end; b += 32) {
            int t = b + lane;
            if (t < end) {
                int i = M.w_idx[t];
                if (i == pr) wp = t;
                else if (i == other_row) wo = t;
                else { double a = fabs(M.w_val[t]); if (a > cmx) cmx = a; }
            }
        }
        wp = warp_max(wp); wo = warp_max(wo); cmx = warp_maxd(cmx);
        if (wp < 0) { if (lane == 0) BLU_CHECK(S, 0); continue; }
        __syncwarp();
        if (lane == 0) {
            const double xrj = M.w_val[wp];
            if (fabs(xrj) > droptol) { M.u_idx[ubase + k - 1] = j; M.u_val[ubase + k - 1] = xrj; }
            else { M.u_idx[ubase + k - 1] = -2; M.u_val[ubase + k - 1] = 0.0; S.flag_a = 1; }
            int kd = 0;
            if (wo < 0) {
                /* fill-in takes the slot of the pivot-row entry: no re-bucketing */
                double x = __dmul_rn(-xrj, q_op);
                double xa = fabs(x);
                if (xa > droptol) {
                    M.w_idx[wp] = other_row; M.w_val[wp] = x;
                    kd = 1;
                    if (xa > cmx) cmx = xa;
                    const u64 kk = M.ckey[j] & ~KEY_PARK;      /* same list position, new maximum */
                    M.ckey[j] = (cmx == 0.0 || cmx < abstol) ? (kk | KEY_PARK) : kk;
                } else {
                    end--;
                    M.w_idx[wp] = M.w_idx[end]; M.w_val[wp] = M.w_val[end];
                    M.lend[j] = end;
                    M.ckey[j] = mkckey(end - beg, cbase + k, cmx, abstol);
                }
            } else {
                end--;
                M.w_idx[wp] = M.w_idx[end]; M.w_val[wp] = M.w_val[end];
                if (wo == end) wo = wp;
                double v = __dsub_rn(M.w_val[wo], __dmul_rn(xrj, q_op));
                M.w_val[wo] = v;
                double xa = fabs(v);
                if (xa <= droptol) {
                    end--;
                    M.w_idx[wo] = M.w_idx[end]; M.w_val[wo] = M.w_val[end];
                    kd = 2; S.flag_b = 1;
                } else if (xa > cmx) cmx = xa;
                M.lend[j] = end;
                M.ckey[j] = mkckey(end - beg, cbase + k, cmx, abstol);
            }
            kind[k] = kd;
            M.colpiv[j] = cmx;
            if (cmx == 0.0 || cmx < abstol) S.need_remove = 1;
            acc_bytes += 12.0 * (oldnz + end - beg);
        }
        __syncwarp();
    }
    if (acc_bytes != 0.0) atomicAdd(&S.elim_bytes, acc_bytes);
    bsync<NT>();

    /* row file update of the other row, pivot.rs:1228-1293; warp 0 */
    if (wid == 0) {
        const int line = m + other_row;
        int beg = M.lbeg[line], end = M.lend[line], cap = M.lcap[line];
        const int oldnz = end - beg;
        if (S.flag_b) {
            /* order-preserving removal of the pivot column and of cancelled columns */
            for (int k = 1 + lane; k <= rnz1; k += 32) if (kind[k] == 2) M.colmark[M.w_idx[rbeg + k]] = 1;
            if (lane == 0) M.colmark[pc] = 1;
            __syncwarp();
            int put = beg;
            for (int b = beg; b < end; b += 32) {
                int t = b + lane;
                int valid = t < end;
                int j = valid ? M.w_idx[t] : 0;
                int keep = valid && M.colmark[j] == 0;
                unsigned km = __ballot_sync(FULLMASK, keep);
                __syncwarp();
                if (keep) M.w_idx[put + __popc(km & lanemask_lt())] = j;
                put += __popc(km);
            }
            __syncwarp();
            for (int k = 1 + lane; k <= rnz1; k += 32) if (kind[k] == 2) M.colmark[M.w_idx[rbeg + k]] = 0;
            if (lane == 0) M.colmark[pc] = 0;
            end = put;
            __syncwarp();
        } else {
            int where = -1;
            for (int b = beg; b < end; b += 32) {
                int t = b + lane;
                int hit = t < end && M.w_idx[t] == pc;
                unsigned hm = __ballot_sync(FULLMASK, hit);
                if (hm) { where = b + __ffs((int)hm) - 1; break; }
            }
            if (where < 0) { if (lane == 0) BLU_CHECK(S, 0); }
            else { if (lane == 0) M.w_idx[where] = M.w_idx[end - 1]; end--; }
            __syncwarp();
        }
        /* count fill-in columns, make room, append them in pivot-row order */
        int nfill = 0;
        for (int k = 1 + lane; k <= rnz1; k += 32) nfill += kind[k] == 1;
        nfill = warp_sum(nfill);
        if (nfill > cap - end) {
            int nz = end - beg;
            int room = nfill + slack_of(M.prm, nz + nfill);
            int np = 0;
            if (lane == 0) { np = atomicAdd(&S.w_used, nz + room); S.nexpand++; }
            np = __shfl_sync(FULLMASK, np, 0);
            for (int t = lane; t < nz; t += 32) M.w_idx[np + t] = M.w_idx[beg + t];
            beg = np; end = np + nz; cap = np + nz + room;
            __syncwarp();
        }
        for (int b = 1; b <= rnz1; b += 32) {
            int k = b + lane;
            int f = k <= rnz1 && kind[k] == 1;
            unsigned fm = __ballot_sync(FULLMASK, f);
            if (f) M.w_idx[end + __popc(fm & lanemask_lt())] = M.w_idx[rbeg + k];
            end += __popc(fm);
        }
        __syncwarp();
        int un = rnz1;
        if (S.flag_a) un = warp_squeeze(M.u_idx, M.u_val, ubase, rnz1);
        if (lane == 0) {
            M.lbeg[line] = beg; M.lend[line] = end; M.lcap[line] = cap;
            M.rkey[other_row] = mkkey(end - beg, S.rstamp);
            S.rstamp += 1;
            S.elim_bytes += 4.0 * (oldnz + end - beg);
            /* L column, pivot.rs:1296-1305 */
            int lput = M.l_begin_p[rank];
            if (fabs(q_op) > droptol) { M.l_idx[lput] = other_row; M.l_val[lput] = q_op; lput++; }
            M.l_idx[lput++] = -1;
            finish_step(S, rank, lput, ubase + un, pivot, 2, rnz1 + 1);
            S.cstamp = cbase + rnz1 + 1;
        }
    }
    bsync<NT>();
}

#include "blu_factor_dense.cuh"

/* ------------------------------------------------------------------ */
/* pivot dispatcher (pivot.rs:48-112) + driver (factorize_bump.rs)     */
/* ------------------------------------------------------------------ */
#define DENSE_MIN_ROWS 4      /* the last pivots are singletons / doubletons: not worth a conversion */
#define DENSE_MAX_ENTRIES 8   /* conversions per factorization (each costs O(kd^2)) */
/* dense-tail dispatch: RES (values in shared memory) is a launch property, uniform over the block */
template <int NT> __device__ __forceinline__ void dense_enter_d(Shm &S) {
    if (S.dv_smem == 1) dense_enter<NT, 1>(S); else if (S.dv_smem == 2) dense_enter<NT, 2>(S); else dense_enter<NT, 0>(S);
}
template <int NT> __device__ __forceinline__ void dense_exit_d(Shm &S) {
    if (S.dv_smem == 1) dense_exit<NT, 1>(S); else if (S.dv_smem == 2) dense_exit<NT, 2>(S); else dense_exit<NT, 0>(S);
}
template <int NT> __device__ __forceinline__ int dense_run_d(Shm &S) {
    return S.dv_smem == 1 ? dense_run<NT, 1>(S) : S.dv_smem == 2 ? dense_run<NT, 2>(S) : dense_run<NT, 0>(S);
}

template <int NT> __device__ void phase_bump(Shm &S) {
    Mat &M = S.M;
    const int m = M.m, tid = threadIdx.x;
    if (tid == 0) {
        S.lput = M.l_begin_p[S.rank]; S.uput = M.u_begin[S.rank];
        int ms = M.prm.maxsearch < 1 ? 1 : M.prm.maxsearch, lv = 0;
        for (int n = m; n > 32 && lv < 7; lv++) n = (n + 31) / 32;
        const int rows = 2 + ms * (lv + 1);
        S.use_tree = S.nact > M.tree_min && ms <= TREE_MAX_SEARCH && M.prm.search_rows == 0 && S.dyn_bytes >= rows * (32 * 12 + 4);
    }
    bsync<NT>();
    if (S.use_tree) ctree_build<NT>(S);
    while (S.rank + S.rankdef < m) {
        i64 t0 = clock64();
        const int kd_first = S.kd_big > S.kd_small ? S.kd_big : S.kd_small;
        if (!S.dense && kd_first > 0 && m - S.rank <= kd_first && m - S.rank >= DENSE_MIN_ROWS &&
            S.rank >= S.dense_block_rank && S.dense_entries < DENSE_MAX_ENTRIES && M.prm.search_rows == 0) {
            if (S.mode == BLU_MODE_HEAD) {      /* the tail kernel takes over from here */
                bsync<NT>();
                if (tid == 0) S.suspend = 1;
                bsync<NT>();
                return;
            }
            /* a two-stage tail starts in HBM/L2 at order kd_big and moves to shared memory at kd_small */
            bsync<NT>();
            if (tid == 0) {
                if (S.kd_big > S.kd_small && m - S.rank > S.kd_small) { S.kd = S.kd_big; S.dv_smem = 2; }
                else { S.kd = S.kd_small; S.dv_smem = S.launch_res; }
                S.kw = S.kd / 32;
                M.dn_key = S.kd == S.kd_small && S.kd_big > S.kd_small ? M.dn_key2 : M.dn_key;
            }
            bsync<NT>();
            dense_enter_d<NT>(S);
            if (tid == 0) { S.t_phase[13] += clock64() - t0; S.use_tree = 0; }      /* at most dense_k columns are left: the scan is cheap */
            if (S.status != BLU_OK) return;
            t0 = clock64();
        }
        if (S.dense) {
            /* the dense pivot loop runs until the pivots are found or a step belongs to the sparse code */
            const int code = dense_run_d<NT>(S);
            if (S.status != BLU_OK) return;
            if (code == DRUN_DONE) continue;
            if (code == DRUN_RESTAGE) {
                t0 = clock64();
                dense_restage<NT>(S);
                if (tid == 0) S.t_phase[13] += clock64() - t0;
                if (S.status != BLU_OK) return;
                continue;
            }
            t0 = clock64();
            dense_exit_d<NT>(S);
            if (tid == 0) S.t_phase[13] += clock64() - t0;
            if (S.status != BLU_OK) return;
            if (code == DRUN_REMOVE) {      /* pivot.rs:96-106 works on the line file */
                t0 = clock64();
                post_remove_cols<NT>(S, S.rank - 1);
                if (tid == 0) S.t_phase[10] += clock64() - t0;
                if (S.status != BLU_OK) return;
                continue;
            }
            /* DRUN_SPARSE_PIVOT: singleton row / column, doubleton column: the same pivot on the line file
             * (the room in L and U was checked by dense_run) */
        } else {
            if (M.prm.search_rows != 0) markowitz_search_rows<NT>(S); else markowitz_search<NT>(S);
            if (tid == 0) S.t_phase[3] += clock64() - t0;
            if (S.status != BLU_OK) return;
            if (S.pivot_row < 0) {
                /* empty column: drop it, no pivot (factorize_bump.rs:23-31) */
                const int pc0 = S.pivot_col;
                bsync<NT>();
                if (tid == 0) { M.ckey[pc0] = KEY_INF; S.ndead++; S.rankdef++; }
                bsync<NT>();
                if (S.use_tree) ctree_update<NT>(S, (const int *)0, 0, pc0);
                continue;
            }
        }
        const int pc = S.pivot_col, pr = S.pivot_row;
        const int rank = S.rank;
        const int nz_col = M.lend[pc] - M.lbeg[pc];
        const int nz_row = M.lend[m + pr] - M.lbeg[m + pr];
        /* room in L and U, pivot.rs:69-81 (S.lput / S.uput mirror l_begin_p[rank] / u_begin[rank]) */
        {
            int room = M.l_mem - S.lput;
            int st = BLU_OK;
            if (room < nz_col) { if (tid == 0) M.info->addmem_l = nz_col - room; st = BLU_REALLOCATE; }
            room = M.u_mem - S.uput;
            if (room < nz_row - 1) { if (tid == 0) M.info->addmem_u = nz_row - 1 - room; st = BLU_REALLOCATE; }
            if (st != BLU_OK) { bsync<NT>(); if (tid == 0) S.status = st; bsync<NT>(); return; }
        }
        t0 = clock64();
        int kind;
        if (nz_row == 1) { pivot_singleton_row<NT>(S); kind = 0; }
        else if (nz_col == 1) { pivot_singleton_col<NT>(S); kind = 1; }
        else if (nz_col == 2) { pivot_doubleton_col<NT>(S); kind = 2; }
        else {
            const bool small = nz_col - 1 <= MAXROW_SMALL;
            if (S.smarks && nz_col <= S.cap && nz_row <= S.cap) pivot_general_fast<NT>(S, small);
            else pivot_general<NT>(S, small);
            kind = small ? 3 : 4;
        }
        if (tid == 0) { S.t_phase[4 + kind] += clock64() - t0; S.n_kind[kind]++; }
        if (S.status != BLU_OK) return;
        t0 = clock64();
        post_remove_cols<NT>(S, rank);
        if (tid == 0) S.t_phase[10] += clock64() - t0;
        if (S.status != BLU_OK) return;
        if (S.use_tree) {
            /* the keys that changed belong to the columns of the pivot row, whose (unlinked) line still sits in W */
            t0 = clock64();
            ctree_update<NT>(S, M.w_idx + M.lbeg[m + pr], nz_row, -1);
            if (tid == 0) S.t_phase[3] += clock64() - t0;
        }
    }
}

#endif
