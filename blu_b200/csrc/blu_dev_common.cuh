/* blu_dev_common.cuh -- per-slot view, shared-memory state and block/warp primitives. */
#ifndef BLU_DEV_COMMON_CUH
#define BLU_DEV_COMMON_CUH

#include "blu_types.h"

#ifndef BLU_EMU
#include <cuda_runtime.h>
#define BLU_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define BLU_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

typedef blu_u64 u64;
typedef blu_i64 i64;

#ifdef BLU_EMU
static inline long long clock64() { return 0; }
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
#endif
#define FULLMASK 0xffffffffu
#define KEY_INF 0xffffffffffffffffull
#define KEY_PARK 0x8000000000000000ull   /* see mkckey */
#define STAMP_BITS 40
#define MAXROW_SMALL 64 /* pivot.rs:22 */
#define MAXCAND 32      /* upper bound on maxsearch honoured by the search kernel */
#define RING_BARS 128       /* completion barriers of the row ring (warps x buffers per warp), in front of the ring */
#define TREE_MAX_SEARCH 4    /* ... when maxsearch is at most this (the frontier of the top-k walk lives in shared memory) */

/* per-slot view of BluDev */
struct Mat {
    int m;
    int l_mem, u_mem, w_mem, bnz_cap, tree_min;
    i64 b_total;
    BluParams prm;
    const i64 *b_begin, *b_end, *b_i;
    const double *b_x;
    int *bt_ptr, *bt_idx; double *bt_val;
    int *pinv, *qinv, *prank, *qrank;
    double *colpiv, *rowpiv;
    int *l_idx; double *l_val;
    int *u_idx; double *u_val;
    int *w_idx; double *w_val;
    int *lbeg, *lend, *lcap;
    u64 *ckey, *rkey, *ctree;
    int *l_begin_p, *u_begin, *l_begin, *lt_begin, *lt_begin_p, *p, *r_begin, *eta_row;
    int *dep_lt, *dep_lc, *dep_uc, *len_uc;
    int *ur_ptr, *ur_idx, *dep_ur; double *ur_val;
    int *pivotcol, *pivotrow;
    int *rowmark, *colmark, *marked, *iwork1, *pstack, *acols, *tmpi;
    u64 *cancelled;
    double *work0, *work1, *gwork;
    double *dn_val; BluKey2 *dn_key, *dn_key2; unsigned *dn_rbits, *dn_cbits;
    BluInfo *info;
};

__device__ __forceinline__ void mat_view(Mat &M, const BluDev &D, int s) {
    const size_t m = (size_t)D.m, S = (size_t)s;
    M.m = D.m;
    M.l_mem = (int)D.l_mem; M.u_mem = (int)D.u_mem; M.w_mem = (int)D.w_mem; M.bnz_cap = (int)D.bnz_cap;
    M.prm = D.prm; M.b_total = D.b_total; M.tree_min = D.tree_min;
    M.b_begin = D.b_begin + S * m; M.b_end = D.b_end + S * m; M.b_i = D.b_i; M.b_x = D.b_x;
    M.bt_ptr = D.bt_ptr + S * (m + 1);
    M.bt_idx = D.bt_idx + S * (size_t)D.bnz_cap; M.bt_val = D.bt_val + S * (size_t)D.bnz_cap;
    M.pinv = D.pinv + S * m; M.qinv = D.qinv + S * m;
    M.prank = D.prank + S * m; M.qrank = D.qrank + S * m;
    M.colpiv = D.colpiv + S * m; M.rowpiv = D.rowpiv + S * m;
    M.l_idx = D.l_idx + S * (size_t)D.l_mem; M.l_val = D.l_val + S * (size_t)D.l_mem;
    M.u_idx = D.u_idx + S * (size_t)D.u_mem; M.u_val = D.u_val + S * (size_t)D.u_mem;
    M.w_idx = D.w_idx + S * 2 * (size_t)D.w_mem; M.w_val = D.w_val + S * 2 * (size_t)D.w_mem;
    if (D.slot_store) {      /* private stores of a basis that outgrew the uniform ones */
        const BluSlotStore &e = D.slot_store[s];
        if (e.l_idx) { M.l_idx = e.l_idx; M.l_val = e.l_val; M.l_mem = (int)e.l_mem; }
        if (e.u_idx) { M.u_idx = e.u_idx; M.u_val = e.u_val; M.u_mem = (int)e.u_mem; }
        if (e.w_idx) { M.w_idx = e.w_idx; M.w_val = e.w_val; M.w_mem = (int)e.w_mem; }
    }
    M.lbeg = D.lbeg + S * 2 * m; M.lend = D.lend + S * 2 * m; M.lcap = D.lcap + S * 2 * m;
    M.ckey = D.ckey + S * m; M.rkey = D.rkey + S * m;
    M.ctree = D.ctree + S * (m / 31 + 72);
    M.l_begin_p = D.l_begin_p + S * (m + 1); M.u_begin = D.u_begin + S * (m + 1);
    M.l_begin = D.l_begin + S * (m + 1); M.lt_begin = D.lt_begin + S * (m + 1);
    M.lt_begin_p = D.lt_begin_p + S * (m + 1); M.p = D.p + S * (m + 1);
    M.r_begin = D.r_begin + S * (m + 1); M.eta_row = D.eta_row + S * (m + 1);
    M.dep_lt = D.dep_lt + S * m; M.dep_lc = D.dep_lc + S * m; M.dep_uc = D.dep_uc + S * m; M.len_uc = D.len_uc + S * m;
    M.ur_ptr = D.ur_ptr; M.ur_idx = D.ur_idx; M.dep_ur = D.dep_ur; M.ur_val = D.ur_val;   /* single object: slot 0 */
    M.pivotcol = D.pivotcol + S * (2 * m + 2); M.pivotrow = D.pivotrow + S * (2 * m + 2);
    M.rowmark = D.rowmark + S * m; M.colmark = D.colmark + S * m; M.marked = D.marked + S * m;
    M.iwork1 = D.iwork1 + S * (2 * m + 2); M.pstack = D.pstack + S * m;
    M.acols = D.acols + S * m; M.tmpi = D.tmpi + S * (4 * m + 4);
    M.cancelled = D.cancelled + S * m;
    M.work0 = D.work0 + S * m; M.work1 = D.work1 + S * m;
    M.gwork = D.gwork + S * (size_t)D.gwork_warps * m;
    {
        const size_t k2 = (size_t)D.dense_k, kd = D.dense_kbig > D.dense_k ? (size_t)D.dense_kbig : k2, kw = kd / 32;
        M.dn_val = D.dn_val + S * kd * kd; M.dn_key = D.dn_key + S * kd * kd;
        M.dn_key2 = D.dn_key2 + (D.dense_kbig > D.dense_k ? S * k2 * k2 : 0);
        M.dn_rbits = D.dn_rbits + S * kd * kw; M.dn_cbits = D.dn_cbits + S * kd * kw;
    }
    M.info = D.info + s;
}

/* block-shared state of one factorization */
struct Shm {
    Mat M;
    int iscr[40];
    u64 kscr[40];
    double dscr[40];
    int status;
    int rank, rankdef;
    int w_used, w_limit;      /* fill pointer and end of the live W half */
    int w_half;
    i64 cstamp, rstamp;
    int nact, ndead;
    int pivot_row, pivot_col;
    int flag_a, flag_b, need_remove;
    int nexpand, ngarbage;
    i64 nsearch;
    i64 factor_flops;
    int cand_col[MAXCAND];
    i64 cand_mc[MAXCAND];
    int cand_row[MAXCAND];
    int ncand;
    int cap;                  /* entries of the smem line caches */
    int *cidx, *ridx; double *cval, *work; /* dynamic smem carve-up */
    /* fast path (m <= SMARK_MAX): line headers of the pivot row's columns / pivot column's rows,
     * and the row/column marks, all in shared memory */
    int *chb, *che, *chc, *rhb, *rhe, *rhc;
    unsigned short *rm; unsigned char *cm;
    int smarks;               /* the shared-memory marks exist */
    int dyn_bytes;            /* size of the dynamic shared memory (reused as scratch outside the pivot loop) */
    int wc, wr;               /* position of the pivot in its column / row (single writer) */
    double elim_bytes; i64 nelim_div;
    i64 t_phase[16]; i64 n_kind[8];
    /* dense tail (blu_factor_dense.cuh) */
    unsigned char *dyn;       /* base of the dynamic shared memory */
    int dense;                /* 1 while the active submatrix lives in the dense arrays */
    int kd, kw;               /* slots per side (multiple of 32) and words per bitmap row */
    int nrs, ncs;             /* row / column slots handed out at entry */
    int dpt, dpc;             /* slots of the pivot row / column */
    unsigned epoch;           /* number of the next dense step (storage-order keys, BluKey2) */
    int dense_entries, dense_block_rank;
    int mode, suspend;        /* BLU_MODE_*; 1 = park for the tail kernel, 2 = park for the build kernel */
    int dv_smem;              /* where the dense arrays of the current stage live: 0 = HBM, 1 = values and bitmaps in shared memory,
                               * 2 = bitmaps in shared memory, values in HBM/L2 (the first stage of a two-stage tail) */
    int launch_res;           /* the launch has room for the dense values and bitmaps of order kd_small in shared memory */
    int kd_small, kd_big;     /* orders of the (last) stage and of the HBM/L2 stage in front of it (0: none) */
    int use_tree, tree_levels, tree_off[8], tree_n[8];   /* min-tree over the column keys (markowitz_search of large bumps) */
    u64 mbar; unsigned mbar_phase;   /* completion barrier of the bulk copies (dense_pivot) and its current phase */
    int lput, uput;           /* fill pointers of L and U (= l_begin_p[rank], u_begin[rank]) */
    int dpcand;               /* stash row of the pivot column's keys */
    /* first stage of a two-stage tail: per-warp ring of row buffers filled and drained by bulk copies (dense_step) */
    unsigned ring_phase[32];  /* per warp: current phase bit of each of its buffers */
    int ring_nbuf;            /* buffers per warp in this stage (0: no ring) */
    i64 dstamp_base;          /* dense tail: stamp - dstamp_base is the 23-bit stamp of the 32-bit search keys (skey32) */
    int cand_done;            /* candidate columns evaluated so far (the warp that finishes last makes the choice) */
    int dgeneral;             /* the chosen pivot is a general dense step (pivot_any / pivot_small) */
};

template <int NT> __device__ __forceinline__ void bsync() {
    if (NT > 32) __syncthreads(); else __syncwarp();
}
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }

__device__ __forceinline__ u64 mkkey(int cnt, i64 stamp) { return ((u64)(unsigned)cnt << STAMP_BITS) | (u64)stamp; }
__device__ __forceinline__ int key_cnt(u64 k) { return (int)((k & ~KEY_PARK) >> STAMP_BITS); }
/* Column key.  A non-empty column whose maximum fell below abstol without the column being emptied (its
 * pivot-row entry was dropped from U, so pivot.rs:96-106 never looked at it) stays in its count list but
 * the search passes over it without counting it (markowitz.rs:88-90 with D6 repaired): the top bit hides
 * it from the candidate scan until an update rewrites the key. */
__device__ __forceinline__ u64 mkckey(int cnt, i64 stamp, double cmx, double abstol) {
    const u64 k = mkkey(cnt, stamp);
    return (cnt > 0 && (cmx == 0.0 || cmx < abstol)) ? (k | KEY_PARK) : k;
}

/* ---- warp primitives ---- */
__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULLMASK, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULLMASK, v, d);
    return v;
}
__device__ __forceinline__ i64 warp_sum64(i64 v) {
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULLMASK, v, d);
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) { int t = __shfl_xor_sync(FULLMASK, v, d); v = t > v ? t : v; }
    return v;
}
__device__ __forceinline__ int warp_min(int v) {
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) { int t = __shfl_xor_sync(FULLMASK, v, d); v = t < v ? t : v; }
    return v;
}
__device__ __forceinline__ double warp_maxd(double v) {
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) { double t = __shfl_xor_sync(FULLMASK, v, d); v = t > v ? t : v; }
    return v;
}
__device__ __forceinline__ double warp_mind(double v) {
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) { double t = __shfl_xor_sync(FULLMASK, v, d); v = t < v ? t : v; }
    return v;
}
#ifndef BLU_EMU
__device__ __forceinline__ unsigned warp_min_u32(unsigned v) { return __reduce_min_sync(FULLMASK, v); }   /* REDUX */
#else
static inline unsigned warp_min_u32(unsigned v) {
    for (int d = 16; d > 0; d >>= 1) { unsigned t = __shfl_xor_sync(FULLMASK, v, d); v = t < v ? t : v; }
    return v;
}
#endif
__device__ __forceinline__ u64 warp_min64(u64 v) {
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) { u64 t = __shfl_xor_sync(FULLMASK, v, d); v = t < v ? t : v; }
    return v;
}
/* sum of doubles in lane order is not needed bit-exactly anywhere in the factorization */
__device__ __forceinline__ double warp_sumd(double v) {
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULLMASK, v, d);
    return v;
}

/* ---- block primitives (every thread of the block must call them) ---- */
template <int NT> __device__ __forceinline__ int block_excl_scan(int v, int *total, int *scr) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = warp_incl_scan(v);
    if (NW == 1) {
        *total = __shfl_sync(FULLMASK, incl, 31);
        return incl - v;
    }
    if (lane == 31) scr[wid] = incl;
    bsync<NT>();
    if (wid == 0) {
        int t = lane < NW ? scr[lane] : 0;
        int ti = warp_incl_scan(t);
        if (lane < NW) scr[lane] = ti - t;
        if (lane == NW - 1) scr[NW] = ti;
    }
    bsync<NT>();
    int res = incl - v + scr[wid];
    *total = scr[NW];
    bsync<NT>();
    return res;
}
template <int NT> __device__ __forceinline__ i64 block_sum64(i64 v, u64 *scr) {
    constexpr int NW = NT / 32;
    v = warp_sum64(v);
    if (NW == 1) return v;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) scr[wid] = (u64)v;
    bsync<NT>();
    i64 r = 0;
    #pragma unroll
    for (int w = 0; w < NW; w++) r += (i64)scr[w];
    bsync<NT>();
    return r;
}
template <int NT> __device__ __forceinline__ int block_max(int v, int *scr) {
    constexpr int NW = NT / 32;
    v = warp_max(v);
    if (NW == 1) return v;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) scr[wid] = v;
    bsync<NT>();
    int r = scr[0];
    #pragma unroll
    for (int w = 1; w < NW; w++) r = scr[w] > r ? scr[w] : r;
    bsync<NT>();
    return r;
}
template <int NT> __device__ __forceinline__ u64 block_min64(u64 v, u64 *scr) {
    constexpr int NW = NT / 32;
    v = warp_min64(v);
    if (NW == 1) return v;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) scr[wid] = v;
    bsync<NT>();
    u64 r = scr[0];
    #pragma unroll
    for (int w = 1; w < NW; w++) r = scr[w] < r ? scr[w] : r;
    bsync<NT>();
    return r;
}
template <int NT> __device__ __forceinline__ double block_maxd(double v, double *scr) {
    constexpr int NW = NT / 32;
    v = warp_maxd(v);
    if (NW == 1) return v;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) scr[wid] = v;
    bsync<NT>();
    double r = scr[0];
    #pragma unroll
    for (int w = 1; w < NW; w++) r = scr[w] > r ? scr[w] : r;
    bsync<NT>();
    return r;
}
template <int NT> __device__ __forceinline__ double block_mind(double v, double *scr) {
    constexpr int NW = NT / 32;
    v = warp_mind(v);
    if (NW == 1) return v;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) scr[wid] = v;
    bsync<NT>();
    double r = scr[0];
    #pragma unroll
    for (int w = 1; w < NW; w++) r = scr[w] < r ? scr[w] : r;
    bsync<NT>();
    return r;
}

/* pull [p, p+bytes) towards the SM ahead of use: one 128-byte line per lane and round (L2 prefetch) */
__device__ __forceinline__ void warp_prefetch_l2(const void *p, int bytes) {
#ifndef BLU_EMU
    const char *c = (const char *)p;
    for (int off = (threadIdx.x & 31) * 128; off < bytes; off += 32 * 128)
#ifdef PF_L1
        asm volatile("prefetch.global.L1 [%0];" ::"l"(c + off));
#else
        asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
#endif
#else
    (void)p; (void)bytes;
#endif
}

/* one address per lane: start pulling the line that holds *p into L1 (used by the sequential sweeps:
 * the 32 line starts of the next 32 pivots are requested together, long before they are needed) */
__device__ __forceinline__ void lane_prefetch(const void *p) {
#ifndef BLU_EMU
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

/* ---- bulk asynchronous copy global -> shared (the TMA engine's 1-D form, cp.async.bulk, completion on an mbarrier):
 * one thread starts the copy of a contiguous, 16-byte aligned segment and the block goes on with other work; whoever
 * needs the data waits for the barrier's phase.  SASS: UBLKCP + SYNCS.  (The emulator copies in place.) ---- */
#ifndef BLU_EMU
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_inval(u64 *bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void bulk_copy_g2s(void *sdst, const void *gsrc, unsigned bytes, u64 *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
/* writes of the generic proxy (ordered before by a barrier) become visible to the bulk-copy engine */
__device__ __forceinline__ void bulk_fence() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(u64 *bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
/* shared -> global bulk copy (completion tracked per thread in bulk async-groups) */
__device__ __forceinline__ void bulk_copy_s2g(void *gdst, const void *ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
/* all but the N most recent groups have read their shared-memory source (the buffer may be refilled) */
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
/* all but the N most recent groups are complete (their writes performed) */
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
/* this thread's shared-memory writes become visible to the bulk-copy engine */
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
/* values that bulk stores rewrite are read past L1 (the async proxy does not keep L1 coherent) */
__device__ __forceinline__ double ld_l2(const double *p) { return __ldcg(p); }
#else
static inline void bulk_copy_s2g(void *gdst, const void *ssrc, unsigned bytes) { memcpy(gdst, ssrc, bytes); }
static inline void bulk_commit() {}
template <int N> static inline void bulk_wait_read() {}
template <int N> static inline void bulk_wait() {}
static inline void fence_async_smem() {}
static inline double ld_l2(const double *p) { return *p; }
/* emulated completion barrier: the word counts completed phases; a waiter yields until the phase of its parity is over */
static inline void mbar_init(u64 *bar, unsigned) { *bar = 0; }
static inline void mbar_inval(u64 *) {}
static inline void bulk_copy_g2s(void *sdst, const void *gsrc, unsigned bytes, u64 *bar) { memcpy(sdst, gsrc, bytes); (*(unsigned *)bar)++; }
static inline void mbar_wait(u64 *bar, unsigned parity) { unsigned *g = (unsigned *)bar; while ((*g & 1u) == parity) emu::yield_wait(g, *g); }
static inline void bulk_fence() {}
#endif

/* record the first failed device-side invariant (kept live like the reference's assert!s) */
#define BLU_CHECK(S, cond) do { if (!(cond)) { if ((S).M.info->internal_error == 0) (S).M.info->internal_error = __LINE__; (S).status = BLU_ERROR_INTERNAL; } } while (0)

__device__ __forceinline__ int slack_of(const BluParams &p, int nz) { return (int)(p.stretch * (double)nz) + p.pad; }

#endif
