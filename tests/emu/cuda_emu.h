/* cuda_emu.h -- a tiny SIMT emulator so that the CUDA kernels of blu_b200/csrc can be
 * compiled with g++ and stepped through on a machine without a GPU.
 *
 * TEST / DEBUG INFRASTRUCTURE ONLY.  The product library (libblu_b200.so) is built by
 * nvcc for sm_100a and never contains, loads or falls back to this code; the emulated
 * build is a separate library (tests/emu/libblu_emu.so) that only `-m "not gpu"` tests
 * load, to check kernel LOGIC (barrier placement, ordered compaction, list surgery)
 * against the oracle on the CPU-only build box.
 *
 * Model: one CUDA block = NT fibers (ucontext) run round-robin by one OS thread;
 * a fiber runs until it reaches __syncthreads / __syncwarp / a *_sync warp primitive.
 * Blocks of a grid run one after another.  `__shared__` becomes `static` (valid
 * because only one block is alive at a time).  Divergent barriers deadlock in the
 * scheduler and abort with a message instead of hanging.
 */
#ifndef CUDA_EMU_H
#define CUDA_EMU_H
#include <ucontext.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <functional>
#include <vector>

struct emu_dim3 { unsigned x, y, z; emu_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef emu_dim3 dim3;
extern emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define warpSize 32
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
typedef void *cudaGraph_t;
typedef void *cudaGraphExec_t;
#define cudaSuccess 0
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };

extern unsigned char *emu_dyn_smem;

namespace emu {
struct Fiber {
    ucontext_t ctx;
    char *stack;
    int state;      /* 0 runnable, 1 waiting, 2 done */
    unsigned *gen;  /* generation counter waited on */
    unsigned mygen;
    void *bt[10]; int nbt;   /* EMU_TRACE=1: call stack of the barrier waited on, printed on deadlock */
};
struct Warp {
    unsigned gen; int arrived; int live;
    uint64_t slot[2][32]; int pred[2][32]; unsigned par;
};
extern Fiber *fibers; extern Warp *warps;
extern int nthreads, cur;
extern unsigned blk_gen; extern int blk_arrived, blk_live;
extern ucontext_t sched_ctx;
void yield_wait(unsigned *gen, unsigned mygen);
void launch(emu_dim3 grid, emu_dim3 block, size_t smem, std::function<void()> body);
inline int lane() { return cur & 31; }
inline Warp &mywarp() { return warps[cur >> 5]; }
inline void warp_barrier() {
    Warp &w = mywarp();
    unsigned g = w.gen;
    if (++w.arrived >= w.live) { w.arrived = 0; w.gen++; }
    else yield_wait(&w.gen, g);
}
}

/* EMU_STRICT=1: every thread of the block must reach a block barrier from the same call site (two frames of the
 * call stack are compared); the first mismatch is reported with both stacks (addr2line -e the library). */
void emu_strict_check();
inline void __syncthreads() {
    emu_strict_check();
    unsigned g = emu::blk_gen;
    if (++emu::blk_arrived >= emu::blk_live) { emu::blk_arrived = 0; emu::blk_gen++; }
    else emu::yield_wait(&emu::blk_gen, g);
}
inline void __syncwarp(unsigned mask = 0xffffffffu) { (void)mask; emu::warp_barrier(); }
inline void __threadfence() {}
inline void __threadfence_block() {}

template <typename T> inline T emu_xchg(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of <= 8 bytes");
    emu::Warp &w = emu::mywarp();
    unsigned p = w.par & 1; /* all lanes read the same parity before the barrier */
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    w.slot[p][emu::lane()] = raw;
    int l = emu::lane();
    emu::warp_barrier();
    if (l == 0) w.par++; /* after the barrier: next call uses the other buffer */
    uint64_t r = (src >= 0 && src < 32) ? w.slot[p][src] : raw;
    T out; memcpy(&out, &r, sizeof(T));
    /* a second barrier keeps lane 0's par++ ordered w.r.t. slow lanes */
    emu::warp_barrier();
    return out;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src, int width = 32) { (void)width; return emu_xchg(v, src & 31); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) { (void)width; return emu_xchg(v, emu::lane() ^ m); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) { (void)width; int s = emu::lane() - (int)d; return emu_xchg(v, s < 0 ? emu::lane() : s); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) { (void)width; int s = emu::lane() + (int)d; return emu_xchg(v, s > 31 ? emu::lane() : s); }
inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned mine = pred ? 1u : 0u, out = 0;
    for (int l = 0; l < 32; l++) { /* 32 exchanges folded into one: use the slot buffer directly */ (void)l; break; }
    emu::Warp &w = emu::mywarp();
    unsigned p = w.par & 1;
    w.slot[p][emu::lane()] = mine;
    int l0 = emu::lane();
    emu::warp_barrier();
    if (l0 == 0) w.par++;
    int base = 0; (void)base;
    int nl = 32;
    /* lanes beyond the block's thread count do not exist */
    int wbase = (emu::cur >> 5) << 5;
    if (wbase + nl > emu::nthreads) nl = emu::nthreads - wbase;
    for (int l = 0; l < nl; l++) if (w.slot[p][l]) out |= 1u << l;
    emu::warp_barrier();
    return out;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) {
    int wbase = (emu::cur >> 5) << 5; int nl = emu::nthreads - wbase; if (nl > 32) nl = 32;
    unsigned full = nl == 32 ? 0xffffffffu : ((1u << nl) - 1);
    return __ballot_sync(m, pred) == full;
}
inline unsigned __activemask() { return 0xffffffffu; }
/* lanes holding the same value (full warps only) */
template <typename T> inline unsigned __match_any_sync(unsigned, T v) {
    static_assert(sizeof(T) <= 8, "match of <= 8 bytes");
    emu::Warp &w = emu::mywarp();
    unsigned p = w.par & 1;
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    w.slot[p][emu::lane()] = raw;
    int l = emu::lane();
    emu::warp_barrier();
    if (l == 0) w.par++;
    unsigned m = 0;
    for (int q = 0; q < 32; q++) if (w.slot[p][q] == raw) m |= 1u << q;
    emu::warp_barrier();
    return m;
}

inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
template <typename T> inline T __ldg(const T *p) { return *p; }

/* atomics: fibers never preempt each other, plain read-modify-write is atomic */
template <typename T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> inline T atomicSub(T *p, T v) { T o = *p; *p = o - v; return o; }
template <typename T> inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <typename T> inline T atomicXor(T *p, T v) { T o = *p; *p = o ^ v; return o; }
template <typename T> inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

/* host runtime shim */
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void *p) { free(p); return 0; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, int) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, int, cudaStream_t) {
    for (size_t r = 0; r < h; r++) memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
    return 0;
}
inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = 0; return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = 0; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
enum { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
inline cudaError_t cudaDeviceGetAttribute(int *v, int attr, int) { *v = attr == cudaDevAttrMultiProcessorCount ? 148 : 232448; return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
#define cudaStreamNonBlocking 1
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8

#define BLU_DYN_SMEM(name) unsigned char *name = emu_dyn_smem
#define BLU_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch(emu_dim3(grid), emu_dim3(block), (smem), [=]() { kernel(__VA_ARGS__); })

#endif
