/* cuda_emu.cpp -- scheduler of the SIMT emulator (test/debug infrastructure, see cuda_emu.h) */
#include "cuda_emu.h"
#include <execinfo.h>
#include <dlfcn.h>

emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
unsigned char *emu_dyn_smem = nullptr;

namespace emu {
Fiber *fibers = nullptr; Warp *warps = nullptr;
int nthreads = 0, cur = 0;
unsigned blk_gen = 0; int blk_arrived = 0, blk_live = 0;
ucontext_t sched_ctx;
static std::function<void()> *g_body;
static const size_t STACK = 256 * 1024;

}
void emu_strict_check() {
    static const int strict = getenv("EMU_STRICT") != nullptr;
    if (!strict) return;
    static void *site[6]; static unsigned site_gen = ~0u; static int site_thread = -1;
    void *bt0[8];
    const int n = backtrace(bt0, 8) - 1;
    void **bt = bt0 + 1;      /* (frame 0 is this function) */
    if (site_gen != emu::blk_gen || emu::blk_arrived == 0) { site_gen = emu::blk_gen; site_thread = emu::cur; for (int q = 0; q < 6; q++) site[q] = q < n ? bt[q] : nullptr; return; }
    for (int q = 3; q < 5; q++)      /* (frames 0-1 are this check and __syncthreads; frame 2 differs when the compiler duplicates a barrier) */
        if ((q < n ? bt[q] : nullptr) != site[q]) {
            Dl_info di;
            fprintf(stderr, "cuda_emu: DIVERGENT block barrier: thread %d at", emu::cur);
            for (int r = 0; r < n; r++) if (dladdr(bt[r], &di) && di.dli_fbase) fprintf(stderr, " +0x%zx", (size_t)((char *)bt[r] - (char *)di.dli_fbase));
            fprintf(stderr, "\n   thread %d at", site_thread);
            for (int r = 0; r < 6; r++) if (site[r] && dladdr(site[r], &di) && di.dli_fbase) fprintf(stderr, " +0x%zx", (size_t)((char *)site[r] - (char *)di.dli_fbase));
            fprintf(stderr, "\n");
            abort();
        }
}
namespace emu {
void yield_wait(unsigned *gen, unsigned mygen) {
    Fiber &f = fibers[cur];
    f.state = 1; f.gen = gen; f.mygen = mygen;
    static const int trace = getenv("EMU_TRACE") != nullptr;
    f.nbt = trace ? backtrace(f.bt, 10) : 0;
    int me = cur;
    swapcontext(&f.ctx, &sched_ctx);
    cur = me;
    threadIdx.x = (unsigned)me;
}

static void fiber_main() {
    (*g_body)();
    Fiber &f = fibers[cur];
    f.state = 2;
    /* an exited thread no longer takes part in barriers */
    Warp &w = warps[cur >> 5];
    w.live--; blk_live--;
    if (w.live > 0 && w.arrived >= w.live) { w.arrived = 0; w.gen++; }
    if (blk_live > 0 && blk_arrived >= blk_live) { blk_arrived = 0; blk_gen++; }
    swapcontext(&f.ctx, &sched_ctx);
}

void launch(emu_dim3 grid, emu_dim3 block, size_t smem, std::function<void()> body) {
    g_body = &body;
    nthreads = (int)block.x;
    blockDim = block; gridDim = grid;
    int nw = (nthreads + 31) / 32;
    fibers = (Fiber *)calloc((size_t)nthreads, sizeof(Fiber));
    warps = (Warp *)calloc((size_t)nw, sizeof(Warp));
    for (int t = 0; t < nthreads; t++) fibers[t].stack = (char *)malloc(STACK);
    unsigned char *dyn = (unsigned char *)aligned_alloc(128, ((smem + 127) / 128 + 1) * 128);
    for (unsigned b = 0; b < grid.x; b++) {
        blockIdx = emu_dim3(b);
        memset(dyn, 0xCD, smem); /* shared memory is not zero-initialised on the GPU either */
        emu_dyn_smem = dyn;
        blk_gen = 0; blk_arrived = 0; blk_live = nthreads;
        for (int w = 0; w < nw; w++) {
            memset(&warps[w], 0, sizeof(Warp));
            int l = nthreads - 32 * w; warps[w].live = l > 32 ? 32 : l;
        }
        for (int t = 0; t < nthreads; t++) {
            Fiber &f = fibers[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK; f.ctx.uc_link = &sched_ctx;
            makecontext(&f.ctx, (void (*)())fiber_main, 0);
            f.state = 0;
        }
        int done = 0;
        /* EMU_SEED != 0: visit the runnable threads in a different pseudo-random order on every scheduler
         * pass.  Between two barriers the real hardware promises no order among lanes either, so code
         * that only works in lane order (a missing __syncwarp between a read and a write of the same
         * word, say) shows up as a parity failure under some seed. */
        unsigned long long rng = 0;
        { const char *e = getenv("EMU_SEED"); rng = e ? strtoull(e, nullptr, 10) * 0x9E3779B97F4A7C15ULL + b : 0; }
        std::vector<int> order((size_t)nthreads);
        for (int t = 0; t < nthreads; t++) order[(size_t)t] = t;
        while (done < nthreads) {
            int progressed = 0;
            if (rng) {
                for (int t = nthreads - 1; t > 0; t--) {
                    rng = rng * 6364136223846793005ULL + 1442695040888963407ULL;
                    int u = (int)((rng >> 33) % (unsigned long long)(t + 1));
                    int tmp = order[(size_t)t]; order[(size_t)t] = order[(size_t)u]; order[(size_t)u] = tmp;
                }
            }
            for (int ti = 0; ti < nthreads; ti++) {
                const int t = order[(size_t)ti];
                Fiber &f = fibers[t];
                if (f.state == 2) continue;
                if (f.state == 1) { if (*f.gen == f.mygen) continue; f.state = 0; }
                cur = t; threadIdx = emu_dim3((unsigned)t);
                swapcontext(&sched_ctx, &f.ctx);
                progressed = 1;
                if (f.state == 2) done++;
            }
            if (!progressed) {
                fprintf(stderr, "cuda_emu: DEADLOCK in block %u (divergent barrier?)\n", b);
                for (int t = 0; t < nthreads; t++)
                    if (fibers[t].state == 1) {
                        fprintf(stderr, "  thread %d waits on %s barrier", t,
                                fibers[t].gen == &blk_gen ? "block" : "warp");
                        for (int q = 0; q < fibers[t].nbt; q++) {      /* offsets into this library, for addr2line -e */
                            Dl_info di;
                            if (dladdr(fibers[t].bt[q], &di) && di.dli_fbase) fprintf(stderr, " +0x%zx", (size_t)((char *)fibers[t].bt[q] - (char *)di.dli_fbase));
                        }
                        fprintf(stderr, "\n");
                    }
                abort();
            }
        }
    }
    for (int t = 0; t < nthreads; t++) free(fibers[t].stack);
    free(fibers); free(warps); free(dyn);
    fibers = nullptr; warps = nullptr; emu_dyn_smem = nullptr;
}
}
