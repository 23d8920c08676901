/* blu_fact_inst.cu -- one instantiation of the factorization kernel per translation unit (compiled once per
 * CTA size with -DFACT_NT=...), so that the six of them build in parallel. */
#ifdef BLU_EMU
#include "cuda_emu.h"
#endif
#include "blu_types.h"
#include "blu_dev_common.cuh"
#include "blu_factor_build.cuh"
#include "blu_fact_launch.h"

#ifndef FACT_NT
#error "compile with -DFACT_NT=32|64|128|256|512|1024"
#endif
#define BLU_CAT2(a, b) a##b
#define BLU_CAT(a, b) BLU_CAT2(a, b)

int BLU_CAT(blu_launch_factorize_, FACT_NT)(cudaStream_t stream, const BluDev &dv, int nslot, int cap, int mode, int kd, int resident, int rerun, size_t smem) {
#ifndef BLU_EMU
    /* static + dynamic shared memory beyond 48 KB needs the opt-in (the static part is ~2.4 KB) */
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_factorize<FACT_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
#endif
    BLU_LAUNCH(k_factorize<FACT_NT>, nslot, FACT_NT, smem, stream, dv, cap, mode, kd, resident, rerun);
    return (int)cudaGetLastError();
}
