/* blu_fact_launch.h -- launchers of k_factorize<NT>, one per translation unit (blu_fact_inst.cu).
 * Return a cudaError_t value (0 = success). */
#ifndef BLU_FACT_LAUNCH_H
#define BLU_FACT_LAUNCH_H
#include "blu_types.h"
#define BLU_DECL_LAUNCH(NT) int blu_launch_factorize_##NT(cudaStream_t stream, const BluDev &dv, int nslot, int cap, int mode, int kd, int resident, int rerun, size_t smem)
BLU_DECL_LAUNCH(32); BLU_DECL_LAUNCH(64); BLU_DECL_LAUNCH(128); BLU_DECL_LAUNCH(256); BLU_DECL_LAUNCH(512); BLU_DECL_LAUNCH(1024);
#endif
