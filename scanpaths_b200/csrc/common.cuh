// Shared helpers for the scanpaths_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/scanpaths_b200.h"

namespace spb {

void set_error(const char *fmt, ...);
void count_launch();
// optional live profiling: CUDA-event pairs around tagged launches (see spb_profile_enable)
enum ProfTag { kTagConvX = 1, kTagConvH = 2, kTagConvP = 3, kTagCell = 4, kTagHead = 5, kTagFeedback = 6, kTagRank1 = 7,
               kTagPrep = 8, kTagWinoIn = 9, kTagScore = 10, kTagSample = 11 };
void prof_begin(int tag, cudaStream_t s);
void prof_end(cudaStream_t s);

#define SPB_CHECK_ARG(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            spb::set_error("%s: %s", __func__, msg);               \
            return SPB_ERR_ARG;                                    \
        }                                                          \
    } while (0)

#define SPB_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            spb::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__));   \
            return SPB_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

#define SPB_LAUNCH_CHECK()                                                                   \
    do {                                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                \
        spb::count_launch();                                                                 \
        if (e__ != cudaSuccess) {                                                            \
            spb::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e__)); \
            return SPB_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

int num_sms();   // SM count of the current device (148 on B200), queried once per device

__device__ __forceinline__ double shfl_up_f64(double v, int delta) {
    return __shfl_up_sync(0xffffffffu, v, delta);
}
__device__ __forceinline__ double shfl_idx_f64(double v, int lane) {
    return __shfl_sync(0xffffffffu, v, lane);
}

}  // namespace spb
