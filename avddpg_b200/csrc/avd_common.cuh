// avd_common.cuh -- error plumbing and launch helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/avddpg_b200.h"

namespace avd {

void set_error(const char* fmt, ...);           // defined in avd_lib.cu
void count_launch(int n = 1);                    // kernels launched by this library so far (avd_kernel_launches)
int sm_count();                                  // cached multiprocessor count of the current device

#define AVD_REQUIRE(cond, ...)                    \
    do {                                          \
        if (!(cond)) {                            \
            ::avd::set_error(__VA_ARGS__);        \
            return AVD_ERR_INVALID_ARG;           \
        }                                         \
    } while (0)

#define AVD_CUDA_OK(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ::avd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return AVD_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define AVD_LAUNCH_OK()                                                                     \
    do {                                                                                    \
        ::avd::count_launch();                                                              \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess) {                                                            \
            ::avd::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return AVD_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

// Persistent-style grid: enough 256-thread CTAs to fill every SM to 2048 threads, never more than
// the work needs (B200: 148 SMs x 8 CTAs).
inline int grid_for(int64_t work_items, int threads = 256, int ctas_per_sm = 8) {
    int64_t need = (work_items + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// Grid for a grid-stride kernel: exactly the number of CTAs that are co-resident (SMs x occupancy), so
// there is a single wave and no tail, capped by the work available.
template <class Kernel>
inline int resident_grid(Kernel kernel, int64_t work_items, int threads = 256, size_t smem = 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = 1;
    }
    int64_t need = (work_items + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count() * occ;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace avd
