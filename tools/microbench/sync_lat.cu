// Micro-benchmark: latencies of the hand-offs the persistent learn kernels are built from (one CTA, sm_100a):
//   mbarrier arrive -> waiter released (try_wait with a suspend-time hint vs a spin loop), tcgen05.mma + commit -> waiter released,
//   tcgen05.ld (32 and 64 columns) + wait::ld, 8 x st.shared.v4 + fence.proxy.async.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I avddpg_b200/csrc -o tools/microbench/sync_lat tools/microbench/sync_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "avd_umma.cuh"
using namespace avd::umma;

__device__ __forceinline__ bool try_wait_nohint(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
template <int MODE>
__device__ __forceinline__ void wait_mode(uint64_t* bar, uint32_t parity) {
    if (MODE == 0) mbar_wait(bar, parity);                          // try_wait + suspend hint (the library's form)
    else if (MODE == 1) { while (!try_wait_nohint(bar, parity)) {} }   // try_wait, default time limit
    else { while (!test_wait(bar, parity)) {} }                     // pure spin
}

// out[0]: arrive -> release, out[1]: mma + commit -> release, out[2]: ld32, out[3]: ld64, out[4]: 8 sts + fence.proxy.async
template <int MODE>
__global__ void __launch_bounds__(128, 1) lat_kernel(long long* out, int reps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tslot;
    __shared__ volatile long long stamp;
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    for (int i = t; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (t == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    if (w == 0) tmem_alloc(&tslot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    long long acc0 = 0, acc1 = 0;
    // ---- arrive -> release
    for (int r = 0; r < reps; ++r) {
        if (w == 0) {
            long long t0 = clock64();
            while (clock64() - t0 < 3000) {}
            if (lane == 0) { stamp = clock64(); mbar_arrive(&bar[0]); }
            __syncwarp();
        } else if (w == 1) {
            wait_mode<MODE>(&bar[0], r & 1);
            long long t1 = clock64();
            acc0 += t1 - stamp;
        }
        __syncthreads();
    }
    // ---- mma + commit -> release
    for (int r = 0; r < reps; ++r) {
        if (w == 0) {
            const uint32_t leader = elect_one();
            long long t0 = clock64();
            while (clock64() - t0 < 3000) {}
            constexpr uint32_t idesc = make_idesc_f16kind(128, 64, false, false, FMT_F16, FMT_F16);
            const uint64_t dA = make_smem_desc(smem_u32(smem), 16, 1024), dB = make_smem_desc(smem_u32(smem + 32768), 16, 1024);
            if (lane == 0) stamp = clock64();
            __syncwarp();
            mma_bf16_p(leader, tmem, dA, dB, idesc, 0);
            mma_commit_p(leader, &bar[1]);
        } else if (w == 1) {
            wait_mode<MODE>(&bar[1], r & 1);
            long long t1 = clock64();
            acc1 += t1 - stamp;
        }
        __syncthreads();
    }
    if (t == 32) { out[0] = acc0 / reps; out[1] = acc1 / reps; }
    // ---- tcgen05.ld
    if (w == 2) {
        float v[64];
        float s = 0;
        long long a2 = 0, a3 = 0;
        for (int r = 0; r < reps; ++r) {
            long long t0 = clock64();
            tmem_ld32(tmem + ((uint32_t)64 << 16), v);
            long long t1 = clock64();
            s += v[r & 31];
            a2 += t1 - t0;
            t0 = clock64();
            tmem_ld64(tmem + ((uint32_t)64 << 16), v);
            t1 = clock64();
            s += v[r & 63];
            a3 += t1 - t0;
        }
        if (lane == 0) { out[2] = a2 / reps; out[3] = a3 / reps; out[7] = (long long)s; }
    }
    // ---- 8 x st.shared.v4 + fence.proxy.async
    if (w == 3) {
        long long a4 = 0, a5 = 0;
        uint8_t* row = smem + 16384 + lane * 128;
        for (int r = 0; r < reps; ++r) {
            const uint4 pk = make_uint4(r, r + 1, r + 2, r + 3);
            long long t0 = clock64();
#pragma unroll
            for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(row + ((k ^ (lane & 7)) << 4)) = pk;
            long long t1 = clock64();
            fence_proxy_async();
            long long t2 = clock64();
            a4 += t1 - t0;
            a5 += t2 - t1;
        }
        if (lane == 0) { out[4] = a4 / reps; out[5] = a5 / reps; }
    }
    tc_fence_before();
    __syncthreads();
    if (w == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int MODE>
void run(const char* name) {
    long long* d;
    cudaMalloc(&d, 64);
    cudaMemset(d, 0, 64);
    cudaFuncSetAttribute(lat_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    lat_kernel<MODE><<<1, 128, 64 * 1024>>>(d, 200);
    long long h[8];
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-34s arrive->release %5lld | mma(N=64)+commit->release %5lld | ld32 %4lld | ld64 %4lld | 8 sts %4lld + fence.proxy.async %4lld cycles [%s]\n", name,
           h[0], h[1], h[2], h[3], h[4], h[5], cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<0>("try_wait + suspend hint (library)");
    run<1>("try_wait, default limit");
    run<2>("test_wait spin");
    return 0;
}
