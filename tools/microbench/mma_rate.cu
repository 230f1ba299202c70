// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, bf16, M = 128) from shared-memory operands in K-major and MN-major
// layouts.  One CTA; one elected thread issues `reps` x 8 MMAs (a K = 128 product), commits, waits.  Prints cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I avddpg_b200/csrc -o /tmp/mma_rate tools/microbench/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "avd_umma.cuh"
using namespace avd::umma;

template <int N, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int reps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tslot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    if (threadIdx.x < 32) {
        const uint32_t leader = elect_one();
        constexpr uint32_t idesc = make_idesc_bf16(128, N, A_MN, B_MN);
        // A at 0 (32 KB: two 64-element halves of 128 rows x 128 B), B at 32 KB
        const uint64_t dA = A_MN ? make_smem_desc(smem_u32(smem), 128 * 128, 1024) : make_smem_desc(smem_u32(smem), 16, 1024);
        const uint64_t dB = B_MN ? make_smem_desc(smem_u32(smem + 32768), 128 * 128, 1024) : make_smem_desc(smem_u32(smem + 32768), 16, 1024);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const uint32_t aoff = A_MN ? ks * 2048 : (ks >> 2) * (128 * 128) + (ks & 3) * 32;
                const uint32_t boff = B_MN ? ks * 2048 : (ks >> 2) * (128 * 128) + (ks & 3) * 32;
                mma_bf16_p(leader, tmem + (r & 1) * 128, desc_add(dA, aoff), desc_add(dB, boff), idesc, ks != 0);
            }
        }
        mma_commit_p(leader, &bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, bool A_MN, bool B_MN>
void run(const char* name) {
    long long* d;
    cudaMalloc(&d, 8);
    auto k = rate_kernel<N, A_MN, B_MN>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    const int reps = 64;
    for (int i = 0; i < 2; ++i) k<<<1, 128, 80 * 1024>>>(d, reps);
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-28s N=%3d: %7.1f cycles per MMA (M=128, K=16)  [%s]\n", name, N, (double)h / (reps * 8), cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<128, false, false>("A K-major,  B K-major");
    run<128, true, false>("A MN-major, B K-major");
    run<128, false, true>("A K-major,  B MN-major");
    run<128, true, true>("A MN-major, B MN-major");
    run<64, false, false>("A K-major,  B K-major");
    run<64, false, true>("A K-major,  B MN-major");
    run<64, true, true>("A MN-major, B MN-major");
    run<16, false, false>("A K-major,  B K-major");
    run<16, true, false>("A MN-major, B K-major");
    return 0;
}
