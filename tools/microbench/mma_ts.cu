// Micro-benchmark: tcgen05.mma with the A operand in TENSOR MEMORY (kind::f16, fp16, M = 128).
//   1. layout check: A[128][64] written with tcgen05.st.32x32b (lane = row, column c = {A[r][2c] low half, A[r][2c+1] high half}),
//      B[128][64] K-major SWIZZLE_128B in shared memory, D = A * B^T read back and compared with the host.
//   2. issue rate of SS (A from shared memory) and TS (A from TMEM) products for N = 128 / 256, alone and while 16 other warps
//      stream st.shared.v4 (the converter warps of the learn kernels) -- shows who owns the 128 B/clk of the shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I avddpg_b200/csrc -o tools/microbench/mma_ts tools/microbench/mma_ts.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "avd_umma.cuh"
using namespace avd::umma;

__device__ __forceinline__ void mma_ts_p(uint32_t leader, uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "setp.ne.b32 q, %5, 0;\n"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- 1. layout check ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) ts_check_kernel(const __half* A, const __half* B, float* D) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int t = threadIdx.x, w = t >> 5;
    for (int i = t; i < 128 * 64; i += 128) {      // B[n][k] -> K-major, 128-byte rows, 16-byte chunks XOR-swizzled with n % 8
        const int n = i >> 6, k = i & 63;
        const uint32_t off = (n >> 3) * 1024 + (n & 7) * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
        *reinterpret_cast<__half*>(smem + off) = B[i];
    }
    if (t == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (t < 32) tmem_alloc(&tslot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    uint32_t r[32];
    for (int c = 0; c < 32; ++c) {
        const __half lo = A[t * 64 + 2 * c], hi = A[t * 64 + 2 * c + 1];
        r[c] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
    }
    tmem_st32(tmem + ((uint32_t)(32 * w) << 16) + 256, r);
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (t < 32) {
        const uint32_t leader = elect_one();
        constexpr uint32_t idesc = make_idesc_f16kind(128, 128, false, false, FMT_F16, FMT_F16);
        const uint64_t dB = make_smem_desc(smem_u32(smem), 16, 1024);
        for (int ks = 0; ks < 4; ++ks) mma_ts_p(leader, tmem, tmem + 256 + ks * 8, desc_add(dB, ks * 32), idesc, ks != 0);
        mma_commit_p(leader, &bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    float v[32];
    for (int c = 0; c < 4; ++c) {
        tmem_ld32(tmem + ((uint32_t)(32 * w) << 16) + 32 * c, v);
        for (int j = 0; j < 32; ++j) D[t * 128 + 32 * c + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (t < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---- 2. rates -------------------------------------------------------------------------------------------
// warp 0 issues; warps 1..16 store to shared memory when `nst` > 0.  out[0] = issuer cycles, out[1] = slowest storing warp's cycles.
template <int N, bool TS>
__global__ void __launch_bounds__(544, 1) ts_rate_kernel(long long* out, int reps, int nst) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    __shared__ long long tmax;
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    constexpr int kBBytes = 2 * N * 128;           // two 64-element K blocks
    for (int i = t; i < (32768 + kBBytes) / 4; i += 544) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (t == 0) { mbar_init(&bar, 1); fence_barrier_init(); tmax = 0; }
    if (t < 32) tmem_alloc(&tslot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    if (w >= 1 && w <= 4) {                          // A in TMEM: 128 lanes x 64 columns at column 256 + 128
        uint32_t r[32];
        for (int c = 0; c < 32; ++c) r[c] = 0x3c003c00u;
        tmem_st32(tmem + ((uint32_t)(32 * (w & 3)) << 16) + 384, r);
        tmem_st32(tmem + ((uint32_t)(32 * (w & 3)) << 16) + 416, r);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (w == 0) {
        const uint32_t leader = elect_one();
        constexpr uint32_t idesc = make_idesc_f16kind(128, N, false, false, FMT_F16, FMT_F16);
        const uint64_t dA = make_smem_desc(smem_u32(smem), 16, 1024);
        const uint64_t dB = make_smem_desc(smem_u32(smem + 32768), 16, 1024);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const uint32_t aoff = (ks >> 2) * (128 * 128) + (ks & 3) * 32;
                const uint32_t boff = (ks >> 2) * (N * 128) + (ks & 3) * 32;
                const uint32_t d = tmem + (N == 128 ? (r & 1) * 128 : 0);
                if (TS) mma_ts_p(leader, d, tmem + 384 + ks * 8, desc_add(dB, boff), idesc, ks != 0);
                else mma_bf16_p(leader, d, desc_add(dA, aoff), desc_add(dB, boff), idesc, ks != 0);
            }
        }
        mma_commit_p(leader, &bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (t == 0) out[0] = t1 - t0;
    } else if (nst > 0) {
        // conflict-free 16-byte stores: a warp writes 512 contiguous bytes per instruction (4 wavefronts)
        uint8_t* dst = smem + 32768 + kBBytes + (w - 1) * 2048 + lane * 16;
        const uint4 val = make_uint4(t, t, t, t);
        long long t0 = clock64();
        for (int i = 0; i < nst; ++i) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(dst + (i & 3) * 512)), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w) : "memory");
        }
        __syncwarp();
        long long t1 = clock64();
        if (lane == 0) atomicMax((unsigned long long*)&tmax, (unsigned long long)(t1 - t0));
    }
    tc_fence_before();
    __syncthreads();
    if (t == 0) out[1] = tmax;
    if (t < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, bool TS>
void rate(int nst) {
    long long* d;
    cudaMalloc(&d, 16);
    auto k = ts_rate_kernel<N, TS>;
    const int smem = 1024 + 32768 + 2 * N * 128 + 16 * 2048;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 256;
    for (int i = 0; i < 2; ++i) k<<<1, 544, smem>>>(d, reps, nst);
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s N=%3d stores/warp=%5d: %7.1f cycles per MMA (M=128,K=16)", TS ? "A in TMEM" : "A in smem", N, nst, (double)h[0] / (reps * 8));
    if (nst) printf("   st.shared: %6.1f B/clk over %lld cycles", 16.0 * nst * 512 / (double)h[1], h[1]);
    printf("  [%s]\n", cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    // ---- layout check
    std::vector<__half> A(128 * 64), B(128 * 64);
    std::vector<float> Af(128 * 64), Bf(128 * 64), D(128 * 128);
    srand(1);
    for (int i = 0; i < 128 * 64; ++i) {
        A[i] = __float2half((rand() % 2001 - 1000) / 1000.0f);
        B[i] = __float2half((rand() % 2001 - 1000) / 1000.0f);
        Af[i] = __half2float(A[i]);
        Bf[i] = __half2float(B[i]);
    }
    __half *dA, *dB;
    float* dD;
    cudaMalloc(&dA, A.size() * 2);
    cudaMalloc(&dB, B.size() * 2);
    cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(ts_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024);
    ts_check_kernel<<<1, 128, 20 * 1024>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0, worst_swapped = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
            double ref = 0, refs = 0;
            for (int k = 0; k < 64; ++k) {
                ref += (double)Af[m * 64 + k] * Bf[n * 64 + k];
                refs += (double)Af[m * 64 + (k ^ 1)] * Bf[n * 64 + k];
            }
            worst = fmax(worst, fabs(ref - D[m * 128 + n]));
            worst_swapped = fmax(worst_swapped, fabs(refs - D[m * 128 + n]));
        }
    printf("layout check [%s]: max |D - A*B^T| = %.3e (low half = even k), %.3e (low half = odd k)\n", cudaGetErrorString(e), worst, worst_swapped);

    // ---- rates
    rate<128, false>(0);
    rate<128, true>(0);
    rate<256, false>(0);
    rate<256, true>(0);
    rate<64, false>(0);
    rate<64, true>(0);
    for (int nst : {1200, 2400, 4800}) {
        rate<128, false>(nst);
        rate<128, true>(nst);
        rate<256, false>(nst);
        rate<256, true>(nst);
    }
    return 0;
}
