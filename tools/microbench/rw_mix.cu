// rw_mix.cu -- what HBM bandwidth does a perfectly coalesced streaming kernel reach for a given read : write byte mix?
// The env-step training launch reads 32 B and writes 76 B per vehicle-step (a 30 : 70 mix); MEASURED_PEAKS.json's figure is a
// 50 : 50 copy.  Each thread moves float4s: NR input streams, NW output streams.   nvcc -arch=sm_100a -O3 rw_mix.cu -o rw_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int NR, int NW>
__global__ void __launch_bounds__(256) mix(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const float4 v = in[r * n + i];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) out[w * n + i] = make_float4(acc.x + w, acc.y, acc.z, acc.w);
    }
}

template <int NR, int NW>
void run(const float4* in, float4* out, size_t n) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = 148 * 8;
    for (int i = 0; i < 3; ++i) mix<NR, NW><<<grid, 256>>>(in, out, n);
    float best = 1e30f;
    for (int i = 0; i < 10; ++i) {
        cudaEventRecord(a);
        mix<NR, NW><<<grid, 256>>>(in, out, n);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double bytes = (double)(NR + NW) * n * 16;
    printf("read streams %d  write streams %d  (%2.0f %% writes): %7.1f GB/s  (%.3f ms, %.2f GB moved)\n", NR, NW, 100.0 * NW / (NR + NW),
           bytes / best / 1e6, best, bytes / 1e9);
}

int main() {
    const size_t n = (size_t)1 << 24;          // 16 Mi float4 = 256 MiB per stream
    float4 *in, *out;
    cudaMalloc(&in, 8 * n * 16); cudaMalloc(&out, 8 * n * 16);
    cudaMemset(in, 0, 8 * n * 16);
    run<1, 0>(in, out, n); run<4, 0>(in, out, n); run<1, 1>(in, out, n); run<2, 2>(in, out, n); run<0, 1>(in, out, n); run<0, 4>(in, out, n);
    run<3, 7>(in, out, n); run<2, 5>(in, out, n); run<1, 2>(in, out, n); run<2, 1>(in, out, n);
    return 0;
}
