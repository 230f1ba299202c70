// avd_peer.cu -- the one exchange step of the federated (interfrl) round as a single NVLink-native kernel:
//
//   every rank has written its per-system partial sums [systems][pitch] (last used column = the local member count or
//   weight sum) into its half of a SYMMETRIC buffer (same offset in every rank's peer-mapped allocation); this kernel
//     1. signals "my partial sums are complete" into every peer's flag word and waits for all peers' signals
//        (st.release.sys / ld.acquire.sys on peer-mapped memory, monotonically increasing epoch, no reset),
//     2. reads the SUM over ranks -- with `multimem.ld_reduce` on the NVLS multicast mapping, i.e. reduced inside the
//        NVSwitch, or, without a multicast mapping, with plain peer loads over NVLink --
//     3. divides by the reduced count / weight sum (federated.py:62, :110) and writes the means to local memory.
// It replaces ncclAllReduce + the finalize kernel of the NCCL transport (avddpg_b200/server/federated.py); the payload is
// 1.2 - 2.5 MB, so the round is latency-bound and one launch with in-switch reduction is what matters.
// Buffers alternate between two halves from round to round: a rank can only pass the barrier of round k+1 after every
// peer has finished reading round k, so half (k & 1) is free again when round k+2 writes it.
#include "avd_common.cuh"

namespace avd {

struct PeerArgs {
    int rank, world;
    uint32_t epoch;
    uint64_t peer_base[AVD_MAX_PEERS];     // peer-mapped base address of every rank's symmetric allocation
    uint64_t multicast_base;               // NVLS multicast mapping of the same allocation (0: none)
    int64_t flag_off, data_off;            // byte offsets inside the allocation: epoch flags [AVD_MAX_PEERS] u32, partial sums
    float* out;                            // [systems][pitch] local result
    int64_t pitch;                         // floats per system row (multiple of 4)
    int n_systems;
    int64_t n;                             // payload columns; column n carries the divisor
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 multimem_sum4(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) fed_exchange_peer_kernel(PeerArgs g) {
    // ---- 1. cross-rank barrier: block 0 signals, every block waits on the LOCAL flag words
    if (blockIdx.x == 0 && threadIdx.x < g.world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t*>(g.peer_base[threadIdx.x] + g.flag_off) + g.rank, g.epoch);
    }
    if (threadIdx.x < g.world) {
        const uint32_t* flag = reinterpret_cast<const uint32_t*>(g.peer_base[g.rank] + g.flag_off) + threadIdx.x;
        while ((int32_t)(ld_acquire_sys(flag) - g.epoch) < 0) {
        }
    }
    __syncthreads();
    // ---- 2./3. reduced sums -> means
    const int64_t row_f4 = g.pitch / 4;
    const int64_t total = (int64_t)g.n_systems * row_f4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(i / row_f4);
        const int64_t c4 = (i - (int64_t)s * row_f4) * 4;
        const int64_t off = g.data_off + ((int64_t)s * g.pitch + c4) * 4;
        const int64_t doff = g.data_off + ((int64_t)s * g.pitch + (g.n & ~(int64_t)3)) * 4;     // float4 that holds the divisor column
        float4 v, dv;
        if (g.multicast_base) {
            v = multimem_sum4(reinterpret_cast<const float*>(g.multicast_base + off));
            dv = multimem_sum4(reinterpret_cast<const float*>(g.multicast_base + doff));
        } else {
            v = make_float4(0.f, 0.f, 0.f, 0.f);
            dv = v;
            for (int r = 0; r < g.world; ++r) {
                const float4 a = *reinterpret_cast<const float4*>(g.peer_base[r] + off), b = *reinterpret_cast<const float4*>(g.peer_base[r] + doff);
                v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
                dv.x += b.x; dv.y += b.y; dv.z += b.z; dv.w += b.w;
            }
        }
        const int k = (int)(g.n & 3);
        const float inv = 1.0f / (k == 0 ? dv.x : k == 1 ? dv.y : k == 2 ? dv.z : dv.w);
        float* dst = g.out + (int64_t)s * g.pitch + c4;
        const float o[4] = {v.x * inv, v.y * inv, v.z * inv, v.w * inv};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (c4 + j < g.n) dst[j] = o[j];
    }
}

}  // namespace avd

extern "C" int avd_fed_exchange_peer(const avd_peer_comm* comm, int64_t flag_offset, int64_t data_offset, float* out, int64_t pitch,
                                     int32_t n_systems, int64_t n, void* stream) {
    using namespace avd;
    AVD_REQUIRE(comm && out, "null argument");
    AVD_REQUIRE(comm->world >= 1 && comm->world <= AVD_MAX_PEERS && comm->rank >= 0 && comm->rank < comm->world, "bad rank / world (max %d peers)", AVD_MAX_PEERS);
    AVD_REQUIRE(pitch % 4 == 0 && n >= 1 && n < pitch && n_systems >= 1, "pitch must be a multiple of 4 floats and hold n + 1 columns");
    AVD_REQUIRE(flag_offset % 4 == 0 && data_offset % 16 == 0, "misaligned offsets");
    PeerArgs g;
    g.rank = comm->rank; g.world = comm->world; g.epoch = comm->epoch;
    for (int r = 0; r < AVD_MAX_PEERS; ++r) g.peer_base[r] = r < comm->world ? comm->peer_base[r] : 0;
    g.multicast_base = comm->multicast_base;
    g.flag_off = flag_offset; g.data_off = data_offset; g.out = out; g.pitch = pitch; g.n_systems = n_systems; g.n = n;
    for (int r = 0; r < comm->world; ++r) AVD_REQUIRE(g.peer_base[r] != 0, "peer %d has no mapped buffer", r);
    const int64_t work = (int64_t)n_systems * (pitch / 4);
    // few CTAs: every one of them spins on the barrier flags, and 1-2.5 MB of payload needs no more than a few hundred loads in flight per SM
    const int grid = (int)std::min<int64_t>((work + 255) / 256, sm_count());
    fed_exchange_peer_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g);
    AVD_LAUNCH_OK();
    return AVD_OK;
}
