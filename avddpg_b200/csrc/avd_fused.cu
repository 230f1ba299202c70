// avd_fused.cu -- fused two-layer forward of the actor / critic on sm_100a:
//
//   x[n] (4 state words [+ action])  --tcgen05.mma (hi/lo split bf16, K = 16)-->  z1 in TMEM
//   --converter warps-->  r1 = relu(z1) as bf16, written straight into the 128-byte-swizzled K-major A tile in shared
//   memory (never to HBM unless the caller wants it for wgrad) plus one sign bit per element for the ReLU backward
//   --tcgen05.mma, W2'^T resident in shared memory (TMA), fp32 accumulators in TMEM-->  z2
//   --epilogue from TMEM (one thread per row)-->  BN(relu(z2 + b2')) . W3 + b3  -> tanh*high | r + gamma*q | q
//
// The layer-1 BatchNormalization (inference affine h1 = r1*sc1 + sh1) is FOLDED into layer 2 by the pack
// kernel:  W2' = diag(sc1) W2,  b2' = b2 + sh1 W2,  so the converters only clamp and the stored activation is r1.
// Reference semantics: agent/model.py:19-37 (actor), 55-83 (critic); workers/trainer.py:493-495, 502-503, 287.
#include <cudaTypedefs.h>
#include <stdlib.h>

#include <algorithm>

#include "avd_common.cuh"
#include "avd_ddpg_layout.cuh"
#include "avd_umma.cuh"

namespace avd {
namespace fused {

using namespace umma;
typedef __nv_bfloat16 bf16;

constexpr int TILE_M = 128, L2N = 128, KB = 64;
constexpr int MAX_KB = 5;                      // up to 320 input features for layer 2
constexpr int PTAB_COLS = MAX_KB * KB;         // 320
constexpr int SLOT_BYTES = TILE_M * KB * 2;    // 16 KB
constexpr int W_BYTES = MAX_KB * L2N * KB * 2; // 80 KB

enum HeadMode { HEAD_NONE = 0, HEAD_ACTOR = 1, HEAD_TARGET = 2, HEAD_Q = 3, HEAD_BWD_ACTION = 4 };

struct Args {
    avd_net_dims d;
    int critic;                 // 0 actor, 1 critic
    int A;
    int64_t R;
    const float* params;        // [A][pstride]
    int64_t pstride;
    const float* b2f;           // [A][128] folded layer-2 bias  b2 + sh1 W2  (pack_fold_kernel)
    const float* s;             // element (n, k) at s[n*s_rs + k*s_cs]
    int64_t s_rs, s_cs;
    const float* act;           // [A*R] (critic)
    bf16* H_out;                // nullable: [A*R][F] bf16 r1 = relu(z1) (operand of the wgrad GEMM)
    uint32_t* mask_out;         // nullable: [A*R][mask_words] sign bits of z1 (column j of word w at bit 31-j, 1 = negative)
    int mask_words;
    float* Z_out;               // nullable: [A*R][128] raw layer-2 product (input of the head-backward kernels)
    int head;                   // HeadMode
    const float* rew;
    float gamma, high;
    float* out;                 // [A*R]
    bf16* DZ_out;               // HEAD_BWD_ACTION: [A*R][128] bf16  d(-mean q)/dz2   (trainer.py:503-506)
    float* loss;                // HEAD_BWD_ACTION: loss[2*agent+1] += -mean(q)  (nullable)
    int tiles_per_agent, total_tiles, tiles_per_cta;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ---------------------------------------------------------------------------------------------------------
// Epilogue role (4 warps, TMEM lane quadrant = warp % 4): bias + ReLU + BN + W3 head on the fp32 accumulator of the
// layer-2 MMA, one output row per thread; optional coalesced z2 / dz2 stores.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_role(const Args& g, uint8_t* smem, float* etab, float* scratch_base, uint64_t* acc_full,
                                              uint64_t* acc_empty, uint32_t tmem_base, int warp, int lane, int tile_begin, int tile_end) {
    const avd_net_dims d = g.d;
        const int q = warp & 3;
        const int et = threadIdx.x - 32;     // 0..127
        int cur_agent = -1;
        for (int t = tile_begin, tc = 0; t < tile_end; ++t, ++tc) {
            const int agent = t / g.tiles_per_agent;
            const int tile_in_agent = t - agent * g.tiles_per_agent;
            if (agent != cur_agent) {        // per-agent head parameters -> shared (broadcast reads below)
                asm volatile("bar.sync 2, 128;" ::: "memory");
                const float* P = g.params + (int64_t)agent * g.pstride;
                int64_t og2, obe2, omu2, ovar2, oW3, ob3;
                if (g.critic) { const CriticOff o = critic_off(d); og2 = o.g2; obe2 = o.be2; omu2 = o.mu2; ovar2 = o.var2; oW3 = o.W3; ob3 = o.b3; }
                else { const ActorOff o = actor_off(d); og2 = o.g2; obe2 = o.be2; omu2 = o.mu2; ovar2 = o.var2; oW3 = o.W3; ob3 = o.b3; }
                {
                    const int c = et;
                    const float inv = 1.0f / sqrtf(P[ovar2 + c] + kBnEps);
                    const float sc = P[og2 + c] * inv;
                    etab[c] = g.b2f[(int64_t)agent * L2N + c];
                    etab[L2N + c] = sc;
                    etab[2 * L2N + c] = P[obe2 + c] - P[omu2 + c] * sc;
                    etab[3 * L2N + c] = P[oW3 + c];
                    if (c == 0) etab[4 * L2N] = P[ob3];
                }
                asm volatile("bar.sync 2, 128;" ::: "memory");
                cur_agent = agent;
            }
            const int buf = tc & 1;
            const uint32_t nt = (uint32_t)tc >> 1;
            mbar_wait(&acc_full[buf], nt & 1);
            tc_fence_after();
            const int64_t row_in_agent = (int64_t)tile_in_agent * TILE_M + q * 32 + lane;
            const bool valid = row_in_agent < g.R;
            const int64_t n = (int64_t)agent * g.R + row_in_agent;
            float acc = 0.0f;
#pragma unroll 1
            for (int c = 0; c < L2N / 32; ++c) {
                float v[32];
                tmem_ld32(tmem_base + (uint32_t)(buf * L2N) + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
                if (g.Z_out) {   // coalesced through the per-warp transpose scratch
                    const int64_t blk_row = (int64_t)tile_in_agent * TILE_M + q * 32;
                    const int rows_here = (int)min((int64_t)32, g.R - blk_row);
                    if (rows_here > 0)
                        store_block_32x32(scratch_base + (warp - 1) * (32 * 33), v,
                                          g.Z_out + ((int64_t)agent * g.R + blk_row) * L2N + c * 32, L2N, rows_here, 32, lane);
                }
                if (g.head == HEAD_BWD_ACTION) {
                    // actor loss -mean(q): dq = -1/R for every row, so dz2 = (z2 > 0) ? -W3*g2*inv2/R : 0 needs no row reduction
                    const float neg_inv_R = -1.0f / (float)g.R;
                    float dz[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = c * 32 + j;
                        const float z = v[j] + etab[col];
                        acc = fmaf(fmaf(fmaxf(z, 0.0f), etab[L2N + col], etab[2 * L2N + col]), etab[3 * L2N + col], acc);
                        dz[j] = z > 0.0f ? neg_inv_R * etab[3 * L2N + col] * etab[L2N + col] : 0.0f;
                    }
                    const int64_t blk_row = (int64_t)tile_in_agent * TILE_M + q * 32;
                    const int rows_here = (int)min((int64_t)32, g.R - blk_row);
                    if (rows_here > 0)
                        store_block_32x32_bf16(scratch_base + (warp - 1) * (32 * 33), dz,
                                               g.DZ_out + ((int64_t)agent * g.R + blk_row) * L2N + c * 32, L2N, rows_here, lane);
                } else if (g.head != HEAD_NONE) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = c * 32 + j;
                        const float z = v[j] + etab[col];
                        acc = fmaf(fmaf(fmaxf(z, 0.0f), etab[L2N + col], etab[2 * L2N + col]), etab[3 * L2N + col], acc);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            if (g.head == HEAD_BWD_ACTION) {
                if (g.loss) {
                    const float part = warp_sum(valid ? -(acc + etab[4 * L2N]) / (float)g.R : 0.0f);
                    if (lane == 0) atomicAdd(g.loss + 2 * agent + 1, part);
                }
            } else if (g.head != HEAD_NONE && valid) {
                const float pre = acc + etab[4 * L2N];
                float o;
                if (g.head == HEAD_ACTOR) o = g.high * tanhf(pre);
                else if (g.head == HEAD_TARGET) o = g.rew[n] + g.gamma * pre;     // trainer.py:494
                else o = pre;
                g.out[n] = o;
            }
        }
}

// =========================================================================================================
// Layer 1 ALSO runs on the tensor cores.
//
// Computing layer 1 on the CUDA cores costs ~9 instructions per element and is issue-bound.  Here the
// state part of layer 1 (ns <= 4 inputs + bias -> 256 features) is ONE tcgen05.mma per 128 output columns with K = 16:
//     A1 row = [ v_hi(5) | v_lo(5) | v_hi(5) | 0 ],  v = (s0, s1, s2, s3, 1)            (bf16 hi/lo split of the fp32 inputs)
//     B1 row = [ W_hi(4), b_hi | W_hi(4), b_hi | W_lo(4), b_lo | 0 ]                    (same split of W1 and b1)
// so z1 = v_hi W_hi + v_lo W_hi + v_hi W_lo carries ~16 mantissa bits (the dropped v_lo W_lo term is 2^-16 relative) --
// tighter than a tf32 MMA -- and lands in TMEM.  CONVERTER warps do tcgen05.ld (one row per thread) -> sign bit ->
// ReLU -> bf16 -> 128B-swizzled A tile of layer 2 (~2.5 instructions per element; the BN affine is folded into W2').
// The 48 action-branch columns of the critic (one input each) stay on the CUDA cores.
//
// TMEM (512 columns): [0,128) [128,256) layer-2 accumulators (double buffered), [256,384) [384,512) layer-1 outputs of
// the two 128-column halves of a tile.  Warps: 0 MMA issuer (+W2^T TMA), 1..4 epilogue, 5..12 converters (two per lane
// quadrant, one 64-column k-block each), 13 input-tile producer.
// =========================================================================================================
namespace v2 {

constexpr int NSLOT2 = 6;
constexpr int NUM_THREADS2 = 32 + 128 + 256 + 32;
constexpr int OFF2_W = NSLOT2 * SLOT_BYTES;              //  96 KB ring
constexpr int OFF2_B1 = OFF2_W + W_BYTES;                // + 80 KB W2^T
constexpr int B1_BYTES = 2 * 256 * 16;                   //   8 KB  W1ext, no-swizzle K-major: [chunk][row][16 B]
constexpr int OFF2_X = OFF2_B1 + B1_BYTES;
constexpr int X_BYTES = 2 * TILE_M * 16;                 //   4 KB per buffer: [chunk][row][16 B]
constexpr int OFF2_CTAB = OFF2_X + 2 * X_BYTES;          // converter tables (action branch): wa[64], ba[64]
constexpr int OFF2_ETAB = OFF2_CTAB + 128 * 4;
constexpr int OFF2_SCRATCH = OFF2_ETAB + 2080;
constexpr int OFF2_BAR = OFF2_SCRATCH + 4 * 32 * 33 * 4;
constexpr int SMEM2_BYTES = OFF2_BAR + 512 + 1024;
static_assert(SMEM2_BYTES <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
static_assert(OFF2_SCRATCH % 16 == 0 && OFF2_BAR % 8 == 0, "alignment");

// no-swizzle K-major descriptor: 8-row x 16-byte core matrices; LBO = distance between the two K chunks,
// SBO = distance between 8-row groups (128 B: rows are packed 16 B apart)
__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ uint32_t pack2(bf16 a, bf16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

__global__ void __launch_bounds__(NUM_THREADS2, 1) fused_forward_v2_kernel(const __grid_constant__ CUtensorMap tmW, Args g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* ctab = reinterpret_cast<float*>(smem + OFF2_CTAB);     // wa[64] | ba[64]
    float* etab = reinterpret_cast<float*>(smem + OFF2_ETAB);
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + OFF2_BAR);
    uint64_t* a_empty = a_full + NSLOT2;
    uint64_t* acc_full = a_empty + NSLOT2;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* w_full = acc_empty + 2;
    uint64_t* z1_full = w_full + 1;
    uint64_t* z1_empty = z1_full + 2;
    uint64_t* x_full = z1_empty + 2;
    uint64_t* x_empty = x_full + 2;
    uint64_t* b1_full = x_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b1_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const avd_net_dims d = g.d;
    const int F = g.critic ? d.l1 + d.la : d.l1;
    const int nkb = (F + KB - 1) / KB;              // 4 (actor) or 5 (critic)
    const int tile_begin = blockIdx.x * g.tiles_per_cta;
    const int tile_end = min(g.total_tiles, tile_begin + g.tiles_per_cta);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmW);
        for (int i = 0; i < NSLOT2; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4);
            mbar_init(&z1_full[i], 1); mbar_init(&z1_empty[i], 8);
            mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1);
        }
        mbar_init(w_full, 1);
        mbar_init(b1_full, 8);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ========================================== MMA issuer ==========================================
        if (lane == 0 && tile_begin < tile_end) {
            constexpr uint32_t idesc = make_idesc_bf16(TILE_M, L2N, false, false);
            const uint32_t w_addr = smem_u32(smem + OFF2_W);
            const uint32_t b1_addr = smem_u32(smem + OFF2_B1);
            const uint32_t x_addr = smem_u32(smem + OFF2_X);
            uint32_t kc = 0, wph = 0, b1ph = 0;
            int agent_w = -1, agent_b1 = -1;
            const int ntiles = tile_end - tile_begin;
            // layer-1 MMA of half h of local tile tcx -> TMEM columns 256 + 128 h
            auto issue_l1 = [&](int tcx, int h) {
                const int agent = (tile_begin + tcx) / g.tiles_per_agent;
                if (agent != agent_b1) {         // converters rebuild W1ext when they reach this tile
                    mbar_wait(b1_full, b1ph);
                    b1ph ^= 1;
                    agent_b1 = agent;
                }
                const int xb = tcx & 1;
                if (h == 0) mbar_wait(&x_full[xb], ((uint32_t)tcx >> 1) & 1);
                mbar_wait(&z1_empty[h], ((uint32_t)tcx & 1) ^ 1);
                tc_fence_after();
                mma_bf16(tmem_base + 256u + (uint32_t)(h * L2N), make_desc_noswz(x_addr + xb * X_BYTES, TILE_M * 16, 128),
                         make_desc_noswz(b1_addr + h * (L2N * 16), 256 * 16, 128), idesc, 0);
                mma_commit(&z1_full[h]);
                if (h == 1) mma_commit(&x_empty[xb]);
            };
            auto issue_l2 = [&](int tc, int kb_lo, int kb_hi) {
                const uint32_t tacc = tmem_base + (uint32_t)((tc & 1) * L2N);
                for (int kb = kb_lo; kb < kb_hi; ++kb, ++kc) {
                    const uint32_t slot = kc % NSLOT2, n = kc / NSLOT2;
                    mbar_wait(&a_full[slot], n & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + slot * SLOT_BYTES);
                    const int nm = min(4, (F - kb * KB) / 16);
                    for (int j = 0; j < nm; ++j)
                        mma_bf16(tacc, make_smem_desc(a_addr + j * 32, 16, 1024), make_smem_desc(w_addr + kb * (L2N * KB * 2) + j * 32, 16, 1024), idesc,
                                 (kb | j) != 0);
                    mma_commit(&a_empty[slot]);
                }
            };
            issue_l1(0, 0);
            issue_l1(0, 1);
            for (int tc = 0; tc < ntiles; ++tc) {
                const int agent = (tile_begin + tc) / g.tiles_per_agent;
                if (agent != agent_w) {
                    if (tc > 0) mbar_wait(&acc_full[(tc - 1) & 1], ((uint32_t)(tc - 1) >> 1) & 1);   // old W2^T no longer read
                    mbar_expect_tx(w_full, (uint32_t)nkb * L2N * KB * 2);
                    for (int kb = 0; kb < nkb; ++kb) tma_load_3d(smem + OFF2_W + kb * (L2N * KB * 2), &tmW, w_full, kb * KB, 0, agent);
                    mbar_wait(w_full, wph);
                    wph ^= 1;
                    agent_w = agent;
                }
                mbar_wait(&acc_empty[tc & 1], (((uint32_t)tc >> 1) & 1) ^ 1);
                tc_fence_after();
                if (tc + 1 < ntiles) issue_l1(tc + 1, 0);
                issue_l2(tc, 0, 2);
                if (tc + 1 < ntiles) issue_l1(tc + 1, 1);
                issue_l2(tc, 2, nkb);
                mma_commit(&acc_full[tc & 1]);
            }
        }
    } else if (warp <= 4) {
        epilogue_role(g, smem, etab, reinterpret_cast<float*>(smem + OFF2_SCRATCH), acc_full, acc_empty, tmem_base, warp, lane, tile_begin, tile_end);
    } else if (warp <= 12) {
        // ========================================== converters ==========================================
        const int cw = warp - 5;             // 0..7
        const int q = warp & 3;              // TMEM lane quadrant of this warp
        const int sub = cw >> 2;             // which 64-column half of a 128-column layer-1 half
        const int ct = threadIdx.x - 160;    // 0..255
        const int row = q * 32 + lane;       // row of the tile owned by this thread
        float* wa_tab = ctab;
        float* ba_tab = wa_tab + 64;
        uint32_t kc_base = 0;
        int cur_agent = -1;
        for (int t = tile_begin, tc = 0; t < tile_end; ++t, ++tc) {
            const int agent = t / g.tiles_per_agent;
            const int tile_in_agent = t - agent * g.tiles_per_agent;
            if (agent != cur_agent) {
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float* P = g.params + (int64_t)agent * g.pstride;
                int64_t oW, ob;
                if (g.critic) { const CriticOff o = critic_off(d); oW = o.Ws; ob = o.bs; }
                else { const ActorOff o = actor_off(d); oW = o.W1; ob = o.b1; }
                {   // thread ct owns layer-1 output column ct (l1 == 256): W1ext row
                    const int n = ct;
                    bf16 whi[4], wlo[4], bhi, blo;
#pragma unroll
                    for (int k = 0; k < 4; ++k) split_bf16(k < d.ns ? P[oW + (int64_t)k * d.l1 + n] : 0.0f, whi[k], wlo[k]);
                    split_bf16(P[ob + n], bhi, blo);
                    const bf16 zero = __float2bfloat16_rn(0.0f);
                    // K index: 0..4 = (W_hi, b_hi) x v_hi | 5..9 = (W_hi, b_hi) x v_lo | 10..14 = (W_lo, b_lo) x v_hi | 15 = 0
                    const uint4 c0 = make_uint4(pack2(whi[0], whi[1]), pack2(whi[2], whi[3]), pack2(bhi, whi[0]), pack2(whi[1], whi[2]));
                    const uint4 c1 = make_uint4(pack2(whi[3], bhi), pack2(wlo[0], wlo[1]), pack2(wlo[2], wlo[3]), pack2(blo, zero));
                    *reinterpret_cast<uint4*>(smem + OFF2_B1 + n * 16) = c0;
                    *reinterpret_cast<uint4*>(smem + OFF2_B1 + 256 * 16 + n * 16) = c1;
                }
                if (g.critic && ct < 64) {   // action branch (la <= 64 columns)
                    const CriticOff o = critic_off(d);
                    float wa = 0.f, ba = 0.f;
                    if (ct < d.la) { wa = P[o.Wa + ct]; ba = P[o.ba + ct]; }
                    wa_tab[ct] = wa; ba_tab[ct] = ba;
                }
                fence_proxy_async();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (lane == 0) mbar_arrive(b1_full);
                cur_agent = agent;
            }
            const int64_t r_in = (int64_t)tile_in_agent * TILE_M + row;
            const bool rvalid = r_in < g.R;
            const int64_t nrow = (int64_t)agent * g.R + (rvalid ? r_in : g.R - 1);
            const float a_val = g.critic ? g.act[nrow] : 0.0f;
            const int64_t blk_row = (int64_t)tile_in_agent * TILE_M + q * 32;       // first row of my warp's 32-row block
            const int rows_here = (int)max((int64_t)0, min((int64_t)32, g.R - blk_row));

            // write one finished 64-column k-block of my row into ring slot `slot` (+ coalesced copy to H_out)
            auto emit = [&](const float* hv, int ncols, int kb, uint32_t slot, uint32_t neg0, uint32_t neg1) {
                uint8_t* slot_base = smem + slot * SLOT_BYTES;
                if (g.mask_out && rvalid) {
                    uint32_t* mrow = g.mask_out + nrow * g.mask_words + 2 * kb;
                    mrow[0] = neg0;
                    mrow[1] = neg1;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j * 8 < ncols) {
                        const uint4 pk = make_uint4(pack_bf16x2(hv[8 * j], hv[8 * j + 1]), pack_bf16x2(hv[8 * j + 2], hv[8 * j + 3]),
                                                    pack_bf16x2(hv[8 * j + 4], hv[8 * j + 5]), pack_bf16x2(hv[8 * j + 6], hv[8 * j + 7]));
                        *reinterpret_cast<uint4*>(slot_base + row * 128 + ((j ^ (row & 7)) << 4)) = pk;      // SWIZZLE_128B
                    }
                }
                if (g.H_out) {   // my warp re-reads its own 32 rows with lane = (row-in-quad, chunk) so that stores coalesce
                    __syncwarp();
                    const int ch = lane & 7, rsub = lane >> 3;
                    if (ch * 8 < ncols) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = i * 4 + rsub;
                            const int trow = q * 32 + rr;
                            const uint4 pk = *reinterpret_cast<const uint4*>(slot_base + trow * 128 + ((ch ^ (trow & 7)) << 4));
                            if (rr < rows_here)
                                *reinterpret_cast<uint4*>(g.H_out + ((int64_t)agent * g.R + blk_row + rr) * F + kb * KB + ch * 8) = pk;
                        }
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[slot]);
            };

            for (int h = 0; h < 2; ++h) {
                const int kb = 2 * h + sub;
                const uint32_t kcx = kc_base + (uint32_t)kb;
                const uint32_t slot = kcx % NSLOT2, n = kcx / NSLOT2;
                mbar_wait(&z1_full[h], (uint32_t)tc & 1);
                tc_fence_after();
                float z[64];
                const uint32_t taddr = tmem_base + 256u + (uint32_t)(h * L2N + sub * 64) + ((uint32_t)(q * 32) << 16);
                tmem_ld32(taddr, z);
                tmem_ld32(taddr + 32, z + 32);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&z1_empty[h]);          // TMEM half may be overwritten by the next tile's layer-1 MMA
                uint32_t neg0 = 0, neg1 = 0;     // sign bits of z1: one funnel shift per element
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    neg0 = __funnelshift_l(__float_as_uint(z[j]), neg0, 1);
                    neg1 = __funnelshift_l(__float_as_uint(z[32 + j]), neg1, 1);
                    z[j] = fmaxf(z[j], 0.0f);
                    z[32 + j] = fmaxf(z[32 + j], 0.0f);
                }
                mbar_wait(&a_empty[slot], (n & 1) ^ 1);
                emit(z, 64, kb, slot, neg0, neg1);
            }
            if (g.critic && sub == 1) {     // action-branch k-block (48 columns, one input): CUDA cores
                const int kb = 4;
                const uint32_t kcx = kc_base + (uint32_t)kb;
                const uint32_t slot = kcx % NSLOT2, n = kcx / NSLOT2;
                float hv[64];
                uint32_t neg0 = 0, neg1 = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float z0 = fmaf(a_val, wa_tab[j], ba_tab[j]), z1 = fmaf(a_val, wa_tab[32 + j], ba_tab[32 + j]);
                    neg0 = __funnelshift_l(__float_as_uint(z0), neg0, 1);
                    neg1 = __funnelshift_l(__float_as_uint(z1), neg1, 1);
                    hv[j] = fmaxf(z0, 0.0f);
                    hv[32 + j] = fmaxf(z1, 0.0f);
                }
                mbar_wait(&a_empty[slot], (n & 1) ^ 1);
                emit(hv, d.la, kb, slot, neg0, neg1);
            }
            kc_base += (uint32_t)nkb;
        }
    } else {
        // ========================================== input-tile producer (warp 13) ==========================================
        for (int t = tile_begin, tc = 0; t < tile_end; ++t, ++tc) {
            const int agent = t / g.tiles_per_agent;
            const int tile_in_agent = t - agent * g.tiles_per_agent;
            const int xb = tc & 1;
            mbar_wait(&x_empty[xb], (((uint32_t)tc >> 1) & 1) ^ 1);
            uint8_t* xbase = smem + OFF2_X + xb * X_BYTES;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = i * 32 + lane;
                const int64_t r_in = (int64_t)tile_in_agent * TILE_M + row;
                const int64_t nrow = (int64_t)agent * g.R + (r_in < g.R ? r_in : g.R - 1);
                bf16 hi[5], lo[5];
#pragma unroll
                for (int k = 0; k < 4; ++k) split_bf16(k < d.ns ? g.s[nrow * g.s_rs + k * g.s_cs] : 0.0f, hi[k], lo[k]);
                hi[4] = __float2bfloat16_rn(1.0f);
                lo[4] = __float2bfloat16_rn(0.0f);
                // K index: 0..4 = v_hi | 5..9 = v_lo | 10..14 = v_hi | 15 = 0
                const uint4 c0 = make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], lo[0]), pack2(lo[1], lo[2]));
                const uint4 c1 = make_uint4(pack2(lo[3], lo[4]), pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], lo[4]));
                *reinterpret_cast<uint4*>(xbase + row * 16) = c0;
                *reinterpret_cast<uint4*>(xbase + TILE_M * 16 + row * 16) = c1;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&x_full[xb]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace v2

static PFN_cuTensorMapEncodeTiled encode_fn() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    }
    return fn;
}

bool supported(const avd_net_dims& d, bool critic) {
    const int F = critic ? d.l1 + d.la : d.l1;
    return d.l2 == L2N && d.ns <= 4 && d.l1 == 256 && F % 16 == 0 && F <= MAX_KB * KB && d.la % 8 == 0 && d.la <= 64;
}

// W2T: bf16 [A][128][F] (K-major copy of the BN-folded layer-2 kernel) and b2f [A][128]: see pack_fold_kernel
int forward(const avd_net_dims& d, bool critic, int A, int64_t R, const float* params, int64_t pstride, const bf16* W2T, const float* b2f,
            const float* s, int64_t s_rs, int64_t s_cs, const float* act, bf16* H_out, uint32_t* mask_out, float* Z_out, int head,
            const float* rew, float gamma, float high, float* out, cudaStream_t st, bf16* DZ_out, float* loss) {
    if (!supported(d, critic)) {
        set_error("fused forward kernel does not support these layer sizes");
        return AVD_ERR_UNSUPPORTED;
    }
    PFN_cuTensorMapEncodeTiled enc = encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return AVD_ERR_CUDA;
    }
    static bool attr_set = false;
    if (!attr_set) {
        AVD_CUDA_OK(cudaFuncSetAttribute(v2::fused_forward_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::SMEM2_BYTES));
        attr_set = true;
    }
    const int F = critic ? d.l1 + d.la : d.l1;
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)F, (cuuint64_t)L2N, (cuuint64_t)A};
    cuuint64_t strides[2] = {(cuuint64_t)F * 2, (cuuint64_t)F * L2N * 2};
    cuuint32_t box[3] = {KB, L2N, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(W2T), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(W2T) failed with %d", (int)r);
        return AVD_ERR_CUDA;
    }
    Args g;
    g.d = d; g.critic = critic ? 1 : 0; g.A = A; g.R = R; g.params = params; g.pstride = pstride; g.b2f = b2f;
    g.s = s; g.s_rs = s_rs; g.s_cs = s_cs; g.act = act; g.H_out = H_out; g.mask_out = mask_out; g.mask_words = 2 * ((F + KB - 1) / KB);
    g.Z_out = Z_out; g.head = head; g.rew = rew;
    g.gamma = gamma; g.high = high; g.out = out; g.DZ_out = DZ_out; g.loss = loss;
    g.tiles_per_agent = (int)((R + TILE_M - 1) / TILE_M);
    g.total_tiles = g.tiles_per_agent * A;
    const int ctas = std::min(g.total_tiles, sm_count());
    g.tiles_per_cta = (g.total_tiles + ctas - 1) / ctas;
    const int grid = (g.total_tiles + g.tiles_per_cta - 1) / g.tiles_per_cta;
    v2::fused_forward_v2_kernel<<<grid, v2::NUM_THREADS2, v2::SMEM2_BYTES, st>>>(tm, g);
    AVD_LAUNCH_OK();
    return AVD_OK;
}

}  // namespace fused
}  // namespace avd
