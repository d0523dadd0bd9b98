// Yacht-Auction B200 engine -- the WHOLE leaf-evaluator forward (YachtNNet.forward,
// yacht/pytorch/YachtNNet.py:62-70) as one persistent tcgen05 kernel: state_to_vec features in, bf16 policy
// logits (padded to 3232 columns), tanh values and each row's largest logit out.  A CTA owns 128 leaves from
// the first Linear to the last: the skip connection (float32) and the accumulators live in TMEM, every weight
// matrix arrives as a pre-swizzled image through cp.async.bulk while an earlier stage computes, and all bias /
// SiLU / LayerNorm / residual / tanh work happens in the tcgen05.ld epilogues.  Stages:
//   input   Linear(59->256) + LN + SiLU                         (1 K-block of 64, N = 256)
//   trunk   nblocks x [LN(SiLU(fc1)), skip + LN(SiLU(fc2))]     (4 K-blocks; two N = 128 halves per layer, each with
//           its own completion barrier: the first epilogue pass over one half runs under the other half's MMAs;
//           activations go back to shared memory as the next bf16, 128-byte-swizzled K-major A operand)
//   value   SiLU(LN_v(h)) -> Linear(256->128) + SiLU -> dot(w2) + b2 -> tanh      (N = 128)
//   policy  SiLU(LN_pi(h)) written ONCE to tensor memory (A operand from TMEM) -> 26 tiles of 128 columns:
//           the freed A tile + the weight region = three 64 KB weight slots, three TMEM accumulators, full /
//           drained mbarriers; warp 15 only produces (copies, MMAs), the other 15 warps run the epilogue
//           (bias, running row maximum, bf16 packing, 256-bit stores)
// Single-thread instructions (MMA, commit, bulk copy) are issued under elect.sync so they compile to
// straight-line SASS.  Every row is computed independently of the batch it sits in (fixed tile shapes, fixed
// accumulation order), so the evaluator is batch-invariant: sharding leaves over GPUs or waves cannot change a bit.
// -DYA_FWD_TIMELINE builds the profiling variant used by profiles/tools/forward_timeline.py.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <atomic>
#include <cstdint>
#include "../../include/yacht_b200.h"
#include "ya_tc.cuh"

namespace {

using namespace ya_tc;

constexpr int kRows = 128, kDim = 256, kParts = 4;
constexpr int kThreads = 512;                       // 16 warps: warp w owns TMEM lanes 32 (w % 4).., column part w / 4
constexpr int kIssuerWarp = 15;                     // its elected lane issues every bulk copy and every MMA
constexpr int kFeat = 59;
constexpr int kPolicyCols = 3232, kPolicyTile = 128, kPolicyTiles = 26;        // 26 * 128 = 3328 >= 3232
constexpr int kABytes = kRows * kDim * 2;            // 64 KB
constexpr int kWBytes = kDim * kDim * 2;             // 128 KB (two 64 KB halves for the N = 128 stages)
constexpr int kPrmFloats = 776;                      // largest parameter block (value head), 16-byte multiple
constexpr int kPiPrmFloats = 2 * kDim + kPolicyTiles * kPolicyTile;            // gamma_pi | beta_pi | bias of all 3,328 columns
constexpr int kActions = 3226;
constexpr int kSmemBytes = 1024 + kABytes + kWBytes + 2 * kPrmFloats * 4 + kPiPrmFloats * 4 + kRows * kParts * 8 + 128;

struct Blob {                                        // byte / float offsets of the host-built blobs (see mcts.py)
    int64_t w_in, w_trunk, w_v, w_pi;
    int64_t p_in, p_trunk, p_v, p_pi_ln, p_pi_bias;
};

#ifdef YA_FWD_TIMELINE                                 // profiling build only (profiles/tools/forward_timeline.py)
__device__ unsigned long long g_timeline[1024];
#define YA_STAMP() do { if (blockIdx.x == 0 && tid == 0 && tl_n < 1024) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_timeline[tl_n++] = t_; } } while (0)
#else
#define YA_STAMP() do { } while (0)
#endif

__device__ __forceinline__ uint32_t a_tile_offset(int r, int c8) {
    int kb = c8 >> 3, chunk = c8 & 7;
    return (uint32_t)(kb * (kRows * 128) + r * 128 + ((chunk ^ (r & 7)) << 4));
}

template <bool F16>
#ifdef YA_EXP_TRUNK_NO_TANH                            // profiling experiment: the trunk epilogue without its MUFU op
#define silu_from_half(x) ((x) * 1.0009765625f)
#endif
__device__ __forceinline__ void pack_store_a(uint8_t* a_tile, int row, int c8_first, const uint32_t (&r)[32]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t p[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = pack2<F16>(__uint_as_float(r[q * 8 + 2 * i]), __uint_as_float(r[q * 8 + 2 * i + 1]));
        *reinterpret_cast<uint4*>(a_tile + a_tile_offset(row, c8_first + q)) = make_uint4(p[0], p[1], p[2], p[3]);
    }
}

// F16: operands (activations, weights) and logits in IEEE half -- the precision of the reference's CUDA predict (fp16
// autocast, yacht/NNet.py:186-193) -- instead of bfloat16; accumulation, LayerNorm and the skip connection are float32
// either way.  tcgen05.mma kind::f16 runs both formats at the same rate.
template <bool F16>
__global__ void __launch_bounds__(kThreads, 1)
ya_k_forward(const float* __restrict__ features, uint16_t* __restrict__ logits, float* __restrict__ values,
             float* __restrict__ row_max, const uint8_t* __restrict__ wblob, const float* __restrict__ pblob, Blob off, int nblocks, int64_t n, float eps,
             const uint64_t* __restrict__ scatter_dst, const uint32_t* __restrict__ scatter_desc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_tile = base;
    uint8_t* w_tile = base + kABytes;
    float* prm_all = reinterpret_cast<float*>(w_tile + kWBytes);
    float* pi_prm = prm_all + 2 * kPrmFloats;
    float2* xchg = reinterpret_cast<float2*>(pi_prm + kPiPrmFloats);
    uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + kParts * kRows);     // [0] weights landed, [1] / [2] MMA done,
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);            // policy head: [3..5] weight slot landed,
                                                                             // [6..8] accumulator ready, [9..11] drained

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane;
    const int part = warp >> 2;
    const int64_t grow = (int64_t)blockIdx.x * kRows + row;
    const bool producer = warp == kIssuerWarp;                        // also an epilogue warp, except in the policy head
    constexpr bool worker = true;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        mbar_init(&bars[3], 1);
        mbar_init(&bars[4], 1);
        for (int i = 5; i < 9; ++i) mbar_init(&bars[i], 1);
        for (int i = 9; i < 12; ++i) mbar_init(&bars[i], 15);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    // This thread's 64 of a row's 256 columns: 32 from each half, so that the epilogue can start on the first
    // N = 128 half of a trunk layer while the tensor core is still on the second.
    const int colv[2] = {part * 32, 128 + part * 32};
    const uint32_t t_skip = t_lane + 256;                             // float32 skip connection: TMEM columns 256..511
    uint32_t w_phase = 0, m_phase = 0, m2_phase = 0;
#ifdef YA_FWD_TIMELINE
    int tl_n = 0;
#endif
    int stage = 0;                                                    // parameter double buffer index = stage & 1

    // one weight image + one parameter block per stage, on one transaction barrier
    auto load_stage = [&](uint8_t* w_dst, int64_t w_src, uint32_t w_bytes, float* p_dst, int64_t p_src, uint32_t p_floats) {
        mbar_expect_tx(&bars[0], w_bytes + p_floats * 4);
        for (uint32_t o = 0; o < w_bytes; o += 32768) bulk_g2s(w_dst + o, wblob + w_src + o, min(32768u, w_bytes - o), &bars[0]);
        if (p_floats) bulk_g2s(p_dst, pblob + p_src, p_floats * 4, &bars[0]);
    };
    auto prm_buf = [&](int buf) { return prm_all + (buf & 1) * kPrmFloats; };
    auto policy_slot = [](int j) { return (2 + 2 * j) % 3; };          // tile j -> 64 KB slot: 2, 1, 0, 2, 1, 0, ...
    // MMA of one stage: A tile (n_kb K-blocks of 64) x weight image at w_src (rows = n_cols) -> TMEM column d_col
    auto run_mma = [&](const uint8_t* w_src, int n_kb, int n_cols, uint32_t d_col) {
        proxy_fence();                                                // A tile written through the generic proxy
        tc_fence_before();
        __syncthreads();
        YA_STAMP();                                                   // [3k] epilogue of the previous stage done
        mbar_wait(&bars[0], w_phase);                                 // weights + parameters landed
        w_phase ^= 1;
        YA_STAMP();                                                   // [3k+1] weights landed
        if (producer) {
            tc_fence_after();
            const uint64_t da = umma_desc(smem_u32(a_tile)), db = umma_desc(smem_u32(w_src));
            const uint32_t idesc = umma_idesc(n_cols, F16);
            if (elect_one()) {
                for (int kb = 0; kb < n_kb; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma(tmem + d_col, umma_desc_advance(da, kb * (kRows * 128) + k * 32),
                             umma_desc_advance(db, kb * (n_cols * 128) + k * 32), (uint32_t)((kb | k) != 0), idesc);
                umma_commit(&bars[1]);
            }
            __syncwarp();
        }
        mbar_wait(&bars[1], m_phase);                                 // accumulator ready; A tile and this weight buffer free
        m_phase ^= 1;
        tc_fence_after();
        YA_STAMP();                                                   // [3k+2] MMA done
    };
    // the four warps that share a TMEM lane quarter (same 32 rows, different column parts) exchange row statistics
    auto worker_sync = [&] { asm volatile("bar.sync %0, 128;" ::"r"(1 + (warp & 3)) : "memory"); };
    auto row_stats = [&](float s, float ss, float& mean, float& rstd) {  // LayerNorm statistics over the 4 threads of a row
        xchg[part * kRows + row] = make_float2(s, ss);
        worker_sync();
#pragma unroll
        for (int p = 1; p < kParts; ++p) {
            float2 o = xchg[((part + p) % kParts) * kRows + row];
            s += o.x; ss += o.y;
        }
        mean = s * (1.0f / kDim);
        rstd = rsqrtf(fmaxf(ss * (1.0f / kDim) - mean * mean, 0.0f) + eps);
        worker_sync();                                                // xchg may be rewritten by the next stage
    };

    // ---------------------------------------------------------------- input stage
    if (producer && elect_one()) load_stage(w_tile, off.w_in, 256 * 128, prm_buf(0), off.p_in, 3 * kDim);
    if (worker) {   // features (float32 [n][59]) -> bf16, K padded to 64: this thread fills chunks 2*part, 2*part+1 of K-block 0
        uint32_t f[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            int c = part * 16 + i;
            f[i] = (grow < n && c < kFeat) ? __float_as_uint(features[grow * kFeat + c]) : 0u;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            uint32_t p[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = pack2<F16>(__uint_as_float(f[q * 8 + 2 * i]), __uint_as_float(f[q * 8 + 2 * i + 1]));
            *reinterpret_cast<uint4*>(a_tile + a_tile_offset(row, part * 2 + q)) = make_uint4(p[0], p[1], p[2], p[3]);
        }
    }
    run_mma(w_tile, 1, kDim, 0);
    if (producer && elect_one()) load_stage(w_tile, off.w_trunk, kWBytes, prm_buf(1), off.p_trunk, 3 * kDim);      // first trunk layer streams in
    if (worker) {   // h = SiLU(LN(z + b)): Linear -> LayerNorm -> SiLU (YachtNNet.py:25-30); also the first skip connection
        const float* prm = prm_all;
        float ps[4] = {0, 0, 0, 0}, pq[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(t_lane + colv[c], r);
            tmem_ld_wait();
            const float* bias = prm + colv[c];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float v = __uint_as_float(r[i]) + bias[i];
                ps[i & 3] += v; pq[i & 3] = fmaf(v, v, pq[i & 3]);
                r[i] = __float_as_uint(v);
            }
            tmem_st32(t_lane + colv[c], r);
        }
        tmem_st_wait();
        float mean, rstd;
        row_stats((ps[0] + ps[1]) + (ps[2] + ps[3]), (pq[0] + pq[1]) + (pq[2] + pq[3]), mean, rstd);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(t_lane + colv[c], r);
            tmem_ld_wait();
            const float* gamma = prm + kDim + colv[c];
            const float* beta = prm + 2 * kDim + colv[c];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float ga = rstd * gamma[i];
                float y = fmaf(__uint_as_float(r[i]), ga, fmaf(-mean, ga, beta[i]));
                r[i] = __float_as_uint(silu_from_half(0.5f * y));
            }
            tmem_st32(t_skip + colv[c], r);
            pack_store_a<F16>(a_tile, row, colv[c] / 8, r);
        }
        tmem_st_wait();
    }
    stage = 1;

    // ---------------------------------------------------------------- residual trunk
    const int layers = 2 * nblocks;
    for (int l = 0; l < layers; ++l, ++stage) {
        const float* prm = prm_all + (stage & 1) * kPrmFloats;
        const bool second = l & 1;                                    // fc2: add the skip connection
        // Two N = 128 halves, each with its own completion barrier: the epilogue's first pass over columns
        // 0..127 runs while the tensor core works on columns 128..255.
        proxy_fence();
        tc_fence_before();
        __syncthreads();
        YA_STAMP();
        mbar_wait(&bars[0], w_phase);
        w_phase ^= 1;
        YA_STAMP();
        if (producer) {
            tc_fence_after();
            const uint64_t da = umma_desc(smem_u32(a_tile)), db = umma_desc(smem_u32(w_tile));
            const uint32_t idesc = umma_idesc(128, F16);
            if (elect_one()) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma(tmem + half * 128, umma_desc_advance(da, kb * (kRows * 128) + k * 32),
                                 umma_desc_advance(db, kb * (kDim * 128) + half * (128 * 128) + k * 32), (uint32_t)((kb | k) != 0), idesc);
                    umma_commit(&bars[1 + half]);
                }
            }
            __syncwarp();
        }
        // prm = bias / 2 | gamma | beta (the host halves the bias: SiLU(x) = t + t * tanh(t), t = x / 2).  The 64
        // activations stay in registers across the statistics exchange (no TMEM round trip).
        uint32_t v0[32], v1[32];
        float ps[4] = {0, 0, 0, 0}, pq[4] = {0, 0, 0, 0};
        mbar_wait(&bars[1], m_phase);
        m_phase ^= 1;
        tc_fence_after();
        tmem_ld32(t_lane + colv[0], v0);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float x = silu_from_half(fmaf(__uint_as_float(v0[i]), 0.5f, prm[colv[0] + i]));
            ps[i & 3] += x; pq[i & 3] = fmaf(x, x, pq[i & 3]);
            v0[i] = __float_as_uint(x);
        }
        mbar_wait(&bars[2], m2_phase);
        m2_phase ^= 1;
        tc_fence_after();
        YA_STAMP();
        if (producer && elect_one()) {                                // next stage's weights under the rest of this epilogue
            if (l + 1 < layers) load_stage(w_tile, off.w_trunk + (int64_t)(l + 1) * kWBytes, kWBytes, prm_buf(stage + 1),
                                           off.p_trunk + (int64_t)(l + 1) * 3 * kDim, 3 * kDim);
            else {                                                    // value head weights; both heads' parameters
                mbar_expect_tx(&bars[0], 65536 + 772 * 4 + 2 * kDim * 4);
                bulk_g2s(w_tile, wblob + off.w_v, 32768, &bars[0]);
                bulk_g2s(w_tile + 32768, wblob + off.w_v + 32768, 32768, &bars[0]);
                bulk_g2s(prm_buf(stage + 1), pblob + off.p_v, 772 * 4, &bars[0]);
                bulk_g2s(pi_prm, pblob + off.p_pi_ln, 2 * kDim * 4, &bars[0]);
            }
        }
        __syncwarp();
        tmem_ld32(t_lane + colv[1], v1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float x = silu_from_half(fmaf(__uint_as_float(v1[i]), 0.5f, prm[colv[1] + i]));
            ps[i & 3] += x; pq[i & 3] = fmaf(x, x, pq[i & 3]);
            v1[i] = __float_as_uint(x);
        }
        float mean, rstd;
        row_stats((ps[0] + ps[1]) + (ps[2] + ps[3]), (pq[0] + pq[1]) + (pq[2] + pq[3]), mean, rstd);
        const float nm = -mean * rstd;
        const float* gamma = prm + kDim;
        const float* beta = prm + 2 * kDim;
        if (second) {                                                 // h += LN(SiLU(fc2(..))); the sum is the next skip
            uint32_t sk[32];
            tmem_ld32(t_skip + colv[0], sk);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
                v0[i] = __float_as_uint(fmaf(fmaf(__uint_as_float(v0[i]), rstd, nm), gamma[colv[0] + i], beta[colv[0] + i]) + __uint_as_float(sk[i]));
            tmem_st32(t_skip + colv[0], v0);
            pack_store_a<F16>(a_tile, row, colv[0] / 8, v0);
            tmem_ld32(t_skip + colv[1], sk);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
                v1[i] = __float_as_uint(fmaf(fmaf(__uint_as_float(v1[i]), rstd, nm), gamma[colv[1] + i], beta[colv[1] + i]) + __uint_as_float(sk[i]));
            tmem_st32(t_skip + colv[1], v1);
            pack_store_a<F16>(a_tile, row, colv[1] / 8, v1);
            tmem_st_wait();
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v0[i] = __float_as_uint(fmaf(fmaf(__uint_as_float(v0[i]), rstd, nm), gamma[colv[0] + i], beta[colv[0] + i]));
            pack_store_a<F16>(a_tile, row, colv[0] / 8, v0);
#pragma unroll
            for (int i = 0; i < 32; ++i) v1[i] = __float_as_uint(fmaf(fmaf(__uint_as_float(v1[i]), rstd, nm), gamma[colv[1] + i], beta[colv[1] + i]));
            pack_store_a<F16>(a_tile, row, colv[1] / 8, v1);
        }
    }

    // ---------------------------------------------------------------- heads: a = SiLU(LN(h; gamma, beta)) from the skip
    // Both heads normalise the same h (YachtNNet.py:38-50): one statistics pass, then the value head's activations go
    // to the shared-memory A tile and the policy head's to tensor-memory columns [0, 128) as packed bf16 pairs
    // (A operand read from TMEM).
    auto head_prep = [&](const float* gv, const float* bv, const float* gp, const float* bp) {
        float ps[4] = {0, 0, 0, 0}, pq[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(t_skip + colv[c], r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) { float v = __uint_as_float(r[i]); ps[i & 3] += v; pq[i & 3] = fmaf(v, v, pq[i & 3]); }
        }
        float mean, rstd;
        row_stats((ps[0] + ps[1]) + (ps[2] + ps[3]), (pq[0] + pq[1]) + (pq[2] + pq[3]), mean, rstd);
        const float nm = -mean * rstd;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32], a[32];
            tmem_ld32(t_skip + colv[c], r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float x = fmaf(__uint_as_float(r[i]), rstd, nm);
                a[i] = __float_as_uint(silu_from_half(0.5f * fmaf(x, gv[colv[c] + i], bv[colv[c] + i])));
                r[i] = __float_as_uint(silu_from_half(0.5f * fmaf(x, gp[colv[c] + i], bp[colv[c] + i])));
            }
            pack_store_a<F16>(a_tile, row, colv[c] / 8, a);
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack2<F16>(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            tmem_st16(t_lane + (uint32_t)(colv[c] / 2), pk);           // K elements colv[c].. = packed columns colv[c] / 2..
        }
        tmem_st_wait();
    };

    // value head (YachtNNet.py:44-50): LN -> SiLU -> Linear(256,128) -> SiLU -> Linear(128,1) -> tanh
    {
        const float* prm = prm_all + (stage & 1) * kPrmFloats;       // gamma_v | beta_v | b1[128] | w2[128] | b2
        mbar_wait(&bars[0], w_phase);                                 // the LayerNorm parameters travel with the weights
        head_prep(prm, prm + kDim, pi_prm, pi_prm + kDim);
        run_mma(w_tile, 4, 128, 128);                                 // accumulator in columns 128..255: 0..127 hold the policy A operand
        if (producer && elect_one()) {
            // The policy head reads its activations from tensor memory, so all 192 KB of operand space (A tile +
            // weight region) become three 64 KB weight slots; tiles 0..2 start streaming now.
            for (int j = 0; j < 3; ++j) {
                uint64_t* bar = &bars[3 + policy_slot(j)];
                mbar_expect_tx(bar, 65536 + (j == 0 ? kPolicyTiles * kPolicyTile * 4 : 0));
                bulk_g2s(base + policy_slot(j) * 65536, wblob + off.w_pi + (int64_t)j * 65536, 32768, bar);
                bulk_g2s(base + policy_slot(j) * 65536 + 32768, wblob + off.w_pi + (int64_t)j * 65536 + 32768, 32768, bar);
                if (j == 0) bulk_g2s(pi_prm + 2 * kDim, pblob + off.p_pi_bias, kPolicyTiles * kPolicyTile * 4, bar);   // every bias
            }
        }
        float dot = 0.0f;
        if (worker) {
            uint32_t r[32];
            tmem_ld32(t_lane + 128 + part * 32, r);
            tmem_ld_wait();
            const float* b1 = prm + 2 * kDim + part * 32;
            const float* w2 = prm + 2 * kDim + 128 + part * 32;
            float acc[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float t = silu_from_half(fmaf(__uint_as_float(r[i]), 0.5f, 0.5f * b1[i]));
                acc[i & 3] = fmaf(t, w2[i], acc[i & 3]);
            }
            dot = (acc[0] + acc[1]) + (acc[2] + acc[3]);
            xchg[part * kRows + row] = make_float2(dot, 0.0f);
        }
        __syncthreads();
        if (worker && part == 0 && grow < n) {
            float s = dot + xchg[1 * kRows + row].x + xchg[2 * kRows + row].x + xchg[3 * kRows + row].x + prm[2 * kDim + 256];
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(s));
            values[grow] = th;
        }
        __syncthreads();
        ++stage;
    }

    // policy head (YachtNNet.py:38-42): Linear(256, 3226) on the activations already sitting in tensor memory, as 26
    // tiles of 128 columns.  Weight tiles rotate over three 64 KB slots and accumulators over TMEM columns 128 /
    // 256 / 384: as soon as tile j's MMAs retire, tile j + 3's weights start streaming into the slot they read,
    // and the MMAs of the following tiles run under the epilogue of tile j.
    {
        mbar_wait(&bars[3 + policy_slot(0)], 0);                      // tile 0 and every bias landed
        const float* bias_all = pi_prm + 2 * kDim;
        auto load_tile = [&](int j) {                                 // producer thread, j >= 3
            uint64_t* bar = &bars[3 + policy_slot(j)];
            uint8_t* dst = base + policy_slot(j) * 65536;
            mbar_expect_tx(bar, 65536);
            bulk_g2s(dst, wblob + off.w_pi + (int64_t)j * 65536, 32768, bar);
            bulk_g2s(dst + 32768, wblob + off.w_pi + (int64_t)j * 65536 + 32768, 32768, bar);
        };
        auto issue_tile = [&](int j) {                                // producer thread: tile j's 16 MMAs
            const uint64_t db = umma_desc(smem_u32(base + policy_slot(j) * 65536));
            const uint32_t idesc = umma_idesc(kPolicyTile, F16);
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_ts(tmem + kPolicyTile + (j % 3) * kPolicyTile, tmem + kb * 32 + k * 8,
                            umma_desc_advance(db, kb * (kPolicyTile * 128) + k * 32), (uint32_t)((kb | k) != 0), idesc);
            umma_commit(&bars[6 + j % 3]);
        };
        proxy_fence();
        tc_fence_before();
        __syncthreads();                                              // the A tile (p) is complete
        tc_fence_after();
        float row_mx[2] = {-3.0e38f, -3.0e38f};
        // Scatter mode (the MCTS path): this row's LEGAL logits go straight into its leaf's row in the tree pool, in legal
        // order (the row layout ya_mcts_select allocated: csrc/ya_mcts.cu "ROWS_L16"), instead of a dense [n][3232] matrix.
        // dst = 0: the row needs no evaluation (its descent ended in a terminal / dead-end node).
        uint16_t* s_dst = nullptr;
        uint32_t s_desc = 0;
        if (scatter_dst && grow < n) {
            s_dst = reinterpret_cast<uint16_t*>(scatter_dst[grow]);
            s_desc = s_dst ? scatter_desc[grow] : 0u;
        }
        if (producer) {
            // Producer: keeps the tensor pipe fed.  Weight slots, accumulators (TMEM columns 128 / 256 / 384) and their
            // barriers all cycle with period 3: tile j needs its weights (requested two tiles ago) and the accumulator
            // drained by the epilogue of tile j - 3; once tile j is queued, tile j - 1 has retired and tile j + 2
            // streams into its slot.
            for (int j = 0; j < kPolicyTiles; ++j) {
                const uint32_t par = (uint32_t)((j / 3) & 1);
                mbar_wait(&bars[3 + policy_slot(j)], par);
                if (j >= 3) mbar_wait(&bars[9 + j % 3], par ^ 1u);
                tc_fence_after();
                if (elect_one()) issue_tile(j);
                __syncwarp();
                if (j >= 1 && j + 2 < kPolicyTiles) {
                    mbar_wait(&bars[6 + (j - 1) % 3], (uint32_t)(((j - 1) / 3) & 1));
                    if (elect_one()) load_tile(j + 2);
                    __syncwarp();
                }
            }
        } else {
            // 15 epilogue warps: the issuer warp's share (rows 96..127, column part 3) goes to warp 11 on top of its own
            const int n_my = (warp == kIssuerWarp - 4) ? 2 : 1;
            for (int j = 0; j < kPolicyTiles; ++j) {
                mbar_wait(&bars[6 + j % 3], (uint32_t)((j / 3) & 1));
                tc_fence_after();
                YA_STAMP();                                           // policy tile j: accumulator ready
                for (int q = 0; q < n_my; ++q) {
                    const int pp = part + q;
                    uint32_t r[32];
#ifdef YA_EXP_POLICY_NO_LD                                 // profiling experiment: the MMA / bulk-copy pipeline alone
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = 0u;
#else
                    tmem_ld32(t_lane + kPolicyTile + (j % 3) * kPolicyTile + pp * 32, r);
                    tmem_ld_wait();
#endif
                    if (q == n_my - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars[9 + j % 3]); // this warp's share of the accumulator is in registers
                    }
                    const int col0 = j * kPolicyTile + pp * 32;
#if defined(YA_EXP_POLICY_NO_LD) || defined(YA_EXP_POLICY_NO_ST)
                    if (col0 < kPolicyCols && r[0] == 0x7FC12345u) {   // profiling experiment: epilogue without its stores
#else
                    if (col0 < kPolicyCols) {
#endif
                        const float* bias = bias_all + col0;
                        float f[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]) + bias[i];
                        float m = row_mx[q];
                        if (col0 + 32 <= kActions) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) m = fmaxf(m, f[i]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (col0 + i < kActions) m = fmaxf(m, f[i]);
                        }
                        row_mx[q] = m;
                        if (s_desc) {
                            // The leaf's logit area keeps every logit at its column's position modulo 16 (csrc/ya_mcts.cu
                            // "Logit area"): a bid row is columns [0, 208); a ten-dice row has, per OPEN category c, a
                            // 272-slot copy of the 16-aligned column window around the category's run [202 + 252 c, +252).
                            // So each of this chunk's two 16-column sectors goes out as ONE aligned 32-byte store per block
                            // that wants it (a sector holding a run boundary may be wanted by both neighbours).
                            uint32_t pk[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) pk[i] = pack2<F16>(f[2 * i], f[2 * i + 1]);
                            auto store_sector = [&](uint16_t* dst, int h) {
                                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst),
                                             "r"(pk[8 * h]), "r"(pk[8 * h + 1]), "r"(pk[8 * h + 2]), "r"(pk[8 * h + 3]),
                                             "r"(pk[8 * h + 4]), "r"(pk[8 * h + 5]), "r"(pk[8 * h + 6]), "r"(pk[8 * h + 7]) : "memory");
                            };
                            if (s_desc & 1u) {                                             // bid row
#pragma unroll
                                for (int h = 0; h < 2; ++h)
                                    if (col0 + 16 * h < 208) store_sector(s_dst + col0 + 16 * h, h);
                            } else if (s_desc >> 13) {                                     // ten dice
                                const uint32_t open = (s_desc >> 1) & 0xFFFu;              // bit c = category c open
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    const int cs = col0 + 16 * h;
                                    const int ra = ((cs + 50) * 4162) >> 20, rb = ((cs + 65) * 4162) >> 20;   // runs of the sector's first / last column
#pragma unroll
                                    for (int t = 0; t < 2; ++t) {
                                        const int r = t ? rb : ra;                         // run r >= 1 = category r - 1
                                        if ((t == 0 || rb != ra) && r >= 1 && r <= 12 && ((open >> (r - 1)) & 1u)) {
                                            const int start = 202 + 252 * (r - 1);
                                            const int rank = __popc(open & ((1u << (r - 1)) - 1u));
                                            store_sector(s_dst + 272 * rank + (cs - (start & ~15)), h);
                                        }
                                    }
                                }
                            } else {                                                       // five dice: subset 0 of every open category
                                const uint32_t open = (s_desc >> 1) & 0xFFFu;
                                const int r0 = ((col0 + 50) * 4162) >> 20;
                                const int end0 = 202 + 252 * r0;                           // category r0 starts here, if inside this chunk
                                if (end0 < col0 + 32 && r0 < 12 && ((open >> r0) & 1u)) {
                                    uint16_t* one = s_dst + __popc(open & ((1u << r0) - 1u));
#pragma unroll
                                    for (int i = 0; i < 16; ++i)
                                        if (col0 + 2 * i == end0) *one = (uint16_t)(pk[i] & 0xFFFFu);
                                }
                            }
                        }
                        if (logits && grow < n) {
                            uint32_t p[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) p[i] = pack2<F16>(f[2 * i], f[2 * i + 1]);
                            uint16_t* dst = logits + grow * kPolicyCols + col0;         // 64 bytes, 32-byte aligned
#pragma unroll
                            for (int h2 = 0; h2 < 2; ++h2)            // 256-bit stores: half the LSU work of 4 x 16 bytes
                                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 16 * h2),
                                             "r"(p[8 * h2]), "r"(p[8 * h2 + 1]), "r"(p[8 * h2 + 2]), "r"(p[8 * h2 + 3]),
                                             "r"(p[8 * h2 + 4]), "r"(p[8 * h2 + 5]), "r"(p[8 * h2 + 6]), "r"(p[8 * h2 + 7]) : "memory");
                        }
                    }
                }
            }
            for (int q = 0; q < n_my; ++q) xchg[(part + q) * kRows + row].x = row_mx[q];
        }
        YA_STAMP();
        // the row's largest logit as the expand kernel will see it (rounding to 16 bits is monotone)
        __syncthreads();
        if (part == 0 && grow < n && row_max) {
            float m = fmaxf(fmaxf(row_mx[0], xchg[1 * kRows + row].x), fmaxf(xchg[2 * kRows + row].x, xchg[3 * kRows + row].x));
            row_max[grow] = round16<F16>(m);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace

#ifdef YA_FWD_TIMELINE
extern "C" int ya_debug_forward_timeline(unsigned long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(unsigned long long) * 1024);
}
#endif

extern "C" int ya_nn_forward(const float* features, void* logits16, float* values, float* row_max, const void* weight_blob,
                             const float* param_blob, const int64_t* offsets, int nblocks, int64_t n, float eps, int fp16,
                             const uint64_t* scatter_dst, const uint32_t* scatter_desc, void* stream) {
    if (n <= 0) return 0;
    if ((reinterpret_cast<uintptr_t>(logits16) | reinterpret_cast<uintptr_t>(weight_blob) |
         reinterpret_cast<uintptr_t>(param_blob)) & 31u) return (int)cudaErrorMisalignedAddress;
    if ((scatter_dst == nullptr) != (scatter_desc == nullptr) || (!logits16 && !scatter_dst)) return (int)cudaErrorInvalidValue;
    // the opt-in to > 48 KB of dynamic shared memory is a per-DEVICE function attribute: one flag per device and
    // kernel variant (set twice by racing threads is harmless; the flag is only published after the call succeeded)
    static std::atomic<bool> configured[2][64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    if (!configured[fp16 ? 1 : 0][dev].load(std::memory_order_acquire)) {
        e = fp16 ? cudaFuncSetAttribute(ya_k_forward<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)
                 : cudaFuncSetAttribute(ya_k_forward<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        configured[fp16 ? 1 : 0][dev].store(true, std::memory_order_release);
    }
    Blob off{offsets[0], offsets[1], offsets[2], offsets[3], offsets[4], offsets[5], offsets[6], offsets[7], offsets[8]};
    int blocks = (int)((n + kRows - 1) / kRows);
    if (fp16)
        ya_k_forward<true><<<blocks, kThreads, kSmemBytes, (cudaStream_t)stream>>>(
            features, static_cast<uint16_t*>(logits16), values, row_max, static_cast<const uint8_t*>(weight_blob), param_blob, off,
            nblocks, n, eps, scatter_dst, scatter_desc);
    else
        ya_k_forward<false><<<blocks, kThreads, kSmemBytes, (cudaStream_t)stream>>>(
            features, static_cast<uint16_t*>(logits16), values, row_max, static_cast<const uint8_t*>(weight_blob), param_blob, off,
            nblocks, n, eps, scatter_dst, scatter_desc);
    return (int)cudaGetLastError();
}
