// Yacht-Auction B200 engine -- the WHOLE leaf-evaluator forward (YachtNNet.forward,
// yacht/pytorch/YachtNNet.py:62-70) as one persistent tcgen05 kernel: state_to_vec features in, bf16 policy
// logits (padded to 3232 columns) and tanh values out.  A CTA owns 128 leaves from the first Linear to the
// last: activations stay in shared memory (bf16, 128-byte-swizzled K-major A operand), the skip connection
// and the accumulators live in TMEM, every weight matrix arrives as a pre-swizzled image through
// cp.async.bulk while the previous stage's epilogue runs, and all bias / SiLU / LayerNorm / residual / tanh
// work happens in the tcgen05.ld epilogues.  Stages:
//   input   Linear(59->256) + LN + SiLU                         (1 K-block of 64, N = 256)
//   trunk   nblocks x [LN(SiLU(fc1)), skip + LN(SiLU(fc2))]     (4 K-blocks, N = 256)
//   value   SiLU(LN_v(h)) -> Linear(256->128) + SiLU -> dot(w2) + b2 -> tanh      (N = 128)
//   policy  SiLU(LN_pi(h)) -> 26 tiles of Linear(256->128 columns) + bias -> bf16 logits
// Every row is computed independently of the batch it sits in (fixed tile shapes, fixed accumulation order),
// so the evaluator is batch-invariant: sharding leaves over GPUs or waves cannot change a single bit.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include "../../include/yacht_b200.h"
#include "ya_tc.cuh"

namespace {

using namespace ya_tc;

constexpr int kRows = 128, kDim = 256, kParts = 4;
constexpr int kWorkers = 512;                       // 16 epilogue warps: warp w owns TMEM lanes 32 (w % 4).., column part w / 4
constexpr int kProducer = kWorkers;                  // lane 0 of a 17th warp issues every bulk copy and every MMA
constexpr int kThreads = kWorkers + 32;
constexpr int kFeat = 59;
constexpr int kPolicyCols = 3232, kPolicyTile = 128, kPolicyTiles = 26;        // 26 * 128 = 3328 >= 3232
constexpr int kABytes = kRows * kDim * 2;            // 64 KB
constexpr int kWBytes = kDim * kDim * 2;             // 128 KB (two 64 KB halves for the N = 128 stages)
constexpr int kPrmFloats = 776;                      // largest parameter block (value head), 16-byte multiple
constexpr int kPiPrmFloats = 2 * kDim + kPolicyTiles * kPolicyTile;            // gamma_pi | beta_pi | bias of all 3,328 columns
constexpr int kActions = 3226;
#ifndef YA_FWD_BULK_CHUNK
#define YA_FWD_BULK_CHUNK 32768
#endif
constexpr int kBulkChunk = YA_FWD_BULK_CHUNK;        // bytes per cp.async.bulk request
constexpr int kSmemBytes = 1024 + kABytes + kWBytes + 2 * kPrmFloats * 4 + kPiPrmFloats * 4 + kRows * kParts * 8 + 64;

struct Blob {                                        // byte / float offsets of the host-built blobs (see mcts.py)
    int64_t w_in, w_trunk, w_v, w_pi;
    int64_t p_in, p_trunk, p_v, p_pi_ln, p_pi_bias;
};

#ifdef YA_FWD_TIMELINE                                 // profiling build only (profiles/tools/forward_timeline.py)
__device__ unsigned long long g_timeline[1024];
#define YA_STAMP() do { if (blockIdx.x == 0 && tid == 0 && tl_n < 1024) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_timeline[tl_n++] = t_; } } while (0)
#define YA_STAMP2() do { if (blockIdx.x == 0 && tl_n < 512) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_timeline[512 + tl_n++] = t_; } } while (0)
#else
#define YA_STAMP() do { } while (0)
#define YA_STAMP2() do { } while (0)
#endif

__device__ __forceinline__ uint32_t a_tile_offset(int r, int c8) {
    int kb = c8 >> 3, chunk = c8 & 7;
    return (uint32_t)(kb * (kRows * 128) + r * 128 + ((chunk ^ (r & 7)) << 4));
}

__device__ __forceinline__ void pack_store_a(uint8_t* a_tile, int row, int c8_first, const uint32_t (&r)[32]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t p[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[q * 8 + 2 * i]), __uint_as_float(r[q * 8 + 2 * i + 1]));
            p[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(a_tile + a_tile_offset(row, c8_first + q)) = make_uint4(p[0], p[1], p[2], p[3]);
    }
}

__global__ void __launch_bounds__(kThreads, 1)
ya_k_forward(const float* __restrict__ features, __nv_bfloat16* __restrict__ logits, float* __restrict__ values,
             float* __restrict__ row_max, const uint8_t* __restrict__ wblob, const float* __restrict__ pblob, Blob off, int nblocks, int64_t n, float eps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_tile = base;
    uint8_t* w_tile = base + kABytes;
    float* prm_all = reinterpret_cast<float*>(w_tile + kWBytes);
    float* pi_prm = prm_all + 2 * kPrmFloats;
    float2* xchg = reinterpret_cast<float2*>(pi_prm + kPiPrmFloats);
    uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + kParts * kRows);     // [0] weights landed, [1] / [2] MMA done,
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);             // [3] / [4] policy weight halves landed,
                                                                             // [5] / [6] policy accumulator drained

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane;
    const int part = warp >> 2;
    const int64_t grow = (int64_t)blockIdx.x * kRows + row;
    const bool worker = warp < kWorkers / 32;                         // the producer warp only keeps the barriers company

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        mbar_init(&bars[3], 1);
        mbar_init(&bars[4], 1);
        mbar_init(&bars[5], kWorkers / 32);
        mbar_init(&bars[6], kWorkers / 32);
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
    const uint32_t t_acc64 = t_lane + (uint32_t)(part * 64);          // my 64 columns of an N = 256 accumulator
    const uint32_t t_skip = t_lane + 256 + (uint32_t)(part * 64);     // my 64 columns of the float32 skip connection
    uint32_t w_phase = 0, m_phase = 0;
    int tl_n = 0; (void)tl_n;
    int stage = 0;                                                    // parameter double buffer index = stage & 1

    // one weight image + one parameter block per stage, on one transaction barrier
    auto load_stage = [&](uint8_t* w_dst, int64_t w_src, uint32_t w_bytes, float* p_dst, int64_t p_src, uint32_t p_floats) {
        mbar_expect_tx(&bars[0], w_bytes + p_floats * 4);
        for (uint32_t o = 0; o < w_bytes; o += 32768) bulk_g2s(w_dst + o, wblob + w_src + o, min(32768u, w_bytes - o), &bars[0]);
        if (p_floats) bulk_g2s(p_dst, pblob + p_src, p_floats * 4, &bars[0]);
    };
    auto prm_buf = [&](int buf) { return prm_all + (buf & 1) * kPrmFloats; };
    // MMA of one stage: A tile (n_kb K-blocks of 64) x weight image at w_src (rows = n_cols) -> TMEM column d_col
    auto run_mma = [&](const uint8_t* w_src, int n_kb, int n_cols, uint32_t d_col) {
        proxy_fence();                                                // A tile written through the generic proxy
        tc_fence_before();
        __syncthreads();
        YA_STAMP();                                                   // [3k] epilogue of the previous stage done
        mbar_wait(&bars[0], w_phase);                                 // weights + parameters landed
        w_phase ^= 1;
        YA_STAMP();                                                   // [3k+1] weights landed
        if (tid == kProducer) {
            tc_fence_after();
            const uint32_t a0 = smem_u32(a_tile), b0 = smem_u32(w_src);
            const uint32_t idesc = umma_idesc(n_cols);
            for (int kb = 0; kb < n_kb; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma(tmem + d_col, umma_desc(a0 + kb * (kRows * 128) + k * 32), umma_desc(b0 + kb * (n_cols * 128) + k * 32),
                         (uint32_t)((kb | k) != 0), idesc);
            umma_commit(&bars[1]);
        }
        mbar_wait(&bars[1], m_phase);                                 // accumulator ready; A tile and this weight buffer free
        m_phase ^= 1;
        tc_fence_after();
        YA_STAMP();                                                   // [3k+2] MMA done
    };
    auto worker_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory"); };   // the 16 epilogue warps only
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
    if (tid == kProducer) load_stage(w_tile, off.w_in, 256 * 128, prm_buf(0), off.p_in, 3 * kDim);
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
            for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(f[q * 8 + 2 * i]), __uint_as_float(f[q * 8 + 2 * i + 1]));
                p[i] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(a_tile + a_tile_offset(row, part * 2 + q)) = make_uint4(p[0], p[1], p[2], p[3]);
        }
    }
    run_mma(w_tile, 1, kDim, 0);
    if (tid == kProducer) load_stage(w_tile, off.w_trunk, kWBytes, prm_buf(1), off.p_trunk, 3 * kDim);      // first trunk layer streams in
    if (worker) {   // h = SiLU(LN(z + b)): Linear -> LayerNorm -> SiLU (YachtNNet.py:25-30); also the first skip connection
        const float* prm = prm_all;
        float ps[4] = {0, 0, 0, 0}, pq[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(t_acc64 + c * 32, r);
            tmem_ld_wait();
            const float* bias = prm + part * 64 + c * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float v = __uint_as_float(r[i]) + bias[i];
                ps[i & 3] += v; pq[i & 3] = fmaf(v, v, pq[i & 3]);
                r[i] = __float_as_uint(v);
            }
            tmem_st32(t_acc64 + c * 32, r);
        }
        tmem_st_wait();
        float mean, rstd;
        row_stats((ps[0] + ps[1]) + (ps[2] + ps[3]), (pq[0] + pq[1]) + (pq[2] + pq[3]), mean, rstd);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(t_acc64 + c * 32, r);
            tmem_ld_wait();
            const float* gamma = prm + kDim + part * 64 + c * 32;
            const float* beta = prm + 2 * kDim + part * 64 + c * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float ga = rstd * gamma[i];
                float y = fmaf(__uint_as_float(r[i]), ga, fmaf(-mean, ga, beta[i]));
                r[i] = __float_as_uint(silu_from_half(0.5f * y));
            }
            tmem_st32(t_skip + c * 32, r);
            pack_store_a(a_tile, row, part * 8 + c * 4, r);
        }
        tmem_st_wait();
    }
    stage = 1;

    // ---------------------------------------------------------------- residual trunk
    const int layers = 2 * nblocks;
    for (int l = 0; l < layers; ++l, ++stage) {
        const float* prm = prm_all + (stage & 1) * kPrmFloats;
        const bool second = l & 1;                                    // fc2: add the skip connection
        run_mma(w_tile, 4, kDim, 0);
        if (tid == kProducer) {                                       // next stage's weights under this epilogue
            if (l + 1 < layers) load_stage(w_tile, off.w_trunk + (int64_t)(l + 1) * kWBytes, kWBytes, prm_buf(stage + 1),
                                           off.p_trunk + (int64_t)(l + 1) * 3 * kDim, 3 * kDim);
            else load_stage(w_tile, off.w_v, 128 * 512, prm_buf(stage + 1), off.p_v, 772);
        }
        if (!worker) continue;
        float ps[4] = {0, 0, 0, 0}, pq[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(t_acc64 + c * 32, r);
            tmem_ld_wait();
            const float* bias = prm + part * 64 + c * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float v = silu_from_half(fmaf(__uint_as_float(r[i]), 0.5f, 0.5f * bias[i]));
                ps[i & 3] += v; pq[i & 3] = fmaf(v, v, pq[i & 3]);
                r[i] = __float_as_uint(v);
            }
            tmem_st32(t_acc64 + c * 32, r);
        }
        tmem_st_wait();
        float mean, rstd;
        row_stats((ps[0] + ps[1]) + (ps[2] + ps[3]), (pq[0] + pq[1]) + (pq[2] + pq[3]), mean, rstd);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32], sk[32];
            tmem_ld32(t_acc64 + c * 32, r);
            if (second) tmem_ld32(t_skip + c * 32, sk);
            tmem_ld_wait();
            const float* gamma = prm + kDim + part * 64 + c * 32;
            const float* beta = prm + 2 * kDim + part * 64 + c * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float ga = rstd * gamma[i];
                float v = fmaf(__uint_as_float(r[i]), ga, fmaf(-mean, ga, beta[i]));
                if (second) { v += __uint_as_float(sk[i]); }
                r[i] = __float_as_uint(v);
            }
            if (second) tmem_st32(t_skip + c * 32, r);
            pack_store_a(a_tile, row, part * 8 + c * 4, r);
        }
        if (second) tmem_st_wait();
    }

    // ---------------------------------------------------------------- heads: a = SiLU(LN(h; gamma, beta)) from the skip
    // to_tmem: the activations go to tensor-memory columns [0, 128) as packed bf16 pairs (A operand read from TMEM)
    auto head_prep = [&](const float* gamma_all, const float* beta_all, bool to_tmem) {
        if (!worker) return;
        float ps[4] = {0, 0, 0, 0}, pq[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(t_skip + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) { float v = __uint_as_float(r[i]); ps[i & 3] += v; pq[i & 3] = fmaf(v, v, pq[i & 3]); }
        }
        float mean, rstd;
        row_stats((ps[0] + ps[1]) + (ps[2] + ps[3]), (pq[0] + pq[1]) + (pq[2] + pq[3]), mean, rstd);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(t_skip + c * 32, r);
            tmem_ld_wait();
            const float* gamma = gamma_all + part * 64 + c * 32;
            const float* beta = beta_all + part * 64 + c * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float ga = rstd * gamma[i];
                float y = fmaf(__uint_as_float(r[i]), ga, fmaf(-mean, ga, beta[i]));
                r[i] = __float_as_uint(silu_from_half(0.5f * y));
            }
            if (to_tmem) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
                    pk[i] = *reinterpret_cast<uint32_t*>(&h);
                }
                tmem_st16(t_lane + (uint32_t)(part * 32 + c * 16), pk);
            } else {
                pack_store_a(a_tile, row, part * 8 + c * 4, r);
            }
        }
        if (to_tmem) tmem_st_wait();
    };

    // value head (YachtNNet.py:44-50): LN -> SiLU -> Linear(256,128) -> SiLU -> Linear(128,1) -> tanh
    {
        const float* prm = prm_all + (stage & 1) * kPrmFloats;       // gamma_v | beta_v | b1[128] | w2[128] | b2
        mbar_wait(&bars[0], w_phase);                                 // the LayerNorm parameters travel with the weights
        head_prep(prm, prm + kDim, false);
        run_mma(w_tile, 4, 128, 0);                                   // (re-waits the same completed phase, then flips it)
        if (tid == kProducer) load_stage(w_tile + 65536, off.w_pi, 65536, pi_prm, off.p_pi_ln, kPiPrmFloats);
        float dot = 0.0f;
        if (worker) {
            uint32_t r[32];
            tmem_ld32(t_lane + part * 32, r);
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

    // policy head (YachtNNet.py:38-42): LN -> SiLU -> Linear(256, 3226) as 26 tiles of 128 columns.  Weight tiles
    // ping-pong between the two 64 KB halves of the weight region (one transaction barrier per half) and
    // accumulators between TMEM columns 0 and 128: as soon as tile j's MMAs retire, tile j + 2 starts streaming
    // into the half they read and tile j + 1's MMAs are issued, all under the epilogue of tile j.
    {
        mbar_wait(&bars[0], w_phase);                                 // tile 0, gamma_pi | beta_pi and every bias landed
        w_phase ^= 1;
        head_prep(pi_prm, pi_prm + kDim, true);
        const float* bias_all = pi_prm + 2 * kDim;
        auto tile_half = [&](int j) { return w_tile + ((j + 1) & 1) * 65536; };     // tile 0 sits in the upper half
        auto load_tile = [&](int j) {                                 // producer thread, j >= 1
            uint64_t* bar = &bars[3 + (j & 1)];
#ifdef YA_FWD_EXP_NOLOAD
            mbar_expect_tx(bar, 1024);
            bulk_g2s(tile_half(j), wblob + off.w_pi + (int64_t)j * 65536, 1024, bar);
#else
            mbar_expect_tx(bar, 65536);
#pragma unroll
            for (int o = 0; o < 65536; o += kBulkChunk) bulk_g2s(tile_half(j) + o, wblob + off.w_pi + (int64_t)j * 65536 + o, kBulkChunk, bar);
#endif
        };
        auto issue_tile = [&](int j) {                                // producer thread: tile j's 16 MMAs
            const uint32_t b0 = smem_u32(tile_half(j));
            const uint32_t idesc = umma_idesc(kPolicyTile);
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_ts(tmem + kPolicyTile + (j & 1) * kPolicyTile, tmem + kb * 32 + k * 8,
                            umma_desc(b0 + kb * (kPolicyTile * 128) + k * 32), (uint32_t)((kb | k) != 0), idesc);
            umma_commit(&bars[1 + (j & 1)]);
        };
        proxy_fence();
        tc_fence_before();
        __syncthreads();                                              // the A tile (p) is complete
        tc_fence_after();
        const uint32_t m1_base = m_phase;                             // parity of bars[1] for policy tile 0
        auto mma_parity = [&](int j) { return (j & 1) ? (uint32_t)((j >> 1) & 1) : (m1_base ^ (uint32_t)((j >> 1) & 1)); };
        float row_mx = -3.0e38f;
        if (tid == kProducer) {
            // Producer: keeps the tensor pipe fed.  Tile j needs its weights (landed), and its accumulator half
            // drained by the epilogue of tile j - 2; tile j + 1's weights go where tile j - 1's were as soon as
            // tile j - 1's MMAs have retired -- by then tile j's MMAs are already queued behind them.
            load_tile(1);                                             // lower half: the value head is done with it
            for (int j = 0; j < kPolicyTiles; ++j) {
                YA_STAMP2();                                          // [5j] loop top
                if (j >= 1) mbar_wait(&bars[3 + (j & 1)], (uint32_t)(((j - 1) >> 1) & 1));
                YA_STAMP2();                                          // [5j+1] weights landed
                if (j >= 2) mbar_wait(&bars[5 + (j & 1)], (uint32_t)(((j - 2) >> 1) & 1));
                YA_STAMP2();                                          // [5j+2] accumulator drained
                tc_fence_after();
                issue_tile(j);
                YA_STAMP2();                                          // [5j+3] MMAs issued
                if (j >= 1 && j + 1 < kPolicyTiles) {
                    mbar_wait(&bars[1 + ((j - 1) & 1)], mma_parity(j - 1));
                    load_tile(j + 1);
                }
                YA_STAMP2();                                          // [5j+4] previous tile retired, next load issued
            }
        } else if (worker) {
            for (int j = 0; j < kPolicyTiles; ++j) {
                mbar_wait(&bars[1 + (j & 1)], mma_parity(j));
                tc_fence_after();
                YA_STAMP();                                           // policy tile j: accumulator ready
                uint32_t r[32];
#ifdef YA_FWD_EXP_NOTMEMLD
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = (uint32_t)(j + i);
#else
                tmem_ld32(t_lane + kPolicyTile + (j & 1) * kPolicyTile + part * 32, r);
                tmem_ld_wait();
#endif
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[5 + (j & 1)]);       // this warp's share of the accumulator is in registers
                const int col0 = j * kPolicyTile + part * 32;
                if (col0 < kPolicyCols) {
                    const float* bias = bias_all + col0;
                    float f[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]) + bias[i];
                    if (col0 + 32 <= kActions) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) row_mx = fmaxf(row_mx, f[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (col0 + i < kActions) row_mx = fmaxf(row_mx, f[i]);
                    }
                    if (grow < n) {
                        uint32_t p[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                            p[i] = *reinterpret_cast<uint32_t*>(&h);
                        }
                        __nv_bfloat16* dst = logits + grow * kPolicyCols + col0;    // 64 bytes, 32-byte aligned
#ifndef YA_FWD_EXP_NOSTORE
#pragma unroll
                        for (int q = 0; q < 2; ++q)                   // 256-bit stores: half the LSU work of 4 x 16 bytes
                            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 16 * q),
                                         "r"(p[8 * q]), "r"(p[8 * q + 1]), "r"(p[8 * q + 2]), "r"(p[8 * q + 3]), "r"(p[8 * q + 4]),
                                         "r"(p[8 * q + 5]), "r"(p[8 * q + 6]), "r"(p[8 * q + 7]) : "memory");
#else
                        if (p[0] == 0x12345678u && p[5] == 0x9abcdef0u) *reinterpret_cast<uint4*>(dst) = make_uint4(p[1] ^ p[9], p[2] ^ p[10], p[3] ^ p[11], p[4] ^ p[15]);
#endif
                    }
                }
            }
        }
        YA_STAMP();
        // the row's largest logit as the expand kernel will see it (bf16 rounding is monotone)
        if (worker) xchg[part * kRows + row].x = row_mx;
        __syncthreads();
        if (worker && part == 0 && grow < n && row_max) {
            float m = fmaxf(fmaxf(row_mx, xchg[1 * kRows + row].x), fmaxf(xchg[2 * kRows + row].x, xchg[3 * kRows + row].x));
            row_max[grow] = __bfloat162float(__float2bfloat16_rn(m));
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

extern "C" int ya_nn_forward(const float* features, void* logits_bf16, float* values, float* row_max, const void* weight_blob,
                             const float* param_blob, const int64_t* offsets, int nblocks, int64_t n, float eps, void* stream) {
    if (n <= 0) return 0;
    if ((reinterpret_cast<uintptr_t>(logits_bf16) | reinterpret_cast<uintptr_t>(weight_blob) |
         reinterpret_cast<uintptr_t>(param_blob)) & 31u) return (int)cudaErrorMisalignedAddress;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(ya_k_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    Blob off{offsets[0], offsets[1], offsets[2], offsets[3], offsets[4], offsets[5], offsets[6], offsets[7], offsets[8]};
    int blocks = (int)((n + kRows - 1) / kRows);
    ya_k_forward<<<blocks, kThreads, kSmemBytes, (cudaStream_t)stream>>>(
        features, static_cast<__nv_bfloat16*>(logits_bf16), values, row_max, static_cast<const uint8_t*>(weight_blob), param_blob, off,
        nblocks, n, eps);
    return (int)cudaGetLastError();
}
