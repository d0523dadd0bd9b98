// Yacht-Auction B200 engine -- the WHOLE leaf-evaluator forward (YachtNNet.forward,
// yacht/pytorch/YachtNNet.py:62-70) as one persistent tcgen05 kernel: state_to_vec features in; tanh values, each
// row's largest logit and the 16-bit policy logits out (scattered straight into the leaves' rows of the tree pool, or
// as a dense matrix padded to 3232 columns).  A CTA owns 128 leaves from the first Linear to the last and two CTAs
// form a pair that shares every weight matrix (tcgen05.mma.cta_group::2, see the kernel); the skip connection
// (float32) and the accumulators live in TMEM, every weight image arrives through cp.async.bulk a stage ahead, and all
// bias / SiLU / LayerNorm / residual / tanh work happens in the tcgen05.ld epilogues (packed FFMA2 / FADD2 arithmetic).
//   input   Linear(59->256) + LN + SiLU                         (1 K-block of 64, N = 256)
//   trunk   nblocks x [LN(SiLU(fc1)), skip + LN(SiLU(fc2))]     (4 K-blocks; two N = 128 halves per layer, each with
//           its own completion barrier: the first epilogue pass over one half runs under the other half's MMAs;
//           activations go back to shared memory as the next 16-bit, 128-byte-swizzled K-major A operand)
//   value   SiLU(LN_v(h)) -> Linear(256->128) + SiLU -> dot(w2) + b2 -> tanh      (N = 128)
//   policy  SiLU(LN_pi(h)) written ONCE to tensor memory (A operand from TMEM) -> 26 tiles of 128 columns:
//           the freed A tile + both weight buffers = six 32 KB half-tile slots, three TMEM accumulators, full /
//           drained mbarriers; warp 15 only produces (copies, MMAs), the other 15 warps run the epilogue
//           (bias, packed running row maximum, 16-bit packing, predicated 256-bit stores)
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
constexpr int kWBytes = kDim * kDim * 2;             // 128 KB per trunk layer in the weight blob ...
constexpr int kWHalf = kWBytes / 2;                  // ... of which each CTA of a pair holds 64 KB (half of the N rows); two buffers
constexpr int kSlotBytes = 32768;                    // one CTA's half of a policy tile (64 of its 128 columns)
constexpr int kSlots = 6;                            // A tile + both weight buffers = 192 KB = six slots in the policy head
constexpr int kPrmFloats = 776;                      // largest parameter block (value head), 16-byte multiple
constexpr int kPiPrmFloats = 2 * kDim + kPolicyTiles * kPolicyTile;            // gamma_pi | beta_pi | bias of all 3,328 columns
constexpr int kBars = 24;
constexpr int kSmemBytes = 1024 + kABytes + 2 * kWHalf + 2 * kPrmFloats * 4 + kPiPrmFloats * 4 + 2 * kRows * kParts * 8 + kBars * 8 + 16;
// mbarriers (per CTA; "leader only" ones are used in rank 0's copy)
enum { B_W = 0,         // [2] weights + parameters of a stage landed in buffer stage & 1 (local bulk copies)
       B_MMA = 2,       // [2] MMAs of the first / second N = 128 half done (commit multicast to both CTAs)
       B_PEER = 4,      // leader only: the other CTA's A tile and weights of this stage are in place
       B_SLOT = 5,      // [6] policy head: this CTA's half of a weight tile landed in the slot
       B_PSLOT = 11,    // [6] leader only: the other CTA's half landed
       B_ACC = 17,      // [3] policy head: accumulator ready (commit multicast)
       B_DRAIN = 20 };  // [3] leader only: accumulator read out by the 15 epilogue warps of BOTH CTAs

struct Blob {                                        // byte / float offsets of the host-built blobs (see mcts.py)
    int64_t w_in, w_trunk, w_v, w_pi;
    int64_t p_in, p_trunk, p_v, p_pi_ln, p_pi_bias;
};

#ifdef YA_FWD_TIMELINE                                 // profiling build only (profiles/tools/forward_timeline.py)
__device__ unsigned long long g_timeline[1024];
#define YA_STAMP() do { if (blockIdx.x == 0 && tid == 0 && tl_n < 1024) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_timeline[tl_n++] = t_; } } while (0)
__device__ unsigned long long g_timeline2[1024];
__device__ unsigned long long g_cta_times[4 * 1024];               // per CTA: start, trunk done, end (globaltimer ns), SM id
#define YA_STAMP2() do { if (blockIdx.x == 0 && tid == 0 && tl2_n < 1024) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_timeline2[tl2_n++] = t_; } } while (0)
#else
#define YA_STAMP() do { } while (0)
#define YA_STAMP2() do { } while (0)
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

// the same from 16 float32 pairs (32 consecutive columns)
template <bool F16>
__device__ __forceinline__ void pack_store_a2(uint8_t* a_tile, int row, int c8_first, const f32x2 (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(a_tile + a_tile_offset(row, c8_first + q)) =
            make_uint4(pack2<F16>(v[4 * q]), pack2<F16>(v[4 * q + 1]), pack2<F16>(v[4 * q + 2]), pack2<F16>(v[4 * q + 3]));
}
__device__ __forceinline__ float hsum2(f32x2 a, f32x2 b) {             // (a.lo + a.hi) + (b.lo + b.hi)
    float a0, a1, b0, b1;
    upk2(a, a0, a1); upk2(b, b0, b1);
    return (a0 + a1) + (b0 + b1);
}
// SiLU of both halves of a pair from t = x / 2: two MUFU.TANH, one packed FMA
__device__ __forceinline__ f32x2 silu2_from_half(f32x2 t) {
    float a, b, ta, tb;
    upk2(t, a, b);
    asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(a));
    asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(b));
#ifdef YA_EXP_TRUNK_NO_TANH
    return t;
#endif
    return fma2(t, pk2(ta, tb), t);
}

// Trunk epilogue, pass 1 (in place): 32 accumulator columns -> x = SiLU(z + b), and the row's running sums of x and x^2
// (four each: columns 0, 1 | 2, 3 mod 4).  half_bias = b / 2.
__device__ __forceinline__ void trunk_pass1(uint32_t (&v)[32], uint32_t taddr, const float* half_bias, f32x2 (&ps)[2], f32x2 (&pq)[2]) {
    tmem_ld32(taddr, v);
    tmem_ld_wait();
    const ulonglong2* bias = reinterpret_cast<const ulonglong2*>(half_bias);
    const f32x2 half2 = pk2(0.5f, 0.5f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const ulonglong2 b = bias[i];
        const f32x2 x0 = silu2_from_half(fma2(pk2u(v[4 * i], v[4 * i + 1]), half2, b.x));
        const f32x2 x1 = silu2_from_half(fma2(pk2u(v[4 * i + 2], v[4 * i + 3]), half2, b.y));
        ps[0] = add2(ps[0], x0); pq[0] = fma2(x0, x0, pq[0]);
        ps[1] = add2(ps[1], x1); pq[1] = fma2(x1, x1, pq[1]);
        upk2u(x0, v[4 * i], v[4 * i + 1]);
        upk2u(x1, v[4 * i + 2], v[4 * i + 3]);
    }
}
// Pass 2 (in place): y = LN(x) (+ the float32 skip connection in tensor memory for fc2: h += LN(SiLU(fc2(..))), the sum is
// the next skip) -> the next layer's A operand in shared memory
template <bool F16, bool SECOND>
__device__ __forceinline__ void trunk_pass2(uint32_t (&v)[32], float rstd, float nm, const float* gamma_p, const float* beta_p,
                                            uint32_t t_skip_addr, uint8_t* a_tile, int row, int c8_first) {
    const ulonglong2* gamma = reinterpret_cast<const ulonglong2*>(gamma_p);
    const ulonglong2* beta = reinterpret_cast<const ulonglong2*>(beta_p);
    const f32x2 rstd2 = pk2(rstd, rstd), nm2 = pk2(nm, nm);
    uint32_t sk[SECOND ? 32 : 1];
    if (SECOND) {
        tmem_ld32(t_skip_addr, reinterpret_cast<uint32_t (&)[32]>(sk));
        tmem_ld_wait();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const ulonglong2 g = gamma[i], b = beta[i];
        f32x2 y0 = fma2(fma2(pk2u(v[4 * i], v[4 * i + 1]), rstd2, nm2), g.x, b.x);
        f32x2 y1 = fma2(fma2(pk2u(v[4 * i + 2], v[4 * i + 3]), rstd2, nm2), g.y, b.y);
        if (SECOND) {
            y0 = add2(y0, pk2u(sk[(4 * i) % (SECOND ? 32 : 1)], sk[(4 * i + 1) % (SECOND ? 32 : 1)]));
            y1 = add2(y1, pk2u(sk[(4 * i + 2) % (SECOND ? 32 : 1)], sk[(4 * i + 3) % (SECOND ? 32 : 1)]));
        }
        upk2u(y0, v[4 * i], v[4 * i + 1]);
        upk2u(y1, v[4 * i + 2], v[4 * i + 3]);
    }
    if (SECOND) tmem_st32(t_skip_addr, v);
    pack_store_a<F16>(a_tile, row, c8_first, v);
}

// F16: operands (activations, weights) and logits in IEEE half -- the precision of the reference's CUDA predict (fp16
// autocast, yacht/NNet.py:186-193) -- instead of bfloat16; accumulation, LayerNorm and the skip connection are float32
// either way.  tcgen05.mma kind::f16 runs both formats at the same rate.
//
// CTA PAIR.  Two CTAs (a cluster of two SMs of one TPC) own 256 leaves, 128 each, and run every matrix product as ONE
// tcgen05.mma.cta_group::2 instruction stream (M = 256) issued by rank 0: each CTA keeps its own 128 activation rows and
// only HALF of every weight matrix (rank r: output columns [64 r, 64 r + 64) of each 128-column block), the tensor cores
// exchange the halves.  Per SM this halves the weight bytes streamed from L2 and the shared-memory reads of the B operand
// (a single-CTA M = 128, N = 128 step reads 8 KB per 64 clocks = the 128 B/clk shared-memory limit; the pair reads 6 KB),
// and it frees 64 KB per CTA: the trunk's weights are double buffered and arrive a whole layer ahead.  Everything after
// the accumulator (epilogues, LayerNorm, skip connection, scatter) is per CTA and unchanged.  Hand-offs: rank 1 tells the
// leader "my A tile and my weight half are in place" with one remote mbarrier arrive per stage; tcgen05.commit multicasts
// completion to both CTAs.
template <bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
ya_k_forward(const float* __restrict__ features, uint16_t* __restrict__ logits, float* __restrict__ values,
             float* __restrict__ row_max, const uint8_t* __restrict__ wblob, const float* __restrict__ pblob, Blob off, int nblocks, int64_t n, float eps,
             const uint64_t* __restrict__ scatter_dst, const uint32_t* __restrict__ scatter_desc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_tile = base;
    uint8_t* w_tiles = base + kABytes;                                // two 64 KB weight buffers
    float* prm_all = reinterpret_cast<float*>(w_tiles + 2 * kWHalf);
    float* pi_prm = prm_all + 2 * kPrmFloats;
    float2* xchg_all = reinterpret_cast<float2*>(pi_prm + kPiPrmFloats);     // two buffers, used alternately
    float2* xchg = xchg_all;
    uint64_t* bars = reinterpret_cast<uint64_t*>(xchg_all + 2 * kParts * kRows);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane;
    const int part = warp >> 2;
    const int64_t grow = (int64_t)blockIdx.x * kRows + row;
    const bool producer = warp == kIssuerWarp;                        // also an epilogue warp, except in the policy head
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    constexpr bool worker = true;

    if (tid == 0) {
        for (int i = 0; i < B_DRAIN; ++i) mbar_init(&bars[i], 1);
        for (int i = B_DRAIN; i < B_DRAIN + 3; ++i) mbar_init(&bars[i], 30);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync();                                                   // both CTAs' barriers exist before anyone arrives remotely
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    // This thread's 64 of a row's 256 columns: 32 from each half, so that the epilogue can start on the first
    // N = 128 half of a trunk layer while the tensor core is still on the second.
    const int colv[2] = {part * 32, 128 + part * 32};
    const uint32_t t_skip = t_lane + 256;                             // float32 skip connection: TMEM columns 256..511
#ifdef YA_FWD_TIMELINE
    if (tid == 0 && blockIdx.x < 1024) {
        unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_));
        g_cta_times[4 * blockIdx.x] = t_; g_cta_times[4 * blockIdx.x + 3] = sm_;
    }
    int tl_n = 0, tl2_n = 0;
    if (blockIdx.x == 0 && tid == 0) {                                // SM clock during the kernel: cycles and ns at both ends
        unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        g_timeline2[1000] = t_; g_timeline2[1001] = (unsigned long long)clock64();
    }
#endif
    const int layers = 2 * nblocks;
    // Stages: 0 = input, 1..layers = trunk, layers + 1 = value head.  Stage s uses weight / parameter buffer s & 1.
    auto w_buf = [&](int s) { return w_tiles + (s & 1) * kWHalf; };
    auto prm_buf = [&](int s) { return prm_all + (s & 1) * kPrmFloats; };
    // this CTA's half of stage s's weight image + the stage's parameter block, on one transaction barrier
    auto load_stage = [&](int s) {
        uint64_t* bar = &bars[B_W + (s & 1)];
        uint8_t* wd = w_buf(s);
        if (s == 0) {
            mbar_expect_tx(bar, 16384 + 3 * kDim * 4);
            bulk_g2s(wd, wblob + off.w_in + rank * 16384, 16384, bar);
            bulk_g2s(prm_buf(s), pblob + off.p_in, 3 * kDim * 4, bar);
        } else if (s <= layers) {
            const uint8_t* src = wblob + off.w_trunk + (int64_t)(s - 1) * kWBytes + rank * kWHalf;
            mbar_expect_tx(bar, kWHalf + 3 * kDim * 4);
            bulk_g2s(wd, src, 32768, bar);
            bulk_g2s(wd + 32768, src + 32768, 32768, bar);
            bulk_g2s(prm_buf(s), pblob + off.p_trunk + (int64_t)(s - 1) * 3 * kDim, 3 * kDim * 4, bar);
        } else {                                                      // value head weights; both heads' parameters
            mbar_expect_tx(bar, 32768 + 772 * 4 + 2 * kDim * 4);
            bulk_g2s(wd, wblob + off.w_v + rank * 32768, 32768, bar);
            bulk_g2s(prm_buf(s), pblob + off.p_v, 772 * 4, bar);
            bulk_g2s(pi_prm, pblob + off.p_pi_ln, 2 * kDim * 4, bar);
        }
    };
    // MMAs of one stage, issued by the leader's elected thread for both CTAs
    auto issue_stage = [&](int s) {
        const uint64_t da = umma_desc(smem_u32(a_tile)), db = umma_desc(smem_u32(w_buf(s)));
        if (s == 0) {                                                 // K = 64, N = 256: each CTA holds 128 weight rows
            const uint32_t idesc = umma_idesc(kDim, F16, 256);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma2(tmem, umma_desc_advance(da, k * 32), umma_desc_advance(db, k * 32), (uint32_t)(k != 0), idesc);
            umma2_commit(&bars[B_MMA]);
            umma2_commit(&bars[B_MMA + 1]);
        } else if (s <= layers) {                                     // two N = 128 halves; per half each CTA holds 64 weight rows
            const uint32_t idesc = umma_idesc(128, F16, 256);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma2(tmem + half * 128, umma_desc_advance(da, kb * (kRows * 128) + k * 32),
                              umma_desc_advance(db, half * 32768 + kb * 8192 + k * 32), (uint32_t)((kb | k) != 0), idesc);
                umma2_commit(&bars[B_MMA + half]);
            }
        } else {                                                      // value head: N = 128 into columns 128..255
            const uint32_t idesc = umma_idesc(128, F16, 256);
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma2(tmem + 128, umma_desc_advance(da, kb * (kRows * 128) + k * 32),
                          umma_desc_advance(db, kb * 8192 + k * 32), (uint32_t)((kb | k) != 0), idesc);
            umma2_commit(&bars[B_MMA]);
            umma2_commit(&bars[B_MMA + 1]);
        }
    };
    // Start of stage s: this CTA's A tile is complete -> (rank 1) tell the leader / (leader) wait for the peer and issue.
    // Then stage s + 1's weights start streaming into the other buffer, whose last reader (stage s - 1) has retired.
    auto begin_stage = [&](int s) {
        proxy_fence();                                                // A tile written through the generic proxy
        tc_fence_before();
        __syncthreads();
        YA_STAMP();                                                   // [3k] epilogue of the previous stage done
        mbar_wait(&bars[B_W + (s & 1)], (uint32_t)((s >> 1) & 1));    // weights + parameters landed
        YA_STAMP();                                                   // [3k+1] weights landed
        if (producer) {
            if (leader) {
                mbar_wait(&bars[B_PEER], (uint32_t)(s & 1));
                tc_fence_after();
                if (elect_one()) issue_stage(s);
            } else {
                if (elect_one()) mbar_arrive_remote(&bars[B_PEER], 0);
            }
            __syncwarp();
            if (s >= 1 && s <= layers && elect_one()) load_stage(s + 1);
            __syncwarp();
        }
    };
    auto wait_mma = [&](int s, int half) {
        mbar_wait(&bars[B_MMA + half], (uint32_t)(s & 1));
        tc_fence_after();
    };
    // the four warps that share a TMEM lane quarter (same 32 rows, different column parts) exchange row statistics
    auto worker_sync = [&] { asm volatile("bar.sync %0, 128;" ::"r"(1 + (warp & 3)) : "memory"); };
    // LayerNorm statistics over the 4 threads of a row.  ONE barrier per exchange: consecutive exchanges alternate between
    // two buffers, and a thread can only reach the exchange after the next (same buffer again) through the barrier of the
    // next one, which every reader of this one has then passed.
    int xb = 0;
    auto row_stats = [&](float s, float ss, float& mean, float& rstd) {
        float2* x = xchg_all + xb * (kParts * kRows);
        xb ^= 1;
        x[part * kRows + row] = make_float2(s, ss);
        worker_sync();
#pragma unroll
        for (int p = 1; p < kParts; ++p) {
            float2 o = x[((part + p) % kParts) * kRows + row];
            s += o.x; ss += o.y;
        }
        mean = s * (1.0f / kDim);
        rstd = rsqrtf(fmaxf(ss * (1.0f / kDim) - mean * mean, 0.0f) + eps);
    };

    // ---------------------------------------------------------------- input stage
    if (producer && elect_one()) {
        load_stage(0);
        load_stage(1);                                                // first trunk layer (or the value head) streams in
    }
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
    begin_stage(0);
    wait_mma(0, 0);
    wait_mma(0, 1);
    YA_STAMP();                                                       // [3k+2] MMA done
    if (worker) {   // h = SiLU(LN(z + b)): Linear -> LayerNorm -> SiLU (YachtNNet.py:25-30); also the first skip connection
        const float* prm = prm_all;
        f32x2 u[2][16];
        f32x2 ps[2] = {0ull, 0ull}, pq[2] = {0ull, 0ull};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(t_lane + colv[c], r);
            tmem_ld_wait();
            const ulonglong2* bias = reinterpret_cast<const ulonglong2*>(prm + colv[c]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const ulonglong2 b = bias[i];
                const f32x2 x0 = add2(pk2u(r[4 * i], r[4 * i + 1]), b.x), x1 = add2(pk2u(r[4 * i + 2], r[4 * i + 3]), b.y);
                ps[0] = add2(ps[0], x0); pq[0] = fma2(x0, x0, pq[0]);
                ps[1] = add2(ps[1], x1); pq[1] = fma2(x1, x1, pq[1]);
                u[c][2 * i] = x0; u[c][2 * i + 1] = x1;
            }
        }
        float mean, rstd;
        row_stats(hsum2(ps[0], ps[1]), hsum2(pq[0], pq[1]), mean, rstd);
        const f32x2 rstd2 = pk2(rstd, rstd), nmean2 = pk2(-mean, -mean), half2 = pk2(0.5f, 0.5f);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const ulonglong2* gamma = reinterpret_cast<const ulonglong2*>(prm + kDim + colv[c]);
            const ulonglong2* beta = reinterpret_cast<const ulonglong2*>(prm + 2 * kDim + colv[c]);
            uint32_t sk[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const ulonglong2 g = gamma[i], b = beta[i];
                const f32x2 ga0 = mul2(rstd2, g.x), ga1 = mul2(rstd2, g.y);
                u[c][2 * i] = silu2_from_half(mul2(half2, fma2(u[c][2 * i], ga0, fma2(nmean2, ga0, b.x))));
                u[c][2 * i + 1] = silu2_from_half(mul2(half2, fma2(u[c][2 * i + 1], ga1, fma2(nmean2, ga1, b.y))));
                upk2u(u[c][2 * i], sk[4 * i], sk[4 * i + 1]);
                upk2u(u[c][2 * i + 1], sk[4 * i + 2], sk[4 * i + 3]);
            }
            tmem_st32(t_skip + colv[c], sk);
            pack_store_a2<F16>(a_tile, row, colv[c] / 8, u[c]);
        }
        tmem_st_wait();
    }

    // ---------------------------------------------------------------- residual trunk
    for (int l = 0; l < layers; ++l) {
        const int stage = l + 1;
        const float* prm = prm_buf(stage);
        const bool second = l & 1;                                    // fc2: add the skip connection
        // Two N = 128 halves, each with its own completion barrier: the epilogue's first pass over columns
        // 0..127 runs while the tensor core works on columns 128..255.
        begin_stage(stage);
        // prm = bias / 2 | gamma | beta (the host halves the bias: SiLU(x) = t + t * tanh(t), t = x / 2).  The 64
        // activations stay in registers across the statistics exchange (no TMEM round trip), as 32 float32 pairs: every
        // arithmetic step is one packed instruction for two neighbouring columns.
        uint32_t v0[32], v1[32];
        f32x2 ps[2] = {0ull, 0ull}, pq[2] = {0ull, 0ull};
        YA_STAMP2();                                                  // [8l] MMAs issued
        wait_mma(stage, 0);
        YA_STAMP2();                                                  // [8l+1] half 0 ready
        trunk_pass1(v0, t_lane + colv[0], prm + colv[0], ps, pq);
        YA_STAMP2();                                                  // [8l+2] pass 1 of half 0 done
        wait_mma(stage, 1);
        YA_STAMP();
        YA_STAMP2();                                                  // [8l+3] half 1 ready
        trunk_pass1(v1, t_lane + colv[1], prm + colv[1], ps, pq);
        float mean, rstd;
        YA_STAMP2();                                                  // [8l+4] pass 1 of half 1 done
        row_stats(hsum2(ps[0], ps[1]), hsum2(pq[0], pq[1]), mean, rstd);
        YA_STAMP2();                                                  // [8l+5] statistics exchanged
        const float nm = -mean * rstd;
        if (second) {
            trunk_pass2<F16, true>(v0, rstd, nm, prm + kDim + colv[0], prm + 2 * kDim + colv[0], t_skip + colv[0], a_tile, row, colv[0] / 8);
            trunk_pass2<F16, true>(v1, rstd, nm, prm + kDim + colv[1], prm + 2 * kDim + colv[1], t_skip + colv[1], a_tile, row, colv[1] / 8);
            tmem_st_wait();
        } else {
            trunk_pass2<F16, false>(v0, rstd, nm, prm + kDim + colv[0], prm + 2 * kDim + colv[0], 0u, a_tile, row, colv[0] / 8);
            trunk_pass2<F16, false>(v1, rstd, nm, prm + kDim + colv[1], prm + 2 * kDim + colv[1], 0u, a_tile, row, colv[1] / 8);
        }
        YA_STAMP2();                                                  // [8l+6] pass 2 done
        YA_STAMP2();
    }

#ifdef YA_FWD_TIMELINE
    if (tid == 0 && blockIdx.x < 1024) {
        unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        g_cta_times[4 * blockIdx.x + 1] = t_;
    }
#endif
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
    // The policy head reads its activations from tensor memory, so all 192 KB of operand space (A tile + both weight
    // buffers) become six 32 KB slots, one per half tile.  Tile j -> slot: 2, 3, 5 (free while the value head still reads
    // the A tile and its weights in slot 4), then 0, 1, 4, and round again.
    auto slot_of = [](int j) { constexpr int m[kSlots] = {2, 3, 5, 0, 1, 4}; return m[j % kSlots]; };
    auto load_tile = [&](int j) {                                     // elected producer thread: this CTA's 64 columns of tile j
        uint64_t* bar = &bars[B_SLOT + slot_of(j)];
        mbar_expect_tx(bar, kSlotBytes + (j == 0 ? kPolicyTiles * kPolicyTile * 4 : 0));
        bulk_g2s(base + slot_of(j) * kSlotBytes, wblob + off.w_pi + (int64_t)j * 65536 + rank * kSlotBytes, kSlotBytes, bar);
        if (j == 0) bulk_g2s(pi_prm + 2 * kDim, pblob + off.p_pi_bias, kPolicyTiles * kPolicyTile * 4, bar);   // every bias
    };
    {
        const int stage = layers + 1;
        const float* prm = prm_buf(stage);                            // gamma_v | beta_v | b1[128] | w2[128] | b2
        mbar_wait(&bars[B_W + (stage & 1)], (uint32_t)((stage >> 1) & 1));   // the LayerNorm parameters travel with the weights
        head_prep(prm, prm + kDim, pi_prm, pi_prm + kDim);
        begin_stage(stage);                                           // accumulator in columns 128..255: 0..127 hold the policy A operand
        if (producer && elect_one())
            for (int j = 0; j < 3; ++j) load_tile(j);
        __syncwarp();
        wait_mma(stage, 0);
        wait_mma(stage, 1);
        YA_STAMP();                                                   // [3k+2] MMA done
        if (producer && elect_one())
            for (int j = 3; j < kSlots; ++j) load_tile(j);            // A tile and value weights are free now
        __syncwarp();
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
    }

    // policy head (YachtNNet.py:38-42): Linear(256, 3226) on the activations already sitting in tensor memory, as 26
    // tiles of 128 columns.  Weight tiles rotate over three 64 KB slots and accumulators over TMEM columns 128 /
    // 256 / 384: as soon as tile j's MMAs retire, tile j + 3's weights start streaming into the slot they read,
    // and the MMAs of the following tiles run under the epilogue of tile j.
    {
        mbar_wait(&bars[B_SLOT + slot_of(0)], 0);                     // tile 0 and every bias landed
        const float* bias_all = pi_prm + 2 * kDim;
        auto issue_tile = [&](int j) {                                // leader's producer thread: tile j's 16 MMAs for both CTAs
            const uint64_t db = umma_desc(smem_u32(base + slot_of(j) * kSlotBytes));
            const uint32_t idesc = umma_idesc(kPolicyTile, F16, 256);
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma2_ts(tmem + kPolicyTile + (j % 3) * kPolicyTile, tmem + kb * 32 + k * 8,
                             umma_desc_advance(db, kb * 8192 + k * 32), (uint32_t)((kb | k) != 0), idesc);
            umma2_commit(&bars[B_ACC + j % 3]);
        };
        proxy_fence();
        tc_fence_before();
        __syncthreads();                                              // the A tile (p) is complete
        tc_fence_after();
        // running row maximum as a packed 16-bit pair: rounding is monotone, so the largest rounded logit is the rounded
        // largest logit -- the value the expand kernel needs (it reads the 16-bit logits)
        constexpr uint32_t kNegInf2 = F16 ? 0xFC00FC00u : 0xFF80FF80u;
        uint32_t row_mx = kNegInf2;
        // Scatter mode (the MCTS path): this row's LEGAL logits go straight into its leaf's row in the tree pool (the row layout
        // ya_mcts_select allocated: csrc/ya_mcts.cu "Logit area"), instead of a dense [n][3232] matrix.  dst = 0: the row needs
        // no evaluation (its descent ended in a terminal / dead-end node).  The policy row is 13 runs of columns -- run 0 = the
        // bid moves [0, 202), run 1 + c = category c, [202 + 252 c, +252) -- and the leaf wants: a bid row run 0 (stored as
        // columns [0, 208), as they are), a ten-dice row the runs of its open categories (one 272-slot block per open category,
        // in order: the 16-aligned column window around the run, so that every 16-column sector keeps its alignment), a
        // five-dice row only the first column of every open category (compact).  `runs` = the wanted runs of the first two
        // kinds; every lane of a warp then walks the same predicated code, whatever mix of leaves the warp holds.
        uint16_t* s_dst = nullptr;
        uint32_t runs = 0, open5 = 0;
        if (scatter_dst && grow < n) {
            s_dst = reinterpret_cast<uint16_t*>(scatter_dst[grow]);
            const uint32_t s_desc = s_dst ? scatter_desc[grow] : 0u;
            const uint32_t open = (s_desc >> 1) & 0xFFFu;             // bit c = category c open
            if (s_desc & 1u) runs = 1u;
            else if (s_desc >> 13) runs = open << 1;
            else open5 = s_desc ? open : 0u;
        }
        if (producer) {
            // Producers (one warp per CTA): keep the tensor pipe fed.  Weight slots cycle with period 6, accumulators (TMEM
            // columns 128 / 256 / 384) with period 3.  Tile j needs both CTAs' halves of its weights (requested five tiles
            // ago; rank 1 reports its half with a remote arrive) and the accumulator drained by both CTAs' epilogues of
            // tile j - 3; once tile j is queued, tile j - 1 has retired and tile j + 5 streams into its slot.
            for (int j = 0; j < kPolicyTiles; ++j) {
                const uint32_t par = (uint32_t)((j / kSlots) & 1);
                mbar_wait(&bars[B_SLOT + slot_of(j)], par);
                if (leader) {
                    mbar_wait(&bars[B_PSLOT + slot_of(j)], par);
                    if (j >= 3) mbar_wait(&bars[B_DRAIN + j % 3], (uint32_t)(((j / 3) - 1) & 1));
                    tc_fence_after();
                    if (elect_one()) issue_tile(j);
                } else {
                    if (elect_one()) mbar_arrive_remote(&bars[B_PSLOT + slot_of(j)], 0);
                }
                __syncwarp();
                if (j >= 1 && j + kSlots - 1 < kPolicyTiles) {
                    mbar_wait(&bars[B_ACC + (j - 1) % 3], (uint32_t)(((j - 1) / 3) & 1));
                    if (elect_one()) load_tile(j + kSlots - 1);
                    __syncwarp();
                }
            }
        } else {
            // 15 epilogue warps.  Rows 96..127 have only three of them (warps 3, 7, 11: the fourth warp of that TMEM lane
            // quarter is the producer), so the quarter's fourth column chunk rotates over the three, one tile each.
            // Per tile a warp pulls its 32 (64) accumulator columns into registers, hands the accumulator back at once,
            // then: bias (the host sets the bias of the padding columns >= 3226 to -inf, so they can never win the
            // maximum), 16-bit packing, packed maximum, stores.
            const bool q3 = (warp & 3) == 3;
            for (int j = 0; j < kPolicyTiles; ++j) {
                const int n_my = (q3 && j % 3 == part) ? 2 : 1;
                mbar_wait(&bars[B_ACC + j % 3], (uint32_t)((j / 3) & 1));
                tc_fence_after();
                YA_STAMP();                                           // policy tile j: accumulator ready
                uint32_t acc[2][32];
#ifdef YA_EXP_POLICY_NO_LD                                 // profiling experiment: the MMA / bulk-copy pipeline alone
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[0][i] = acc[1][i] = 0u;
#else
                tmem_ld32(t_lane + kPolicyTile + (j % 3) * kPolicyTile + part * 32, acc[0]);
                if (n_my == 2) tmem_ld32(t_lane + kPolicyTile + (j % 3) * kPolicyTile + 96, acc[1]);
                tmem_ld_wait();
#endif
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {                                      // this warp's share of the accumulator is in registers
                    if (leader) mbar_arrive(&bars[B_DRAIN + j % 3]);
                    else mbar_arrive_remote(&bars[B_DRAIN + j % 3], 0);
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (q >= n_my) break;
                    const uint32_t (&r)[32] = acc[q];
                    const int col0 = j * kPolicyTile + (q ? 96 : part * 32);
#if defined(YA_EXP_POLICY_NO_LD) || defined(YA_EXP_POLICY_NO_ST)
                    if (col0 < kPolicyCols && r[0] == 0x7FC12345u) {   // profiling experiment: epilogue without its stores
#else
                    if (col0 < kPolicyCols) {
#endif
                        const ulonglong2* bias = reinterpret_cast<const ulonglong2*>(bias_all + col0);
                        uint32_t pk[16];
                        uint32_t m0 = row_mx, m1 = kNegInf2;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const ulonglong2 b = bias[i];
                            pk[2 * i] = pack2<F16>(add2(pk2u(r[4 * i], r[4 * i + 1]), b.x));
                            pk[2 * i + 1] = pack2<F16>(add2(pk2u(r[4 * i + 2], r[4 * i + 3]), b.y));
                            m0 = max16x2<F16>(m0, pk[2 * i]);
                            m1 = max16x2<F16>(m1, pk[2 * i + 1]);
                        }
                        row_mx = max16x2<F16>(m0, m1);
                        auto store_sector = [&](uint16_t* dst, int h) {    // 16 columns = one aligned 32-byte store
                            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst),
                                         "r"(pk[8 * h]), "r"(pk[8 * h + 1]), "r"(pk[8 * h + 2]), "r"(pk[8 * h + 3]),
                                         "r"(pk[8 * h + 4]), "r"(pk[8 * h + 5]), "r"(pk[8 * h + 6]), "r"(pk[8 * h + 7]) : "memory");
                        };
                        if (scatter_dst) {
                            // Bid and ten-dice rows: each of the chunk's two 16-column sectors lies in one or two runs (a
                            // sector that holds a run boundary may be wanted by both neighbours); which runs is the same for
                            // every lane, whether the lane's leaf wants them is one bit of `runs`.
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int cs = col0 + 16 * h;
                                const int ra = ((cs + 50) * 4162) >> 20, rb = ((cs + 65) * 4162) >> 20;   // runs of the sector's first / last column
#pragma unroll
                                for (int t = 0; t < 2; ++t) {
                                    const int rr = t ? rb : ra;
                                    if ((t == 1 && rb == ra) || rr > 12) continue;
                                    if ((runs >> rr) & 1u) {
                                        const int start = rr ? 202 + 252 * (rr - 1) : 0;
                                        const int before = __popc(runs & ((1u << rr) - 1u));   // wanted runs in front of this one
                                        store_sector(s_dst + 272 * before + (cs - (start & ~15)), h);
                                    }
                                }
                            }
                            // Five-dice rows: subset 0 of every open category, i.e. the first column of a run, if one starts
                            // inside this chunk (run starts are even columns: the low half of a packed pair).
                            const int r0 = ((col0 + 50) * 4162) >> 20;
                            const int end0 = 202 + 252 * r0;                               // category r0 starts here
                            if (end0 < col0 + 32 && r0 < 12 && ((open5 >> r0) & 1u)) {
                                uint16_t* one = s_dst + __popc(open5 & ((1u << r0) - 1u));
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if (col0 + 2 * i == end0) *one = (uint16_t)(pk[i] & 0xFFFFu);
                            }
                        }
                        if (logits && grow < n) {
                            uint16_t* dst = logits + grow * kPolicyCols + col0;         // 64 bytes, 32-byte aligned
                            store_sector(dst, 0);
                            store_sector(dst + 16, 1);
                        }
                    }
                }
            }
        }
        xchg[part * kRows + row].x = hmax16x2<F16>(row_mx);           // (the producer warp's entry stays -inf)
        YA_STAMP();
        // the row's largest logit as the expand kernel will see it (rounding to 16 bits is monotone)
        __syncthreads();
        if (part == 0 && grow < n && row_max)
            row_max[grow] = fmaxf(fmaxf(xchg[0 * kRows + row].x, xchg[1 * kRows + row].x), fmaxf(xchg[2 * kRows + row].x, xchg[3 * kRows + row].x));
    }
#ifdef YA_FWD_TIMELINE
    if (tid == 0 && blockIdx.x < 1024) {
        unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        g_cta_times[4 * blockIdx.x + 2] = t_;
    }
    if (blockIdx.x == 0 && tid == 0) {
        unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        g_timeline2[1002] = t_; g_timeline2[1003] = (unsigned long long)clock64();
    }
#endif
    tc_fence_before();
    cluster_sync();                                                   // neither CTA leaves while the other may still signal it
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace

#ifdef YA_FWD_TIMELINE
extern "C" int ya_debug_forward_timeline(unsigned long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(unsigned long long) * 1024);
}
extern "C" int ya_debug_forward_cta_times(unsigned long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_cta_times, sizeof(unsigned long long) * 4 * 1024);
}
extern "C" int ya_debug_forward_timeline2(unsigned long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_timeline2, sizeof(unsigned long long) * 1024);
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
    int blocks = 2 * (int)((n + 2 * kRows - 1) / (2 * kRows));       // CTA pairs: 256 leaves per cluster
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
