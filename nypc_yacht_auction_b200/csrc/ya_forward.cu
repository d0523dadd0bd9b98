// Yacht-Auction B200 engine -- the WHOLE leaf-evaluator forward (YachtNNet.forward,
// yacht/pytorch/YachtNNet.py:62-70) as one persistent tcgen05 kernel: state_to_vec features in; tanh values, each
// row's largest logit and the 16-bit policy logits out (scattered straight into the leaves' rows of the tree pool, or
// as a dense matrix padded to 3232 columns).  A CTA owns 128 (ya_k_forward) or 2 x 128 (ya_k_forward2) leaves from the
// first Linear to the last, and two CTAs form a pair that shares every weight matrix (tcgen05.mma.cta_group::2, see
// "CTA pair" below); accumulators and the residual stream live in TMEM, every weight image arrives through
// cp.async.bulk ahead of its use, and all bias / SiLU / LayerNorm / residual / tanh work happens in the tcgen05.ld
// epilogues (packed FFMA2 / FADD2 arithmetic).
//   input   Linear(59->256) + LN + SiLU                         (1 K-block of 64, N = 256)
//   trunk   nblocks x [LN(SiLU(fc1)), skip + LN(SiLU(fc2))]     (4 K-blocks; two N = 128 halves per layer, each with
//           its own completion barrier: the first epilogue pass over one half runs under the other half's MMAs;
//           activations go back to shared memory as the next 16-bit, 128-byte-swizzled K-major A operand)
//   value   SiLU(LN_v(h)) -> Linear(256->128) + SiLU -> dot(w2) + b2 -> tanh      (N = 128)
//   policy  SiLU(LN_pi(h)) written ONCE to tensor memory (A operand from TMEM) -> 26 tiles of 128 columns through six
//           32 KB half-tile slots (all of the operand space), TMEM accumulators with full / drained mbarriers; warp 15
//           only produces (copies, MMAs), the other 15 warps run the epilogue (bias, packed running row maximum, 16-bit
//           packing, predicated 256-bit stores)
// ARITHMETIC (the same in both kernels, so a row's outputs do not depend on which one ran, on the batch around it, on
// the wave or on the GPU count): operands (activations, weights) and logits are 16-bit -- IEEE half by default, the
// precision of the reference's CUDA predict (fp16 autocast, yacht/NNet.py:186-193), or bfloat16; accumulation, bias,
// SiLU, LayerNorm and the residual sum are float32; the residual stream h is STORED between blocks as IEEE half (in both
// operand modes): with half operands the stored value is exactly the next block's fc1 operand, so nothing is rounded that
// the next matrix product would not round anyway.  Measured against the float32 module (4,096 random rows): logit rms
// error 1.20e-3 (1.06e-3 with a float32 residual stream); the reference's own autocast arithmetic, which rounds every
// Linear output and every SiLU to half, emulated step by step: 1.50e-3 ("Stated tolerance").
// CTA PAIR.  Two CTAs (a cluster of two SMs of one TPC) run every matrix product as ONE tcgen05.mma.cta_group::2
// instruction stream (M = 256) issued by rank 0: each CTA keeps its own activation rows and only HALF of every weight
// matrix (rank r: output columns [64 r, 64 r + 64) of each 128-column block), the tensor cores exchange the halves.  Per
// SM this halves the weight bytes streamed from L2 and the shared-memory reads of the B operand (a single-CTA M = 128,
// N = 128 step reads 8 KB per 64 clocks = the 128 B/clk shared-memory limit; the pair reads 6 KB).  Hand-offs: rank 1
// tells the leader "my operand tile, my weight half and my accumulator are ready" with one remote mbarrier arrive per
// MMA group; tcgen05.commit multicasts completion to both CTAs.
// TWO TILES PER CTA (ya_k_forward2, waves of more than 148 x 128 leaves).  One 128-row tile cannot keep an SM busy: a
// layer is 1.2 us of tensor work followed by 2.6 us of epilogue (MUFU-bound SiLU: 16 tanh per clock and SM) that the
// next layer's MMAs must wait for.  With two tiles X and Y per CTA the same 16 warps alternate E(X, l), E(Y, l),
// E(X, l + 1) ... and the MMAs of the tile they are NOT working on run underneath: the tensor pipe order is X.h0 X.h1
// Y.h0 Y.h1 of layer l, then layer l + 1, one accumulator (256 columns) is shared by both tiles -- a half is re-issued as
// soon as every warp has pulled the previous tile's half into registers -- and the two half-precision residual streams
// take 128 columns each.  Shared memory: two A tiles + ONE 64 KB weight buffer whose halves are refilled as soon as
// tile Y's MMAs on them retire.  The policy head does both tiles per weight slot.
// Single-thread instructions (MMA, commit, bulk copy) are issued under elect.sync so they compile to straight-line
// SASS.  -DYA_FWD_TIMELINE builds the profiling variant used by profiles/tools/forward_timeline.py.
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
constexpr int kWHalf = kWBytes / 2;                  // ... of which each CTA of a pair holds 64 KB (half of the N rows)
constexpr int kSlotBytes = 32768;                    // one CTA's half of a policy tile (64 of its 128 columns) / of a trunk N half
constexpr int kSlots = 6;                            // operand space (A tiles + weight buffers) = 192 KB = six slots in the policy head
constexpr int kPrmFloats = 776;                      // largest parameter block (value head), 16-byte multiple
constexpr int kPiPrmFloats = 2 * kDim + kPolicyTiles * kPolicyTile;            // gamma_pi | beta_pi | bias of all 3,328 columns
constexpr int kBars = 32;
constexpr int kSmemBytes = 1024 + 3 * kABytes + 2 * kPrmFloats * 4 + kPiPrmFloats * 4 + 2 * kRows * kParts * 8 + kBars * 8 + 16;
constexpr int kSkipCol = 256;                        // TMEM: residual stream of tile t in columns 256 + 128 t (packed half pairs)

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
#define YA_STAMP2() do { if (blockIdx.x == 0 && tid == 0 && tl2_n < 1000) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_timeline2[tl2_n++] = t_; } } while (0)
#define YA_CTA_TIME(k) do { if (tid == 0 && blockIdx.x < 1024) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_cta_times[4 * blockIdx.x + (k)] = t_; } } while (0)
#else
#define YA_STAMP() do { } while (0)
#define YA_STAMP2() do { } while (0)
#define YA_CTA_TIME(k) do { } while (0)
#endif

__device__ __forceinline__ uint32_t a_tile_offset(int r, int c8) {
    int kb = c8 >> 3, chunk = c8 & 7;
    return (uint32_t)(kb * (kRows * 128) + r * 128 + ((chunk ^ (r & 7)) << 4));
}
// 32 consecutive columns of one row, already packed to 16 bits, into the swizzled A tile
__device__ __forceinline__ void store_a_packed(uint8_t* a_tile, int row, int c8_first, const uint32_t (&pk)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(a_tile + a_tile_offset(row, c8_first + q)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
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
    return fma2(t, pk2(ta, tb), t);
}

// Per-thread geometry and the LayerNorm statistics exchange.  Thread = one row (TMEM lane) x 64 of its 256 columns:
// 32 from each N = 128 half, so that an epilogue can start on the first half while the tensor core is on the second.
struct Lane {
    int row, part, warp;
    int col[2];
    uint32_t t_lane;                                  // TMEM address of this thread's lane, column 0
    float2* xchg_all;                                 // two exchange buffers, used alternately
    int xb;
    float eps;
    // the four warps that share a TMEM lane quarter (same 32 rows, different column parts)
    __device__ __forceinline__ void quarter_sync() const { asm volatile("bar.sync %0, 128;" ::"r"(1 + (warp & 3)) : "memory"); }
    // ONE barrier per exchange: consecutive exchanges alternate between two buffers, and a thread can only reach the
    // exchange after the next (same buffer again) through the barrier of the next one, which every reader of this one
    // has then passed.
    __device__ __forceinline__ void row_stats(float s, float ss, float& mean, float& rstd) {
        float2* x = xchg_all + xb * (kParts * kRows);
        xb ^= 1;
        x[part * kRows + row] = make_float2(s, ss);
        quarter_sync();
#pragma unroll
        for (int p = 1; p < kParts; ++p) {
            float2 o = x[((part + p) % kParts) * kRows + row];
            s += o.x; ss += o.y;
        }
        mean = s * (1.0f / kDim);
        rstd = rsqrtf(fmaxf(ss * (1.0f / kDim) - mean * mean, 0.0f) + eps);
    }
};

// features (float32 [n][59]) -> 16-bit, K padded to 64: this thread fills chunks 2 * part, 2 * part + 1 of K-block 0.
// The global loads are issued at the very top of the kernel (they are in flight during barrier and TMEM set-up).
__device__ __forceinline__ void fetch_features(float (&f)[16], int part, const float* __restrict__ features, int64_t grow, int64_t n) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        int c = part * 16 + i;
        f[i] = (grow < n && c < kFeat) ? features[grow * kFeat + c] : 0.0f;
    }
}
// A full tile's 128 feature rows are contiguous in global memory (30,208 bytes): ONE bulk copy, issued before the TMEM /
// cluster set-up, parks them in the part of the A tile the input stage does not read (K-blocks 1..3); from there every
// thread picks its 16 columns (row stride 59 words: odd, conflict-free).  Ragged tiles and unaligned feature buffers use
// the per-thread global loads above.
constexpr int kFeatTileBytes = kRows * kFeat * 4;
constexpr int kFeatStage = 16384;                     // offset of the parking area inside an A tile
__device__ __forceinline__ void unpark_features(float (&f)[16], int row, int part, const uint8_t* a_tile) {
    const float* raw = reinterpret_cast<const float*>(a_tile + kFeatStage) + row * kFeat;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        int c = part * 16 + i;
        f[i] = c < kFeat ? raw[c] : 0.0f;
    }
}
template <bool F16>
__device__ __forceinline__ void store_features(const float (&f)[16], int row, int part, uint8_t* a_tile) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        uint32_t p[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = pack2<F16>(f[q * 8 + 2 * i], f[q * 8 + 2 * i + 1]);
        *reinterpret_cast<uint4*>(a_tile + a_tile_offset(row, part * 2 + q)) = make_uint4(p[0], p[1], p[2], p[3]);
    }
}

// Input stage epilogue, h = SiLU(LN(z + b)): Linear -> LayerNorm -> SiLU (YachtNNet.py:25-30) -> the first residual
// stream (tensor memory, half pairs) and the first trunk operand.  `loaded()` runs once the accumulator is in registers.
template <bool F16, class Loaded>
__device__ __forceinline__ void input_epilogue(Lane& L, const float* prm, uint32_t t_skip, uint8_t* a_tile, Loaded&& loaded) {
    f32x2 u[2][16];
    f32x2 ps[2] = {0ull, 0ull}, pq[2] = {0ull, 0ull};
    {
        uint32_t r0[32], r1[32];
        tmem_ld32(L.t_lane + L.col[0], r0);
        tmem_ld32(L.t_lane + L.col[1], r1);
        tmem_ld_wait();
        loaded();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint32_t (&r)[32] = c ? r1 : r0;
            const ulonglong2* bias = reinterpret_cast<const ulonglong2*>(prm + L.col[c]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const ulonglong2 b = bias[i];
                const f32x2 x0 = add2(pk2u(r[4 * i], r[4 * i + 1]), b.x), x1 = add2(pk2u(r[4 * i + 2], r[4 * i + 3]), b.y);
                ps[0] = add2(ps[0], x0); pq[0] = fma2(x0, x0, pq[0]);
                ps[1] = add2(ps[1], x1); pq[1] = fma2(x1, x1, pq[1]);
                u[c][2 * i] = x0; u[c][2 * i + 1] = x1;
            }
        }
    }
    float mean, rstd;
    L.row_stats(hsum2(ps[0], ps[1]), hsum2(pq[0], pq[1]), mean, rstd);
    const f32x2 rstd2 = pk2(rstd, rstd), nmean2 = pk2(-mean, -mean), half2 = pk2(0.5f, 0.5f);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const ulonglong2* gamma = reinterpret_cast<const ulonglong2*>(prm + kDim + L.col[c]);
        const ulonglong2* beta = reinterpret_cast<const ulonglong2*>(prm + 2 * kDim + L.col[c]);
        uint32_t hk[16], ak[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const ulonglong2 g = gamma[i], b = beta[i];
            const f32x2 ga0 = mul2(rstd2, g.x), ga1 = mul2(rstd2, g.y);
            const f32x2 h0 = silu2_from_half(mul2(half2, fma2(u[c][2 * i], ga0, fma2(nmean2, ga0, b.x))));
            const f32x2 h1 = silu2_from_half(mul2(half2, fma2(u[c][2 * i + 1], ga1, fma2(nmean2, ga1, b.y))));
            hk[2 * i] = pack2<true>(h0); hk[2 * i + 1] = pack2<true>(h1);
            if (!F16) { ak[2 * i] = pack2<false>(h0); ak[2 * i + 1] = pack2<false>(h1); }
        }
        tmem_st16(t_skip + (uint32_t)(L.col[c] / 2), hk);
        store_a_packed(a_tile, L.row, L.col[c] / 8, F16 ? hk : ak);
    }
    tmem_st_wait();
}

// Trunk epilogue, pass 1 (in place, after the tcgen05.ld): 32 accumulator columns -> x = SiLU(z + b), and the row's
// running sums of x and x^2 (four each: columns 0, 1 | 2, 3 mod 4).  half_bias = b / 2 (the host halves it: SiLU(x) =
// t + t * tanh(t), t = x / 2).
__device__ __forceinline__ void trunk_pass1(uint32_t (&v)[32], const float* half_bias, f32x2 (&ps)[2], f32x2 (&pq)[2]) {
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
// Pass 2: y = LN(x) -> the next layer's A operand in shared memory.  SECOND (fc2): h += y first, the sum goes back to the
// residual stream as half pairs (with half operands the same words are the next operand).  The residual load is issued
// before the LayerNorm arithmetic and awaited after it.
template <bool F16, bool SECOND>
__device__ __forceinline__ void trunk_pass2(uint32_t (&v)[32], float rstd, float nm, const float* gamma_p, const float* beta_p,
                                            uint32_t t_skip16, uint8_t* a_tile, int row, int c8_first) {
    const ulonglong2* gamma = reinterpret_cast<const ulonglong2*>(gamma_p);
    const ulonglong2* beta = reinterpret_cast<const ulonglong2*>(beta_p);
    const f32x2 rstd2 = pk2(rstd, rstd), nm2 = pk2(nm, nm);
    uint32_t sk[16];
    if (SECOND) tmem_ld16(t_skip16, sk);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const ulonglong2 g = gamma[i], b = beta[i];
        upk2u(fma2(fma2(pk2u(v[4 * i], v[4 * i + 1]), rstd2, nm2), g.x, b.x), v[4 * i], v[4 * i + 1]);
        upk2u(fma2(fma2(pk2u(v[4 * i + 2], v[4 * i + 3]), rstd2, nm2), g.y, b.y), v[4 * i + 2], v[4 * i + 3]);
    }
    uint32_t ak[16];
    if (SECOND) {
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const f32x2 h = add2(pk2u(v[2 * i], v[2 * i + 1]), unpack_h2(sk[i]));
            sk[i] = pack2<true>(h);
            if (!F16) ak[i] = pack2<false>(h);
        }
        tmem_st16(t_skip16, sk);
        store_a_packed(a_tile, row, c8_first, F16 ? sk : ak);
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) ak[i] = pack2<F16>(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
        store_a_packed(a_tile, row, c8_first, ak);
    }
}

// Heads: a = SiLU(LN(h; gamma, beta)).  Both heads normalise the same h (YachtNNet.py:38-50): one statistics pass, then
// the value head's activations go to the shared-memory A tile and the policy head's to tensor memory (t_api, packed
// 16-bit pairs: the A operand is read from TMEM).
template <bool F16>
__device__ __forceinline__ void head_prep(Lane& L, uint32_t t_skip, const float* gv, const float* bv, const float* gp, const float* bp,
                                          uint8_t* a_tile, uint32_t t_api) {
    f32x2 u[2][16];
    f32x2 ps[2] = {0ull, 0ull}, pq[2] = {0ull, 0ull};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t h[16];
        tmem_ld16(t_skip + (uint32_t)(L.col[c] / 2), h);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const f32x2 x0 = unpack_h2(h[2 * i]), x1 = unpack_h2(h[2 * i + 1]);
            ps[0] = add2(ps[0], x0); pq[0] = fma2(x0, x0, pq[0]);
            ps[1] = add2(ps[1], x1); pq[1] = fma2(x1, x1, pq[1]);
            u[c][2 * i] = x0; u[c][2 * i + 1] = x1;
        }
    }
    float mean, rstd;
    L.row_stats(hsum2(ps[0], ps[1]), hsum2(pq[0], pq[1]), mean, rstd);
    const float nm = -mean * rstd;
    const f32x2 rstd2 = pk2(rstd, rstd), nm2 = pk2(nm, nm), half2 = pk2(0.5f, 0.5f);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const ulonglong2* gv2 = reinterpret_cast<const ulonglong2*>(gv + L.col[c]);
        const ulonglong2* bv2 = reinterpret_cast<const ulonglong2*>(bv + L.col[c]);
        const ulonglong2* gp2 = reinterpret_cast<const ulonglong2*>(gp + L.col[c]);
        const ulonglong2* bp2 = reinterpret_cast<const ulonglong2*>(bp + L.col[c]);
        uint32_t av[16], ap[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const ulonglong2 g1 = gv2[i], b1 = bv2[i], g2 = gp2[i], b2 = bp2[i];
            const f32x2 x0 = fma2(u[c][2 * i], rstd2, nm2), x1 = fma2(u[c][2 * i + 1], rstd2, nm2);
            av[2 * i] = pack2<F16>(silu2_from_half(mul2(half2, fma2(x0, g1.x, b1.x))));
            av[2 * i + 1] = pack2<F16>(silu2_from_half(mul2(half2, fma2(x1, g1.y, b1.y))));
            ap[2 * i] = pack2<F16>(silu2_from_half(mul2(half2, fma2(x0, g2.x, b2.x))));
            ap[2 * i + 1] = pack2<F16>(silu2_from_half(mul2(half2, fma2(x1, g2.y, b2.y))));
        }
        store_a_packed(a_tile, L.row, L.col[c] / 8, av);
        tmem_st16(t_api + (uint32_t)(L.col[c] / 2), ap);              // K elements col.. = packed columns col / 2..
    }
    tmem_st_wait();
}

// Value head after its MMA (YachtNNet.py:44-50): SiLU(z + b1) . w2 over this thread's 32 of the 128 columns
__device__ __forceinline__ float value_partial(const Lane& L, uint32_t t_acc, const float* prm) {
    uint32_t r[32];
    tmem_ld32(t_acc + L.part * 32, r);
    tmem_ld_wait();
    const float* b1 = prm + 2 * kDim + L.part * 32;
    const float* w2 = prm + 2 * kDim + 128 + L.part * 32;
    float acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        float t = silu_from_half(fmaf(__uint_as_float(r[i]), 0.5f, 0.5f * b1[i]));
        acc[i & 3] = fmaf(t, w2[i], acc[i & 3]);
    }
    return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// Where a leaf's policy logits go.  Scatter mode (the MCTS path): the row's LEGAL logits go straight into its leaf's row
// in the tree pool (the row layout ya_mcts_select allocated: csrc/ya_mcts.cu "Logit area"), instead of a dense
// [n][3232] matrix.  dst = 0: the row needs no evaluation (its descent ended in a terminal / dead-end node).  The policy
// row is 13 runs of columns -- run 0 = the bid moves [0, 202), run 1 + c = category c, [202 + 252 c, +252) -- and the leaf
// wants: a bid row run 0 (stored as columns [0, 208), as they are), a ten-dice row the runs of its open categories (one
// 272-slot block per open category, in order: the 16-aligned column window around the run, so that every 16-column
// sector keeps its alignment), a five-dice row only the first column of every open category (compact).  `runs` = the
// wanted runs of the first two kinds; every lane of a warp then walks the same predicated code, whatever mix of leaves
// the warp holds.
struct RowOut {
    uint16_t* s_dst;
    uint32_t runs, open5;
    uint16_t* dense;                                  // this row of the dense logit matrix, or null
    uint32_t row_mx;                                  // running row maximum as a packed 16-bit pair
    bool scatter;
    __device__ __forceinline__ void init(const uint64_t* scatter_dst, const uint32_t* scatter_desc, uint16_t* logits, int64_t grow, int64_t n, uint32_t neg_inf2) {
        s_dst = nullptr; runs = 0; open5 = 0; row_mx = neg_inf2;
        scatter = scatter_dst != nullptr;
        dense = (logits && grow < n) ? logits + grow * kPolicyCols : nullptr;
        if (scatter_dst && grow < n) {
            s_dst = reinterpret_cast<uint16_t*>(scatter_dst[grow]);
            const uint32_t s_desc = s_dst ? scatter_desc[grow] : 0u;
            const uint32_t open = (s_desc >> 1) & 0xFFFu;             // bit c = category c open
            if (s_desc & 1u) runs = 1u;
            else if (s_desc >> 13) runs = open << 1;
            else open5 = s_desc ? open : 0u;
        }
    }
};
// One chunk of a policy tile: 32 accumulator columns [col0, col0 + 32) of this thread's row -> bias (the host sets the
// bias of the padding columns >= 3226 to -inf, so they can never win the maximum), 16-bit packing, packed maximum
// (rounding is monotone, so the largest rounded logit is the rounded largest logit: the value the expand kernel needs),
// stores.
template <bool F16>
__device__ __forceinline__ void policy_chunk(const uint32_t (&r)[32], int col0, const float* bias_all, RowOut& o) {
    constexpr uint32_t kNegInf2 = F16 ? 0xFC00FC00u : 0xFF80FF80u;
    const ulonglong2* bias = reinterpret_cast<const ulonglong2*>(bias_all + col0);
    uint32_t pk[16];
    uint32_t m0 = o.row_mx, m1 = kNegInf2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const ulonglong2 b = bias[i];
        pk[2 * i] = pack2<F16>(add2(pk2u(r[4 * i], r[4 * i + 1]), b.x));
        pk[2 * i + 1] = pack2<F16>(add2(pk2u(r[4 * i + 2], r[4 * i + 3]), b.y));
        m0 = max16x2<F16>(m0, pk[2 * i]);
        m1 = max16x2<F16>(m1, pk[2 * i + 1]);
    }
    o.row_mx = max16x2<F16>(m0, m1);
    auto store_sector = [&](uint16_t* dst, int h) {                    // 16 columns = one aligned 32-byte store
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst),
                     "r"(pk[8 * h]), "r"(pk[8 * h + 1]), "r"(pk[8 * h + 2]), "r"(pk[8 * h + 3]),
                     "r"(pk[8 * h + 4]), "r"(pk[8 * h + 5]), "r"(pk[8 * h + 6]), "r"(pk[8 * h + 7]) : "memory");
    };
    if (o.scatter) {
        // Bid and ten-dice rows: each of the chunk's two 16-column sectors lies in one or two runs (a sector that holds a
        // run boundary may be wanted by both neighbours); which runs is the same for every lane, whether the lane's leaf
        // wants them is one bit of `runs`.
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int cs = col0 + 16 * h;
            const int ra = ((cs + 50) * 4162) >> 20, rb = ((cs + 65) * 4162) >> 20;   // runs of the sector's first / last column
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int rr = t ? rb : ra;
                if ((t == 1 && rb == ra) || rr > 12) continue;
                if ((o.runs >> rr) & 1u) {
                    const int start = rr ? 202 + 252 * (rr - 1) : 0;
                    const int before = __popc(o.runs & ((1u << rr) - 1u));   // wanted runs in front of this one
                    store_sector(o.s_dst + 272 * before + (cs - (start & ~15)), h);
                }
            }
        }
        // Five-dice rows: subset 0 of every open category, i.e. the first column of a run, if one starts inside this
        // chunk (run starts are even columns: the low half of a packed pair).
        const int r0 = ((col0 + 50) * 4162) >> 20;
        const int end0 = 202 + 252 * r0;                               // category r0 starts here
        if (end0 < col0 + 32 && r0 < 12 && ((o.open5 >> r0) & 1u)) {
            uint16_t* one = o.s_dst + __popc(o.open5 & ((1u << r0) - 1u));
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (col0 + 2 * i == end0) *one = (uint16_t)(pk[i] & 0xFFFFu);
        }
    }
    if (o.dense) {
        store_sector(o.dense + col0, 0);
        store_sector(o.dense + col0 + 16, 1);
    }
}

// ================================================================================================ one tile per CTA
// mbarriers (per CTA; "leader only" ones are used in rank 0's copy)
enum { B_W = 0,         // [2] weights + parameters of a stage landed in buffer stage & 1 (local bulk copies)
       B_MMA = 2,       // [2] MMAs of the first / second N = 128 half done (commit multicast to both CTAs)
       B_PEER = 4,      // leader only: the other CTA's A tile and weights of this stage are in place
       B_SLOT = 5,      // [6] policy head: this CTA's half of a weight tile landed in the slot
       B_PSLOT = 11,    // [6] leader only: the other CTA's half landed
       B_ACC = 17,      // [3] policy head: accumulator ready (commit multicast)
       B_DRAIN = 20,    // [3] leader only: accumulator read out by the 15 epilogue warps of BOTH CTAs
       B_FEAT = 23 };   // the tile's feature rows landed in the parking area

template <bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
ya_k_forward(const float* __restrict__ features, uint16_t* __restrict__ logits, float* __restrict__ values,
             float* __restrict__ row_max, const uint8_t* __restrict__ wblob, const float* __restrict__ pblob, Blob off, int nblocks, int64_t n, float eps,
             const uint64_t* __restrict__ scatter_dst, const uint32_t* __restrict__ scatter_desc, bool feat_aligned) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_tile = base;
    uint8_t* w_tiles = base + kABytes;                                // two 64 KB weight buffers
    float* prm_all = reinterpret_cast<float*>(base + 3 * kABytes);
    float* pi_prm = prm_all + 2 * kPrmFloats;
    float2* xchg = reinterpret_cast<float2*>(pi_prm + kPiPrmFloats);
    uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + 2 * kParts * kRows);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane;
    const int part = warp >> 2;
    const int64_t grow = (int64_t)blockIdx.x * kRows + row;
    const bool producer = warp == kIssuerWarp;                        // also an epilogue warp, except in the policy head
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const bool feat_bulk = feat_aligned && (int64_t)(blockIdx.x + 1) * kRows <= n;   // a full tile of an aligned feature matrix
    float feat[16];
    if (!feat_bulk) fetch_features(feat, part, features, grow, n);

    if (tid == 0) {
        for (int i = 0; i < B_DRAIN; ++i) mbar_init(&bars[i], 1);
        for (int i = B_DRAIN; i < B_DRAIN + 3; ++i) mbar_init(&bars[i], 30);
        mbar_init(&bars[B_FEAT], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (feat_bulk) {
            proxy_fence();
            mbar_expect_tx(&bars[B_FEAT], kFeatTileBytes);
            bulk_g2s(a_tile + kFeatStage, features + (int64_t)blockIdx.x * kRows * kFeat, kFeatTileBytes, &bars[B_FEAT]);
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync();                                                   // both CTAs' barriers exist before anyone arrives remotely
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    Lane L;
    L.row = row; L.part = part; L.warp = warp;
    L.col[0] = part * 32; L.col[1] = 128 + part * 32;
    L.t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    L.xchg_all = xchg; L.xb = 0; L.eps = eps;
    const uint32_t t_lane = L.t_lane;
    const uint32_t t_skip = t_lane + kSkipCol;                        // residual stream: TMEM columns 256..383 (half pairs)
#ifdef YA_FWD_TIMELINE
    if (tid == 0 && blockIdx.x < 1024) {
        unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_));
        g_cta_times[4 * blockIdx.x + 3] = sm_;
    }
    YA_CTA_TIME(0);
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

    // ---------------------------------------------------------------- input stage
    if (producer && elect_one()) {
        load_stage(0);
        load_stage(1);                                                // first trunk layer (or the value head) streams in
    }
    if (feat_bulk) {
        mbar_wait(&bars[B_FEAT], 0u);
        unpark_features(feat, row, part, a_tile);
    }
    store_features<F16>(feat, row, part, a_tile);
    begin_stage(0);
    wait_mma(0, 0);
    wait_mma(0, 1);
    YA_STAMP();                                                       // [3k+2] MMA done
    input_epilogue<F16>(L, prm_buf(0), t_skip, a_tile, [] {});

    // ---------------------------------------------------------------- residual trunk
    for (int l = 0; l < layers; ++l) {
        const int stage = l + 1;
        const float* prm = prm_buf(stage);                            // bias / 2 | gamma | beta
        const bool second = l & 1;                                    // fc2: add the skip connection
        // Two N = 128 halves, each with its own completion barrier: the epilogue's first pass over columns
        // 0..127 runs while the tensor core works on columns 128..255.  The 64 activations stay in registers across
        // the statistics exchange (no TMEM round trip).
        begin_stage(stage);
        uint32_t v0[32], v1[32];
        f32x2 ps[2] = {0ull, 0ull}, pq[2] = {0ull, 0ull};
        YA_STAMP2();                                                  // [8l] MMAs issued
        wait_mma(stage, 0);
        YA_STAMP2();                                                  // [8l+1] half 0 ready
        tmem_ld32(t_lane + L.col[0], v0);
        tmem_ld_wait();
        trunk_pass1(v0, prm + L.col[0], ps, pq);
        YA_STAMP2();                                                  // [8l+2] pass 1 of half 0 done
        wait_mma(stage, 1);
        YA_STAMP();
        YA_STAMP2();                                                  // [8l+3] half 1 ready
        tmem_ld32(t_lane + L.col[1], v1);
        tmem_ld_wait();
        trunk_pass1(v1, prm + L.col[1], ps, pq);
        float mean, rstd;
        YA_STAMP2();                                                  // [8l+4] pass 1 of half 1 done
        L.row_stats(hsum2(ps[0], ps[1]), hsum2(pq[0], pq[1]), mean, rstd);
        YA_STAMP2();                                                  // [8l+5] statistics exchanged
        const float nm = -mean * rstd;
        if (second) {
            trunk_pass2<F16, true>(v0, rstd, nm, prm + kDim + L.col[0], prm + 2 * kDim + L.col[0], t_skip + L.col[0] / 2, a_tile, row, L.col[0] / 8);
            trunk_pass2<F16, true>(v1, rstd, nm, prm + kDim + L.col[1], prm + 2 * kDim + L.col[1], t_skip + L.col[1] / 2, a_tile, row, L.col[1] / 8);
            tmem_st_wait();
        } else {
            trunk_pass2<F16, false>(v0, rstd, nm, prm + kDim + L.col[0], prm + 2 * kDim + L.col[0], 0u, a_tile, row, L.col[0] / 8);
            trunk_pass2<F16, false>(v1, rstd, nm, prm + kDim + L.col[1], prm + 2 * kDim + L.col[1], 0u, a_tile, row, L.col[1] / 8);
        }
        YA_STAMP2();                                                  // [8l+6] pass 2 done
        YA_STAMP2();
    }
    YA_CTA_TIME(1);

    // ---------------------------------------------------------------- heads
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
    // value head (YachtNNet.py:44-50): LN -> SiLU -> Linear(256,128) -> SiLU -> Linear(128,1) -> tanh
    {
        const int stage = layers + 1;
        const float* prm = prm_buf(stage);                            // gamma_v | beta_v | b1[128] | w2[128] | b2
        mbar_wait(&bars[B_W + (stage & 1)], (uint32_t)((stage >> 1) & 1));   // the LayerNorm parameters travel with the weights
        head_prep<F16>(L, t_skip, prm, prm + kDim, pi_prm, pi_prm + kDim, a_tile, t_lane);
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
        const float dot = value_partial(L, t_lane + 128, prm);
        xchg[part * kRows + row] = make_float2(dot, 0.0f);
        __syncthreads();
        if (part == 0 && grow < n) {
            float s = dot + xchg[1 * kRows + row].x + xchg[2 * kRows + row].x + xchg[3 * kRows + row].x + prm[2 * kDim + 256];
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(s));
            values[grow] = th;
        }
        __syncthreads();
    }

    // policy head (YachtNNet.py:38-42): Linear(256, 3226) on the activations already sitting in tensor memory, as 26
    // tiles of 128 columns.  Weight tiles rotate over six 32 KB slots and accumulators over TMEM columns 128 / 256 / 384:
    // as soon as tile j's MMAs retire, tile j + 5's weights start streaming into the slot they read, and the MMAs of the
    // following tiles run under the epilogue of tile j.
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
        __syncthreads();
        tc_fence_after();
        constexpr uint32_t kNegInf2 = F16 ? 0xFC00FC00u : 0xFF80FF80u;
        RowOut out;
        out.init(scatter_dst, scatter_desc, logits, grow, n, kNegInf2);
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
            // Per tile a warp pulls its 32 (64) accumulator columns into registers and hands the accumulator back at once.
            const bool q3 = (warp & 3) == 3;
            for (int j = 0; j < kPolicyTiles; ++j) {
                const int n_my = (q3 && j % 3 == part) ? 2 : 1;
                mbar_wait(&bars[B_ACC + j % 3], (uint32_t)((j / 3) & 1));
                tc_fence_after();
                YA_STAMP();                                           // policy tile j: accumulator ready
                uint32_t acc[2][32];
                tmem_ld32(t_lane + kPolicyTile + (j % 3) * kPolicyTile + part * 32, acc[0]);
                if (n_my == 2) tmem_ld32(t_lane + kPolicyTile + (j % 3) * kPolicyTile + 96, acc[1]);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {                                      // this warp's share of the accumulator is in registers
                    if (leader) mbar_arrive(&bars[B_DRAIN + j % 3]);
                    else mbar_arrive_remote(&bars[B_DRAIN + j % 3], 0);
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (q >= n_my) break;
                    const int col0 = j * kPolicyTile + (q ? 96 : part * 32);
                    if (col0 < kPolicyCols) policy_chunk<F16>(acc[q], col0, bias_all, out);
                }
            }
        }
        xchg[part * kRows + row].x = hmax16x2<F16>(out.row_mx);       // (the producer warp's entry stays -inf)
        YA_STAMP();
        // the row's largest logit as the expand kernel will see it
        __syncthreads();
        if (part == 0 && grow < n && row_max)
            row_max[grow] = fmaxf(fmaxf(xchg[0 * kRows + row].x, xchg[1 * kRows + row].x), fmaxf(xchg[2 * kRows + row].x, xchg[3 * kRows + row].x));
    }
#ifdef YA_FWD_TIMELINE
    YA_CTA_TIME(2);
    if (blockIdx.x == 0 && tid == 0) {
        unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        g_timeline2[1002] = t_; g_timeline2[1003] = (unsigned long long)clock64();
    }
#endif
    tc_fence_before();
    cluster_sync();                                                   // neither CTA leaves while the other may still signal it
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ================================================================================================ two tiles per CTA
// Weight PIECES, consumed in this order by both tiles: 0 = input image (16 KB), 1 + 2 l + h = half h of trunk layer l
// (32 KB), 2 layers + 1 = value head (32 KB).  Piece p lives in weight slot p & 1; piece p + 2 is requested as soon as
// tile Y's MMAs on piece p retire.  The parameters of a stage ride with its first piece (piece 0 / the odd pieces).
// MMA GROUPS in tensor-pipe order: g = 2 s + t for tile t of stage s (0 = input, 1..layers = trunk, layers + 1 = value).
enum { C_WL = 0,        // [2] weight piece (+ parameters) landed in slot p & 1
       C_MMA = 2,       // [2 t + h] MMAs of tile t, half h of the current stage done (commit multicast)
       C_FREE = 6,      // [2] accumulator half h pulled into registers by all 16 warps of this CTA (count 16)
       C_PEER = 8,      // [2] leader only: rank 1 is ready for the next group's half h
       C_SLOT = 10,     // [6] policy head: this CTA's half of a weight tile landed in the slot
       C_PSLOT = 16,    // [6] leader only: the other CTA's half landed
       C_ACC = 22,      // [2] policy head: accumulator of tile t ready (commit multicast)
       C_DRAIN = 24,    // [2] leader only: accumulator of tile t read out by the 15 epilogue warps of BOTH CTAs (count 30)
       C_VPEER = 29,    // [2] leader only: rank 1 is ready for tile t's value-head MMAs.  Own barriers: the two value hand-offs follow
                        //     each other with no MMA completion in between, so on a shared barrier rank 1 could report twice before
                        //     a slower leader has looked once -- and the leader would wait for the parity of a phase long gone
       C_FEAT = 31,     // both tiles' feature rows landed in their parking areas
       C_RET = 26 };    // [3] policy head: both tiles' MMAs on weight tile j retired (slot j % 3; commit multicast).  A barrier of its
                        //     own, three tiles deep: a producer that tests "tile j - 1 retired" on C_ACC a moment after tile j has
                        //     ALSO retired would see the parity of two phases ago as pending and wait for a tile it has to load first

template <bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
ya_k_forward2(const float* __restrict__ features, uint16_t* __restrict__ logits, float* __restrict__ values,
              float* __restrict__ row_max, const uint8_t* __restrict__ wblob, const float* __restrict__ pblob, Blob off, int nblocks, int64_t n, float eps,
              const uint64_t* __restrict__ scatter_dst, const uint32_t* __restrict__ scatter_desc, bool feat_aligned) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* w_slots = base + 2 * kABytes;                            // A tiles of X and Y, then two 32 KB weight slots
    float* prm_all = reinterpret_cast<float*>(base + 3 * kABytes);
    float* pi_prm = prm_all + 2 * kPrmFloats;
    float2* xchg = reinterpret_cast<float2*>(pi_prm + kPiPrmFloats);
    uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + 2 * kParts * kRows);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane;
    const int part = warp >> 2;
    const bool producer = warp == kIssuerWarp;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    auto a_tile = [&](int t) { return base + t * kABytes; };
    auto grow_of = [&](int t) { return ((int64_t)blockIdx.x * 2 + t) * kRows + row; };
    const bool feat_bulk = feat_aligned && ((int64_t)blockIdx.x * 2 + 2) * kRows <= n;   // two full tiles of an aligned feature matrix
    float feat0[16], feat1[16];
    if (!feat_bulk) {
        fetch_features(feat0, part, features, grow_of(0), n);
        fetch_features(feat1, part, features, grow_of(1), n);
    }

    if (tid == 0) {
        for (int i = 0; i < kBars; ++i) mbar_init(&bars[i], (i == C_FREE || i == C_FREE + 1) ? 16 : (i == C_DRAIN || i == C_DRAIN + 1) ? 30 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (feat_bulk) {
            proxy_fence();
            mbar_expect_tx(&bars[C_FEAT], 2 * kFeatTileBytes);
            const float* src = features + (int64_t)blockIdx.x * 2 * kRows * kFeat;
            bulk_g2s(base + kFeatStage, src, kFeatTileBytes, &bars[C_FEAT]);
            bulk_g2s(base + kABytes + kFeatStage, src + kRows * kFeat, kFeatTileBytes, &bars[C_FEAT]);
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    Lane L;
    L.row = row; L.part = part; L.warp = warp;
    L.col[0] = part * 32; L.col[1] = 128 + part * 32;
    L.t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    L.xchg_all = xchg; L.xb = 0; L.eps = eps;
    const uint32_t t_lane = L.t_lane;
    auto t_skip = [&](int t) { return t_lane + kSkipCol + 128 * t; };
#ifdef YA_FWD_TIMELINE
    if (tid == 0 && blockIdx.x < 1024) {
        unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_));
        g_cta_times[4 * blockIdx.x + 3] = sm_;
    }
    YA_CTA_TIME(0);
    int tl_n = 0, tl2_n = 0;
    if (blockIdx.x == 0 && tid == 0) {
        unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        g_timeline2[1000] = t_; g_timeline2[1001] = (unsigned long long)clock64();
    }
#endif
    const int layers = 2 * nblocks;
    const int last_piece = 2 * layers + 1;
    auto prm_buf = [&](int s) { return prm_all + (s & 1) * kPrmFloats; };
    auto piece_bar = [&](int p) { return &bars[C_WL + (p & 1)]; };
    auto piece_par = [](int p) { return (uint32_t)((p >> 1) & 1); };
    // elected producer thread: this CTA's half of piece p (+ the parameters of the stage it opens)
    auto load_piece = [&](int p) {
        if (p > last_piece) return;
        uint64_t* bar = piece_bar(p);
        uint8_t* wd = w_slots + (p & 1) * kSlotBytes;
        if (p == 0) {
            mbar_expect_tx(bar, 16384 + 3 * kDim * 4);
            bulk_g2s(wd, wblob + off.w_in + rank * 16384, 16384, bar);
            bulk_g2s(prm_buf(0), pblob + off.p_in, 3 * kDim * 4, bar);
        } else if (p == last_piece) {                                 // value head weights; both heads' parameters
            mbar_expect_tx(bar, 32768 + 772 * 4 + 2 * kDim * 4);
            bulk_g2s(wd, wblob + off.w_v + rank * 32768, 32768, bar);
            bulk_g2s(prm_buf(layers + 1), pblob + off.p_v, 772 * 4, bar);
            bulk_g2s(pi_prm, pblob + off.p_pi_ln, 2 * kDim * 4, bar);
        } else {
            const int l = (p - 1) >> 1, h = (p - 1) & 1;
            mbar_expect_tx(bar, kSlotBytes + (h == 0 ? 3 * kDim * 4 : 0));
            bulk_g2s(wd, wblob + off.w_trunk + (int64_t)l * kWBytes + rank * kWHalf + h * kSlotBytes, kSlotBytes, bar);
            if (h == 0) bulk_g2s(prm_buf(l + 1), pblob + off.p_trunk + (int64_t)l * 3 * kDim, 3 * kDim * 4, bar);
        }
    };
    // Producer warp, both CTAs: hand-off for half h (or, whole = true, both halves at once) of MMA group (t, s).  Local
    // conditions first -- the accumulator half is free in this CTA (`free_phase` >= 0: phase of C_FREE to wait for), the
    // weight piece has landed -- then rank 1 reports to the leader and the leader, once rank 1 has reported, issues.
    // The operand tile of (t, s) is complete by program order (a CTA barrier closed the epilogue that wrote it).
    auto issue_group = [&](int t, int s, int h, bool whole, int free_phase) {
        const int g = 2 * s + t;
        const int piece = s == 0 ? 0 : (s <= layers ? 2 * s - 1 + h : last_piece);
        if (free_phase >= 0) {
            mbar_wait(&bars[C_FREE + h], (uint32_t)(free_phase & 1));
            if (whole) mbar_wait(&bars[C_FREE + 1], (uint32_t)(free_phase & 1));
        }
        mbar_wait(piece_bar(piece), piece_par(piece));
        const bool value_head = s > layers;
        if (!leader) {
            if (elect_one()) {
                if (value_head) mbar_arrive_remote(&bars[C_VPEER + t], 0);
                else {
                    mbar_arrive_remote(&bars[C_PEER + h], 0);
                    if (whole) mbar_arrive_remote(&bars[C_PEER + 1], 0);
                }
            }
        } else {
            if (value_head) mbar_wait(&bars[C_VPEER + t], 0u);
            else {
                mbar_wait(&bars[C_PEER + h], (uint32_t)(g & 1));
                if (whole) mbar_wait(&bars[C_PEER + 1], (uint32_t)(g & 1));
            }
            tc_fence_after();
            if (elect_one()) {
                const uint64_t da = umma_desc(smem_u32(a_tile(t))), db = umma_desc(smem_u32(w_slots + (piece & 1) * kSlotBytes));
                if (s == 0) {                                         // K = 64, N = 256: each CTA holds 128 weight rows
                    const uint32_t idesc = umma_idesc(kDim, F16, 256);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma2(tmem, umma_desc_advance(da, k * 32), umma_desc_advance(db, k * 32), (uint32_t)(k != 0), idesc);
                    umma2_commit(&bars[C_MMA + 2 * t]);
                    umma2_commit(&bars[C_MMA + 2 * t + 1]);
                } else {                                              // trunk half / value head: N = 128, each CTA holds 64 weight rows
                    const uint32_t idesc = umma_idesc(128, F16, 256);
                    const uint32_t d = s <= layers ? tmem + h * 128 : tmem + kSkipCol + 128 * t;   // value head: over the dead residual stream
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma2(d, umma_desc_advance(da, kb * (kRows * 128) + k * 32),
                                  umma_desc_advance(db, kb * 8192 + k * 32), (uint32_t)((kb | k) != 0), idesc);
                    umma2_commit(&bars[C_MMA + 2 * t + h]);
                    if (whole) umma2_commit(&bars[C_MMA + 2 * t + 1]);
                }
            }
        }
        __syncwarp();
    };
    auto wait_mma = [&](int t, int s, int h) {
        mbar_wait(&bars[C_MMA + 2 * t + h], (uint32_t)(s & 1));
        tc_fence_after();
    };
    auto arrive_free = [&](int h) {                                    // this warp has its share of accumulator half h in registers
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[C_FREE + h]);
    };
    auto close_epilogue = [&] {                                        // the operand tile just written is complete in this CTA
        proxy_fence();
        tc_fence_before();
        __syncthreads();
    };

    // ---------------------------------------------------------------- input stage
    if (producer && elect_one()) {
        load_piece(0);
        load_piece(1);
    }
    YA_STAMP();                                                        // [0] start
    if (feat_bulk) {
        mbar_wait(&bars[C_FEAT], 0u);
        unpark_features(feat0, row, part, a_tile(0));
        unpark_features(feat1, row, part, a_tile(1));
    }
    store_features<F16>(feat0, row, part, a_tile(0));
    store_features<F16>(feat1, row, part, a_tile(1));
    close_epilogue();
    YA_STAMP();                                                        // [1] features in shared memory
    if (producer) issue_group(0, 0, 0, true, -1);
    mbar_wait(piece_bar(0), piece_par(0));                             // stage 0's parameters
    YA_STAMP();                                                        // [2] input weights and parameters landed
    int ev = 0;                                                        // epilogues so far = phase of C_FREE
    for (int t = 0; t < 2; ++t) {
        wait_mma(t, 0, 0);
        wait_mma(t, 0, 1);
        YA_STAMP();                                                    // [3], [5] input MMA of tile t done
        if (t == 1 && producer && elect_one()) load_piece(2);          // the input image is no longer read
        __syncwarp();
        input_epilogue<F16>(L, prm_buf(0), t_skip(t), a_tile(t), [&] {
            arrive_free(0);
            arrive_free(1);
            if (producer) {
                if (t == 0) issue_group(1, 0, 0, true, ev);
                else if (layers > 0) { issue_group(0, 1, 0, false, ev); issue_group(0, 1, 1, false, ev); }
            }
        });
        ++ev;
        close_epilogue();
        YA_STAMP();                                                    // [4], [6] input epilogue of tile t closed
    }

    // ---------------------------------------------------------------- residual trunk: E(X, l), E(Y, l), E(X, l + 1), ...
    for (int l = 0; l < layers; ++l) {
        const int stage = l + 1;
        const float* prm = prm_buf(stage);
        const bool second = l & 1;
        mbar_wait(piece_bar(2 * stage - 1), piece_par(2 * stage - 1)); // this stage's parameters (every thread, once per stage)
        for (int t = 0; t < 2; ++t) {
            // the group that follows (t, stage) on the tensor pipe: the other tile, same stage (t = 0) or next stage (t = 1)
            const int nt = t ^ 1, ns = stage + t;
            const bool has_next = ns <= layers;
            uint32_t v0[32], v1[32];
            f32x2 ps[2] = {0ull, 0ull}, pq[2] = {0ull, 0ull};
            YA_STAMP2();
            wait_mma(t, stage, 0);
            YA_STAMP2();
            if (t == 1 && producer && elect_one()) load_piece(2 * stage + 1);   // tile Y is done with this layer's first half
            __syncwarp();
            tmem_ld32(t_lane + L.col[0], v0);
            tmem_ld_wait();
            arrive_free(0);
            if (producer && has_next) issue_group(nt, ns, 0, false, ev);
            trunk_pass1(v0, prm + L.col[0], ps, pq);
            YA_STAMP2();
            wait_mma(t, stage, 1);
            YA_STAMP2();
            if (t == 1 && producer && elect_one()) load_piece(2 * stage + 2);
            __syncwarp();
            tmem_ld32(t_lane + L.col[1], v1);
            tmem_ld_wait();
            arrive_free(1);
            if (producer && has_next) issue_group(nt, ns, 1, false, ev);
            ++ev;
            trunk_pass1(v1, prm + L.col[1], ps, pq);
            float mean, rstd;
            YA_STAMP2();
            L.row_stats(hsum2(ps[0], ps[1]), hsum2(pq[0], pq[1]), mean, rstd);
            YA_STAMP2();
            const float nm = -mean * rstd;
            uint8_t* at = a_tile(t);
            const uint32_t ts = t_skip(t);
            if (second) {
                trunk_pass2<F16, true>(v0, rstd, nm, prm + kDim + L.col[0], prm + 2 * kDim + L.col[0], ts + L.col[0] / 2, at, row, L.col[0] / 8);
                trunk_pass2<F16, true>(v1, rstd, nm, prm + kDim + L.col[1], prm + 2 * kDim + L.col[1], ts + L.col[1] / 2, at, row, L.col[1] / 8);
                tmem_st_wait();
            } else {
                trunk_pass2<F16, false>(v0, rstd, nm, prm + kDim + L.col[0], prm + 2 * kDim + L.col[0], 0u, at, row, L.col[0] / 8);
                trunk_pass2<F16, false>(v1, rstd, nm, prm + kDim + L.col[1], prm + 2 * kDim + L.col[1], 0u, at, row, L.col[1] / 8);
            }
            YA_STAMP2();
            close_epilogue();
            YA_STAMP2();
        }
    }
    YA_CTA_TIME(1);

    // ---------------------------------------------------------------- heads
    // Operand space in the policy head: slots 0, 1 = A tile of X, 2, 3 = A tile of Y, 4, 5 = the weight slots.  Slot 4 is
    // free now (the last trunk half has retired), 0 and 1 once X's value MMAs retire, 2, 3 and 5 (value weights) after Y's.
    auto slot_of = [](int j) { constexpr int m[kSlots] = {4, 0, 1, 2, 3, 5}; return m[j % kSlots]; };
    auto load_tile = [&](int j) {
        uint64_t* bar = &bars[C_SLOT + slot_of(j)];
        mbar_expect_tx(bar, kSlotBytes + (j == 0 ? kPolicyTiles * kPolicyTile * 4 : 0));
        bulk_g2s(base + slot_of(j) * kSlotBytes, wblob + off.w_pi + (int64_t)j * 65536 + rank * kSlotBytes, kSlotBytes, bar);
        if (j == 0) bulk_g2s(pi_prm + 2 * kDim, pblob + off.p_pi_bias, kPolicyTiles * kPolicyTile * 4, bar);
    };
    {
        const int stage = layers + 1;
        const float* prm = prm_buf(stage);                            // gamma_v | beta_v | b1[128] | w2[128] | b2
        mbar_wait(piece_bar(last_piece), piece_par(last_piece));
        if (producer && elect_one()) load_tile(0);
        __syncwarp();
        // policy A operands: X in TMEM columns [0, 128), Y in [128, 256) (the shared accumulator is dead); value
        // accumulators over the residual streams, each dead once its tile's heads are prepared
        YA_STAMP();                                                    // [7] trunk done, head parameters landed
        for (int t = 0; t < 2; ++t) {
            head_prep<F16>(L, t_skip(t), prm, prm + kDim, pi_prm, pi_prm + kDim, a_tile(t), t_lane + 128 * t);
            close_epilogue();
            if (producer) issue_group(t, stage, 0, true, -1);
            YA_STAMP();                                                // [8], [9] heads of tile t prepared
        }
        for (int t = 0; t < 2; ++t) {
            wait_mma(t, stage, 0);
            wait_mma(t, stage, 1);
            if (producer && elect_one()) {
                if (t == 0) { load_tile(1); load_tile(2); }
                else { load_tile(3); load_tile(4); load_tile(5); }
            }
            __syncwarp();
            const float dot = value_partial(L, t_skip(t), prm);
            xchg[part * kRows + row] = make_float2(dot, 0.0f);
            __syncthreads();
            const int64_t grow = grow_of(t);
            if (part == 0 && grow < n) {
                float s = dot + xchg[1 * kRows + row].x + xchg[2 * kRows + row].x + xchg[3 * kRows + row].x + prm[2 * kDim + 256];
                float th;
                asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(s));
                values[grow] = th;
            }
            __syncthreads();
            YA_STAMP();                                                // [10], [11] value of tile t out
        }
    }

    // policy head: per weight slot (tile j of 128 columns) the MMAs of X then Y, each into its own accumulator (TMEM
    // columns 256 / 384); (t, j + 1) is issued once both CTAs' epilogues have pulled (t, j) out.
    {
        mbar_wait(&bars[C_SLOT + slot_of(0)], 0);                     // tile 0 and every bias landed
        const float* bias_all = pi_prm + 2 * kDim;
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        constexpr uint32_t kNegInf2 = F16 ? 0xFC00FC00u : 0xFF80FF80u;
        RowOut out[2];
        out[0].init(scatter_dst, scatter_desc, logits, grow_of(0), n, kNegInf2);
        out[1].init(scatter_dst, scatter_desc, logits, grow_of(1), n, kNegInf2);
        if (producer) {
            for (int j = 0; j < kPolicyTiles; ++j) {
                const uint32_t par = (uint32_t)((j / kSlots) & 1);
                mbar_wait(&bars[C_SLOT + slot_of(j)], par);
                if (leader) mbar_wait(&bars[C_PSLOT + slot_of(j)], par);
                else if (elect_one()) mbar_arrive_remote(&bars[C_PSLOT + slot_of(j)], 0);
                __syncwarp();
                if (leader) {
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        if (j >= 1) mbar_wait(&bars[C_DRAIN + t], (uint32_t)((j - 1) & 1));
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t db = umma_desc(smem_u32(base + slot_of(j) * kSlotBytes));
                            const uint32_t idesc = umma_idesc(kPolicyTile, F16, 256);
#pragma unroll
                            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma2_ts(tmem + kSkipCol + 128 * t, tmem + 128 * t + kb * 32 + k * 8,
                                             umma_desc_advance(db, kb * 8192 + k * 32), (uint32_t)((kb | k) != 0), idesc);
                            umma2_commit(&bars[C_ACC + t]);
                            if (t == 1) umma2_commit(&bars[C_RET + j % 3]);
                        }
                        __syncwarp();
                    }
                }
                if (j >= 1 && j + kSlots - 1 < kPolicyTiles) {        // tile j - 1 has retired in both accumulators: refill its slot
                    mbar_wait(&bars[C_RET + (j - 1) % 3], (uint32_t)(((j - 1) / 3) & 1));
                    if (elect_one()) load_tile(j + kSlots - 1);
                    __syncwarp();
                }
            }
        } else {
            const bool q3 = (warp & 3) == 3;
            for (int j = 0; j < kPolicyTiles; ++j) {
                const int n_my = (q3 && j % 3 == part) ? 2 : 1;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    mbar_wait(&bars[C_ACC + t], (uint32_t)(j & 1));
                    tc_fence_after();
                    YA_STAMP();
                    uint32_t acc[2][32];
                    tmem_ld32(t_skip(t) + part * 32, acc[0]);
                    if (n_my == 2) tmem_ld32(t_skip(t) + 96, acc[1]);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (leader) mbar_arrive(&bars[C_DRAIN + t]);
                        else mbar_arrive_remote(&bars[C_DRAIN + t], 0);
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if (q >= n_my) break;
                        const int col0 = j * kPolicyTile + (q ? 96 : part * 32);
                        if (col0 < kPolicyCols) policy_chunk<F16>(acc[q], col0, bias_all, out[t]);
                    }
                }
            }
        }
        YA_STAMP();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            __syncthreads();
            xchg[part * kRows + row].x = hmax16x2<F16>(out[t].row_mx);
            __syncthreads();
            const int64_t grow = grow_of(t);
            if (part == 0 && grow < n && row_max)
                row_max[grow] = fmaxf(fmaxf(xchg[0 * kRows + row].x, xchg[1 * kRows + row].x), fmaxf(xchg[2 * kRows + row].x, xchg[3 * kRows + row].x));
        }
    }
#ifdef YA_FWD_TIMELINE
    YA_CTA_TIME(2);
    if (blockIdx.x == 0 && tid == 0) {
        unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        g_timeline2[1002] = t_; g_timeline2[1003] = (unsigned long long)clock64();
    }
#endif
    tc_fence_before();
    cluster_sync();
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

// tiles_per_cta: 1 or 2; 0 = by size (two tiles per CTA as soon as one tile per SM cannot cover the leaves)
extern "C" int ya_nn_forward_tiles(const float* features, void* logits16, float* values, float* row_max, const void* weight_blob,
                                   const float* param_blob, const int64_t* offsets, int nblocks, int64_t n, float eps, int fp16,
                                   const uint64_t* scatter_dst, const uint32_t* scatter_desc, int tiles_per_cta, void* stream) {
    if (n <= 0) return 0;
    if ((reinterpret_cast<uintptr_t>(logits16) | reinterpret_cast<uintptr_t>(weight_blob) |
         reinterpret_cast<uintptr_t>(param_blob)) & 31u) return (int)cudaErrorMisalignedAddress;
    if ((scatter_dst == nullptr) != (scatter_desc == nullptr) || (!logits16 && !scatter_dst)) return (int)cudaErrorInvalidValue;
    if (tiles_per_cta < 0 || tiles_per_cta > 2) return (int)cudaErrorInvalidValue;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    // the opt-in to > 48 KB of dynamic shared memory is a per-DEVICE function attribute: one flag per device (set twice by
    // racing threads is harmless; the flag is only published after every call succeeded)
    static std::atomic<bool> configured[64];
    static std::atomic<int> sm_count[64];
    if (!configured[dev].load(std::memory_order_acquire)) {
        if ((e = cudaFuncSetAttribute(ya_k_forward<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)) != cudaSuccess) return (int)e;
        if ((e = cudaFuncSetAttribute(ya_k_forward<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)) != cudaSuccess) return (int)e;
        if ((e = cudaFuncSetAttribute(ya_k_forward2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)) != cudaSuccess) return (int)e;
        if ((e = cudaFuncSetAttribute(ya_k_forward2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)) != cudaSuccess) return (int)e;
        int sms = 0;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
        sm_count[dev].store(sms, std::memory_order_relaxed);
        configured[dev].store(true, std::memory_order_release);
    }
    if (tiles_per_cta == 0) tiles_per_cta = n > (int64_t)(sm_count[dev].load(std::memory_order_relaxed) & ~1) * kRows ? 2 : 1;
    Blob off{offsets[0], offsets[1], offsets[2], offsets[3], offsets[4], offsets[5], offsets[6], offsets[7], offsets[8]};
    const int64_t per_pair = 2 * kRows * tiles_per_cta;              // leaves per CTA pair
    const int blocks = 2 * (int)((n + per_pair - 1) / per_pair);
    auto launch = [&](auto kernel) {
        kernel<<<blocks, kThreads, kSmemBytes, (cudaStream_t)stream>>>(
            features, static_cast<uint16_t*>(logits16), values, row_max, static_cast<const uint8_t*>(weight_blob), param_blob, off,
            nblocks, n, eps, scatter_dst, scatter_desc, (reinterpret_cast<uintptr_t>(features) & 15u) == 0);
    };
    if (tiles_per_cta == 2) { if (fp16) launch(ya_k_forward2<true>); else launch(ya_k_forward2<false>); }
    else { if (fp16) launch(ya_k_forward<true>); else launch(ya_k_forward<false>); }
    return (int)cudaGetLastError();
}

extern "C" int ya_nn_forward(const float* features, void* logits16, float* values, float* row_max, const void* weight_blob,
                             const float* param_blob, const int64_t* offsets, int nblocks, int64_t n, float eps, int fp16,
                             const uint64_t* scatter_dst, const uint32_t* scatter_desc, void* stream) {
    return ya_nn_forward_tiles(features, logits16, values, row_max, weight_blob, param_blob, offsets, nblocks, n, eps, fp16,
                               scatter_dst, scatter_desc, 0, stream);
}
