// Yacht-Auction B200 engine -- fused normalisation/activation epilogues for the leaf evaluator.
//
// The dense contractions of the yacht NNet (yacht/pytorch/YachtNNet.py:8-70) stay cuBLASLt GEMMs
// issued by PyTorch; everything between two GEMMs (SiLU, LayerNorm, residual add -- three library
// kernels and three HBM round trips per layer in the stock forward, 60 % of its device time at
// 16,384 leaves) is one pass here: a warp owns a row of 256 bf16 activations (one 16-byte load per
// lane), statistics in float32 via warp shuffles, one 16-byte store.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include "../../include/yacht_b200.h"

namespace {

constexpr int kH = 256;            // hidden width handled by the fast path (main.py:40)
constexpr int kWarps = 8;

__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

struct Row8 { float v[8]; };

__device__ __forceinline__ Row8 load8(const __nv_bfloat16* p) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
    Row8 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y; }
    return r;
}

__device__ __forceinline__ void store8(__nv_bfloat16* p, const Row8& r) {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = raw;
}

__device__ __forceinline__ void layer_norm8(Row8& x, const Row8& gamma, const Row8& beta, float eps) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x.v[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    const float mean = s * (1.0f / kH);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float d = x.v[i] - mean; q += d * d; }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
    const float rstd = rsqrtf(q * (1.0f / kH) + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) x.v[i] = (x.v[i] - mean) * rstd * gamma.v[i] + beta.v[i];
}

// MODE 0: out = SiLU(LN(x))              (input layer and the two heads: Linear -> LN -> SiLU / LN -> SiLU -> Linear)
// MODE 1: out = LN(SiLU(x))              (first half of a residual block, YachtNNet.py:17-19)
// MODE 2: out = res + LN(SiLU(x))        (second half + skip connection, :20-21)
// MODE 3: out = SiLU(LN_a(x)), out2 = SiLU(LN_b(x))   (both heads read the trunk output once)
constexpr int kRowsPerWarp = 1;     // measured on B200: 1 row per warp (8.2 us per 16,384 rows) beats 2 or 4 rows in flight

template <int MODE>
__global__ void __launch_bounds__(kWarps * 32)
ya_k_ln_act(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
            const __nv_bfloat16* __restrict__ res, const __nv_bfloat16* __restrict__ gamma2,
            const __nv_bfloat16* __restrict__ beta2, __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out2,
            int64_t n, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    const Row8 g = load8(gamma + lane * 8), b = load8(beta + lane * 8);
    Row8 g2 = g, b2 = b;
    if (MODE == 3) { g2 = load8(gamma2 + lane * 8); b2 = load8(beta2 + lane * 8); }
    for (int64_t row0 = ((int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5)) * kRowsPerWarp; row0 < n;
         row0 += nwarps * kRowsPerWarp) {
        Row8 v[kRowsPerWarp], r[kRowsPerWarp];
#pragma unroll
        for (int k = 0; k < kRowsPerWarp; ++k) {
            const int64_t row = row0 + k;
            if (row < n) {
                v[k] = load8(x + row * kH + lane * 8);
                if (MODE == 2) r[k] = load8(res + row * kH + lane * 8);
            }
        }
#pragma unroll
        for (int k = 0; k < kRowsPerWarp; ++k) {
            const int64_t row = row0 + k;
            if (row >= n) break;                                   // warp-uniform
            if (MODE == 1 || MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[k].v[i] = silu(v[k].v[i]);
            }
            if (MODE == 3) {
                Row8 w = v[k];
                layer_norm8(w, g2, b2, eps);
#pragma unroll
                for (int i = 0; i < 8; ++i) w.v[i] = silu(w.v[i]);
                store8(out2 + row * kH + lane * 8, w);
            }
            layer_norm8(v[k], g, b, eps);
            if (MODE == 0 || MODE == 3) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[k].v[i] = silu(v[k].v[i]);
            }
            if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[k].v[i] += r[k].v[i];
            }
            store8(out + row * kH + lane * 8, v[k]);
        }
    }
}

}  // namespace

extern "C" int ya_nn_ln_act(int mode, const void* x, const void* gamma, const void* beta, const void* residual,
                            const void* gamma2, const void* beta2, void* out, void* out2, int64_t n, int64_t hidden,
                            float eps, void* stream) {
    if (n <= 0) return 0;
    if (hidden != kH) return (int)cudaErrorInvalidValue;
    auto X = static_cast<const __nv_bfloat16*>(x);
    auto G = static_cast<const __nv_bfloat16*>(gamma);
    auto B = static_cast<const __nv_bfloat16*>(beta);
    auto R = static_cast<const __nv_bfloat16*>(residual);
    auto G2 = static_cast<const __nv_bfloat16*>(gamma2);
    auto B2 = static_cast<const __nv_bfloat16*>(beta2);
    auto O = static_cast<__nv_bfloat16*>(out);
    auto O2 = static_cast<__nv_bfloat16*>(out2);
    int blocks = (int)((n + kWarps * kRowsPerWarp - 1) / (kWarps * kRowsPerWarp));
    if (blocks > 148 * 8) blocks = 148 * 8;
    cudaStream_t s = (cudaStream_t)stream;
    switch (mode) {
        case 0: ya_k_ln_act<0><<<blocks, kWarps * 32, 0, s>>>(X, G, B, R, G2, B2, O, O2, n, eps); break;
        case 1: ya_k_ln_act<1><<<blocks, kWarps * 32, 0, s>>>(X, G, B, R, G2, B2, O, O2, n, eps); break;
        case 2: ya_k_ln_act<2><<<blocks, kWarps * 32, 0, s>>>(X, G, B, R, G2, B2, O, O2, n, eps); break;
        case 3: ya_k_ln_act<3><<<blocks, kWarps * 32, 0, s>>>(X, G, B, R, G2, B2, O, O2, n, eps); break;
        default: return (int)cudaErrorInvalidValue;
    }
    return (int)cudaGetLastError();
}
