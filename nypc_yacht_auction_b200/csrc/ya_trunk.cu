// Yacht-Auction B200 engine -- the residual trunk of the leaf evaluator as ONE persistent tcgen05 kernel.
//
// YachtNNet's trunk (yacht/pytorch/YachtNNet.py:8-21,64-66) is nblocks x [ h = LN1(SiLU(fc1(x))),
// x = x + LN2(SiLU(fc2(h))) ] on 256-wide rows: 12 GEMMs of [rows,256] x [256,256] whose outputs are
// immediately normalised row by row.  Run layer by layer through a library this costs a GEMM launch, an
// epilogue launch and two HBM round trips of the activations per layer.  Here a CTA owns 128 rows for the
// WHOLE trunk:
//   * activations (A operand, bf16, 128B-swizzled K-major) never leave shared memory,
//   * the skip connection lives in TMEM as float32 (columns 256..511),
//   * each layer is 16 tcgen05.mma (M=128, N=256, K=16; accumulator = TMEM columns 0..255) issued by one
//     thread, completion signalled through tcgen05.commit -> mbarrier,
//   * the next layer's 128 KB weight image (pre-swizzled on the host) streams in with cp.async.bulk
//     (complete_tx on an mbarrier) while the epilogue of the current layer runs,
//   * the epilogue reads the accumulator with tcgen05.ld (thread = row, so LayerNorm statistics need only
//     one exchange between the four threads that share a row), applies bias + SiLU + LayerNorm
//     (+ skip connection), and writes the bf16 result straight into the swizzled A tile of the next layer.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include "../../include/yacht_b200.h"
#include "ya_tc.cuh"

namespace {

using namespace ya_tc;

constexpr int kRows = 128;                 // rows per CTA = UMMA M
constexpr int kDim = 256;                  // hidden width = UMMA N = K
constexpr int kThreads = 512;              // 16 warps: warps w, w+4, w+8, w+12 share TMEM lanes 32*(w%4).., 64 columns each
constexpr int kParts = kThreads / 128;     // threads per row
constexpr int kCols = kDim / kParts;       // columns per thread
constexpr int kChunks = kCols / 32;        // 32-column TMEM accesses per thread and pass
constexpr int kABytes = kRows * kDim * 2;  // 64 KB
constexpr int kWBytes = kDim * kDim * 2;   // 128 KB
constexpr int kParamFloats = 3 * kDim;     // bias, gamma, beta
constexpr int kSmemBytes = 1024 + kABytes + kWBytes + 2 * kParamFloats * 4 + kRows * kParts * 8 + 64;

// byte offset of 8 consecutive bf16 (columns c8*8 .. c8*8+7) of row r inside the swizzled A tile
__device__ __forceinline__ uint32_t a_tile_offset(int r, int c8) {
    int kb = c8 >> 3, chunk = c8 & 7;
    return (uint32_t)(kb * (kRows * 128) + r * 128 + ((chunk ^ (r & 7)) << 4));
}

// kinds[l]: 1 -> out = LN(SiLU(z)) ; 2 -> out = skip + LN(SiLU(z)) and skip = out      (z = act x W_l^T + bias_l)
__global__ void __launch_bounds__(kThreads, 1)
ya_k_trunk(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, const uint8_t* __restrict__ wimg,
           const float* __restrict__ params, const int32_t* __restrict__ kinds, int layers, int64_t n, float eps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms need 1024-byte alignment
    uint8_t* a_tile = base;                                   // 64 KB, 1024-aligned
    uint8_t* w_tile = base + kABytes;                         // 128 KB, 1024-aligned
    float* prm_all = reinterpret_cast<float*>(w_tile + kWBytes);         // 2 x (bias | gamma | beta), double-buffered per layer
    float2* xchg = reinterpret_cast<float2*>(prm_all + 2 * kParamFloats);  // [kParts][128] partial (sum, sumsq)
    uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + kParts * kRows);  // [0] weights landed, [1] mma done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane;                   // TMEM lane = row of the tile
    const int half = warp >> 2;                               // which kCols-wide column slice this thread owns
    const int64_t grow = (int64_t)blockIdx.x * kRows + row;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
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
    const uint32_t t_acc = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * kCols);
    const uint32_t t_skip = t_acc + 256;

    // first weight image + its bias / gamma / beta (one transaction barrier for both)
    if (tid == 0) {
        mbar_expect_tx(&bars[0], kWBytes + kParamFloats * 4);
        for (int kb = 0; kb < 4; ++kb) bulk_g2s(w_tile + kb * 32768, wimg + kb * 32768, 32768, &bars[0]);
        bulk_g2s(prm_all, params, kParamFloats * 4, &bars[0]);
    }
    // prologue: my half row of the input -> swizzled A tile (bf16) and skip connection (float32, TMEM)
    {
        const uint4* src = reinterpret_cast<const uint4*>(x + grow * kDim + half * kCols);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {                    // chunks of 32 columns
            uint32_t f[32];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 v = grow < n ? src[c * 4 + q] : make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(a_tile + a_tile_offset(row, half * (kCols / 8) + c * 4 + q)) = v;
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) { f[q * 8 + 2 * t] = w[t] << 16; f[q * 8 + 2 * t + 1] = w[t] & 0xFFFF0000u; }
            }
            tmem_st32(t_skip + c * 32, f);
        }
        tmem_st_wait();
    }

    uint32_t w_phase = 0, m_phase = 0;
    for (int l = 0; l < layers; ++l) {
        const int kind = kinds[l];
        const float* prm = prm_all + (l & 1) * kParamFloats;  // double-buffered: layer l+1's copy lands in the other half
        proxy_fence();                                        // A tile written through the generic proxy
        tc_fence_before();
        __syncthreads();
        mbar_wait(&bars[0], w_phase);                         // weights + parameters of layer l have landed (all threads observe it)
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a0 = smem_u32(a_tile), b0 = smem_u32(w_tile);
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma(tmem, umma_desc(a0 + kb * (kRows * 128) + k * 32), umma_desc(b0 + kb * (kDim * 128) + k * 32),
                         (uint32_t)((kb | k) != 0), umma_idesc(kDim));
            umma_commit(&bars[1]);
        }
        w_phase ^= 1;
        mbar_wait(&bars[1], m_phase);                         // accumulator ready; A and W tiles free again
        m_phase ^= 1;
        tc_fence_after();
        if (tid == 0 && l + 1 < layers) {                     // next layer's weights stream in under the epilogue
            mbar_expect_tx(&bars[0], kWBytes + kParamFloats * 4);
            const uint8_t* src = wimg + (int64_t)(l + 1) * kWBytes;
            for (int kb = 0; kb < 4; ++kb) bulk_g2s(w_tile + kb * 32768, src + kb * 32768, 32768, &bars[0]);
            bulk_g2s(prm_all + ((l + 1) & 1) * kParamFloats, params + (int64_t)(l + 1) * kParamFloats, kParamFloats * 4, &bars[0]);
        }
        // ---- pass 1: z + bias -> SiLU, row statistics, values parked back in TMEM
        // (TMEM loads are software-pipelined: chunk c+1 is in flight while chunk c is processed)
        float s = 0.0f, ss = 0.0f;
        {
            float ps[4] = {0.0f, 0.0f, 0.0f, 0.0f}, pq[4] = {0.0f, 0.0f, 0.0f, 0.0f};   // 4 independent chains
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                uint32_t r[32];
                tmem_ld32(t_acc + c * 32, r);
                tmem_ld_wait();
                const float* bias = prm + half * kCols + c * 32;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float v = silu_from_half(fmaf(__uint_as_float(r[i]), 0.5f, 0.5f * bias[i]));
                    ps[i & 3] += v; pq[i & 3] = fmaf(v, v, pq[i & 3]);
                    r[i] = __float_as_uint(v);
                }
                tmem_st32(t_acc + c * 32, r);
            }
            s = (ps[0] + ps[1]) + (ps[2] + ps[3]);
            ss = (pq[0] + pq[1]) + (pq[2] + pq[3]);
        }
        tmem_st_wait();
        xchg[half * kRows + row] = make_float2(s, ss);
        __syncthreads();
#pragma unroll
        for (int p = 1; p < kParts; ++p) {
            float2 o = xchg[((half + p) % kParts) * kRows + row];
            s += o.x; ss += o.y;
        }
        const float mean = s * (1.0f / kDim);
        const float rstd = rsqrtf(fmaxf(ss * (1.0f / kDim) - mean * mean, 0.0f) + eps);
        // ---- pass 2: LayerNorm (+ skip), bf16 into the next layer's A tile (and to HBM after the last layer)
        const bool last = l + 1 == layers;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            uint32_t r[32], sk[32];
            tmem_ld32(t_acc + c * 32, r);
            if (kind == 2) tmem_ld32(t_skip + c * 32, sk);
            tmem_ld_wait();
            const float* gamma = prm + kDim + half * kCols + c * 32;
            const float* beta = prm + 2 * kDim + half * kCols + c * 32;
            uint32_t packed[16];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float ga = rstd * gamma[i];                      // (v - mean) * rstd * gamma + beta as two FMAs
                float v = fmaf(__uint_as_float(r[i]), ga, fmaf(-mean, ga, beta[i]));
                if (kind == 2) { v += __uint_as_float(sk[i]); sk[i] = __float_as_uint(v); }
                r[i] = __float_as_uint(v);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                __nv_bfloat162 p = __floats2bfloat162_rn(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
                packed[i] = *reinterpret_cast<uint32_t*>(&p);
            }
            if (kind == 2) tmem_st32(t_skip + c * 32, sk);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 v = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                *reinterpret_cast<uint4*>(a_tile + a_tile_offset(row, half * (kCols / 8) + c * 4 + q)) = v;
                if (last && grow < n) reinterpret_cast<uint4*>(out + grow * kDim + half * kCols)[c * 4 + q] = v;
            }
        }
        if (kind == 2) tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace

extern "C" int ya_nn_trunk(const void* x, void* out, const void* weight_images, const float* params, const int32_t* kinds,
                           int layers, int64_t n, int64_t hidden, float eps, void* stream) {
    if (n <= 0 || layers <= 0) return 0;
    if (hidden != kDim) return (int)cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(weight_images)) & 15u)
        return (int)cudaErrorMisalignedAddress;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(ya_k_trunk, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    int blocks = (int)((n + kRows - 1) / kRows);
    ya_k_trunk<<<blocks, kThreads, kSmemBytes, (cudaStream_t)stream>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), static_cast<const uint8_t*>(weight_images),
        params, kinds, layers, n, eps);
    return (int)cudaGetLastError();
}
