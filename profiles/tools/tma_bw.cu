// Micro-benchmark: how fast can ONE SM (or a cluster pair) stream an L2-resident weight image into shared memory?
//   mode 0: cp.async.bulk (1-D), mode 1: cp.async.bulk.tensor.2d (tensor map, 64 x 256 boxes of 128-byte rows),
//   mode 2: 1-D bulk, cluster of 2, each CTA fetches half and multicasts it to both.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tma_bw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    for (int spin = 0; !ok; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (spin > (1 << 22)) __trap();                                // never hang the box
    }
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* b, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(b)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}

constexpr int kStage = 65536, kStages = 3;

template <int MODE>
__global__ void __launch_bounds__(128, 1) k_stream(const uint8_t* __restrict__ src, const __grid_constant__ CUtensorMap tmap,
                                                   int64_t total, int req, unsigned long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + kStages * kStage);
    uint32_t rank = 0;
    if (MODE == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&bars[i], 1); mbar_init(&bars[kStages + i], 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (MODE == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    if (threadIdx.x == 0) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        const int n = (int)(total / kStage);
        auto issue = [&](int s) {
            uint8_t* dst = base + (s % kStages) * kStage;
            uint64_t* bar = &bars[s % kStages];
            const uint8_t* g = src + ((int64_t)s * kStage) % (total);
            mbar_expect(bar, kStage);
            if (MODE == 0) {
                for (int o = 0; o < kStage; o += req)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + o)), "l"(g + o), "r"(req), "r"(smem_u32(bar)) : "memory");
            } else if (MODE == 1) {
                const int rows = req / 128;                            // box = 64 bf16 x rows
                for (int o = 0; o < kStage; o += req) {
                    int r0 = (int)((((int64_t)s * kStage) % total + o) / 128);
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(smem_u32(dst + o)), "l"(&tmap), "r"(0), "r"(r0), "r"(smem_u32(bar)) : "memory");
                }
                (void)rows;
            } else {
                const int half = kStage / 2;
                for (int o = 0; o < half; o += req)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                                 ::"r"(smem_u32(dst + rank * half + o)), "l"(g + rank * half + o), "r"(req), "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
            }
        };
        for (int s = 0; s < kStages - 1 && s < n; ++s) issue(s);
        for (int s = 0; s < n; ++s) {
            if (s + kStages - 1 < n) {
                const int nxt = s + kStages - 1;
                if (MODE == 2 && nxt >= kStages)                        // the slot must have been consumed by BOTH CTAs
                    mbar_wait(&bars[kStages + nxt % kStages], (uint32_t)((nxt / kStages - 1) & 1));
                issue(nxt);
            }
            mbar_wait(&bars[s % kStages], (uint32_t)((s / kStages) & 1));
            if (MODE == 2) { mbar_arrive_cluster(&bars[kStages + s % kStages], 0); mbar_arrive_cluster(&bars[kStages + s % kStages], 1); }
        }
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        out[blockIdx.x] = t1 - t0;
    }
    if (MODE == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int64_t total = 4 << 20;                                      // 4 MB image: L2 resident
    uint8_t* src; cudaMalloc(&src, total); cudaMemset(src, 1, total);
    unsigned long long* out; cudaMalloc(&out, 148 * 8);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    const int smem = kStages * kStage + 1024 + 64;
    cudaFuncSetAttribute(k_stream<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_stream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_stream<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode = 0; mode < 3; ++mode)
        for (int req : {8192, 16384, 32768})
            for (int ctas : {1, 2, 128}) {
                if (mode == 2 && ctas == 1) continue;
                CUtensorMap tm; memset(&tm, 0, sizeof(tm));
                if (mode == 1) {
                    cuuint64_t dims[2] = {64, (cuuint64_t)(total / 128)}; cuuint64_t strides[1] = {128};
                    cuuint32_t box[2] = {64, (cuuint32_t)(req / 128)}; cuuint32_t es[2] = {1, 1};
                    if (req / 128 > 256) continue;
                    CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, strides, box, es,
                                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
                }
                for (int rep = 0; rep < 3; ++rep) {
                    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
                    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = mode == 2 ? 2 : 1;
                    at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; cfg.attrs = at; cfg.numAttrs = 1;
                    cudaError_t e;
                    if (mode == 0) e = cudaLaunchKernelEx(&cfg, k_stream<0>, (const uint8_t*)src, tm, total, req, out);
                    else if (mode == 1) e = cudaLaunchKernelEx(&cfg, k_stream<1>, (const uint8_t*)src, tm, total, req, out);
                    else e = cudaLaunchKernelEx(&cfg, k_stream<2>, (const uint8_t*)src, tm, total, req, out);
                    e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("mode %d req %d ctas %d: %s\n", mode, req, ctas, cudaGetErrorString(e)); return 1; }
                }
                unsigned long long h[148]; cudaMemcpy(h, out, ctas * 8, cudaMemcpyDeviceToHost);
                unsigned long long mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
                printf("mode %d req %5d ctas %3d: %.1f us for %lld KB per CTA -> %.1f GB/s per SM\n", mode, req, ctas, mx / 1e3,
                       (long long)(total >> 10), total / (double)mx);
            }
    return 0;
}
