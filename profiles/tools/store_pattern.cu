// Write bandwidth of the policy head's store patterns: 16,384 rows x 6,528 B (12 x 272 16-bit slots) = 107 MB.
//  A: a warp instruction writes 32 B for each of 32 rows (thread = row; what a tcgen05.ld 32x32b epilogue gives)
//  B: 8 lanes x 32 B = 256 contiguous bytes for each of 4 rows
//  C: 32 lanes x 32 B = 1 KB contiguous of one row
// 128 CTAs x 512 threads, each CTA owns 128 rows, like the forward kernel.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int kRowBytes = 12 * 272 * 2;
__device__ __forceinline__ void st32(void* p, uint32_t v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(512) k(uint8_t* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* cta = out + (size_t)blockIdx.x * 128 * kRowBytes;
    const int sectors = kRowBytes / 32;              // 204 sectors per row
    if (MODE == 0) {                                 // warp w: rows 32 (w % 4) + lane, sectors w / 4, w / 4 + 4, ...
        uint8_t* row = cta + (size_t)((warp & 3) * 32 + lane) * kRowBytes;
        for (int s = 2 * (warp >> 2); s < sectors; s += 8) { st32(row + s * 32, s); st32(row + s * 32 + 32, s); }
    } else if (MODE == 1) {                          // 8 lanes per row
        for (int r = warp * 8; r < warp * 8 + 8; r += 4) {
            uint8_t* row = cta + (size_t)(r + (lane >> 3)) * kRowBytes;
            for (int s = lane & 7; s < sectors; s += 8) st32(row + s * 32, s);
        }
    } else {                                         // whole warp on one row
        for (int r = warp * 8; r < warp * 8 + 8; ++r) {
            uint8_t* row = cta + (size_t)r * kRowBytes;
            for (int s = lane; s < sectors; s += 32) st32(row + s * 32, s);
        }
    }
}
template <int MODE> void run(const char* name, uint8_t* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) k<MODE><<<128, 512>>>(out);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) k<MODE><<<128, 512>>>(out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double bytes = 128.0 * 128 * kRowBytes;
    printf("%-40s %7.1f us  %6.2f TB/s\n", name, ms * 50, bytes / (ms / 20 * 1e-3) * 1e-12);
}
int main() {
    uint8_t* out; cudaMalloc(&out, (size_t)128 * 128 * kRowBytes);
    run<0>("A: 32 rows x 32 B per instruction", out);
    run<1>("B: 4 rows x 256 B per instruction", out);
    run<2>("C: 1 row x 1 KB per instruction", out);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
