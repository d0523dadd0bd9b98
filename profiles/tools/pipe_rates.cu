// Issue / pipe rates of the instructions the forward kernel's epilogue is built from, measured the way the epilogue runs
// them: 16 warps per SM (4 per scheduler), 8 independent chains per thread.  Prints warp-instructions per clock per SM
// and element operations per clock per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o _dbg/pipe_rates pipe_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 2048

template <int OP>
__global__ void __launch_bounds__(512, 1) k(float* out, float a, float b, long long* cycles) {
    float v[CHAINS * 2];
#pragma unroll
    for (int i = 0; i < CHAINS * 2; ++i) v[i] = a * (threadIdx.x + i);
    uint64_t aa, bb;
    asm("mov.b64 %0, {%1,%2};" : "=l"(aa) : "f"(a), "f"(a));
    asm("mov.b64 %0, {%1,%2};" : "=l"(bb) : "f"(b), "f"(b));
    uint32_t ah = 0x3c003c00u, bh = 0x2c002c00u;
    __shared__ float4 sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = make_float4(a, b, a, b);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) {            // FFMA
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[c]) : "f"(a), "f"(b));
            } else if (OP == 1) {     // FFMA2
                uint64_t p;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(p) : "f"(v[2 * c]), "f"(v[2 * c + 1]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(aa), "l"(bb));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(v[2 * c]), "=f"(v[2 * c + 1]) : "l"(p));
            } else if (OP == 2) {     // MUFU.TANH f32
                asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[c]));
            } else if (OP == 3) {     // tanh.approx.f16x2
                uint32_t& u = reinterpret_cast<uint32_t&>(v[c]);
                asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u));
            } else if (OP == 4) {     // HFMA2
                uint32_t& u = reinterpret_cast<uint32_t&>(v[c]);
                asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u) : "r"(ah), "r"(bh));
            } else if (OP == 5) {     // F2FP pack
                uint32_t u;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(v[2 * c]), "f"(v[2 * c + 1]));
                v[2 * c] = __uint_as_float(u | 0x3c000000u);
            } else if (OP == 6) {     // FFMA + FMNMX alternating (two pipes)
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[2 * c]) : "f"(a), "f"(b));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(v[2 * c + 1]) : "f"(a));
            } else if (OP == 7) {     // LDS.128 broadcast
                float4 q;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"((uint32_t)__cvta_generic_to_shared(&sh[(c + it) & 63])));
                v[c] += q.x;
            } else if (OP == 8) {     // FADD2
                uint64_t p;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(p) : "f"(v[2 * c]), "f"(v[2 * c + 1]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(aa));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(v[2 * c]), "=f"(v[2 * c + 1]) : "l"(p));
            } else if (OP == 9) {     // FFMA2 + MUFU.TANH mixed 1:1 (the SiLU mix)
                uint64_t p;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(p) : "f"(v[2 * c]), "f"(v[2 * c + 1]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(aa), "l"(bb));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(v[2 * c]), "=f"(v[2 * c + 1]) : "l"(p));
                asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[2 * c]));
            } else if (OP == 10) {    // ex2.approx.ftz.f32
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[c]));
            } else if (OP == 11) {    // HMNMX2
                uint32_t& u = reinterpret_cast<uint32_t&>(v[c]);
                asm volatile("max.f16x2 %0, %0, %1;" : "+r"(u) : "r"(ah));
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS * 2; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, int elems_per_instr, float* out, long long* cyc) {
    k<OP><<<148, 512>>>(out, 1.0001f, 0.0001f, cyc);
    cudaDeviceSynchronize();
    k<OP><<<148, 512>>>(out, 1.0001f, 0.0001f, cyc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double winstr = 16.0 * CHAINS * ITERS;      // warp instructions per SM (of the measured op)
    printf("%-28s %8.3f warp-instr/clk/SM  %8.1f elem-ops/clk/SM   (%lld cycles)\n", name, winstr / c, winstr * 32 * elems_per_instr / c, c);
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
    run<0>("FFMA", 1, out, cyc);
    run<1>("FFMA2", 2, out, cyc);
    run<8>("FADD2", 2, out, cyc);
    run<2>("MUFU.TANH f32", 1, out, cyc);
    run<10>("MUFU.EX2 f32", 1, out, cyc);
    run<3>("tanh.approx.f16x2", 2, out, cyc);
    run<4>("HFMA2", 2, out, cyc);
    run<11>("HMNMX2", 2, out, cyc);
    run<5>("F2FP.f16x2 (+LOP)", 2, out, cyc);
    run<6>("FFMA+FMNMX pair (per pair)", 2, out, cyc);
    run<7>("LDS.128 broadcast (+FADD)", 4, out, cyc);
    run<9>("FFMA2+MUFU.TANH (per pair)", 3, out, cyc);
    cudaError_t e = cudaGetLastError();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
