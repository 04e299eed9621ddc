// Microbenchmark: throughput of un-fused multiply+add on sm_100a, scalar vs packed f32x2.
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad false, so the packed
// un-fused form is written as fma.rn.f32x2(a, b, nz) (nz = -0.0 from a kernel argument, opaque to
// ptxas: one rounding, identical to mul) followed by add.rn.f32x2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o fp32x2 fp32x2.cu && ./fp32x2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { return ((u64)__float_as_uint(hi) << 32) | __float_as_uint(lo); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

#define CH 8
// mode 0: scalar FMUL + FADD (2*CH*2 flop-instr per iter per "pair" -> counts as 4*CH lane ops)
template <int MODE>
__global__ void k(float* out, int iters, float sa, float sb, u64 nz, u64 one)
{
    float a[2 * CH], acc[2 * CH];
    for (int i = 0; i < 2 * CH; ++i) { a[i] = sa + i * 0.001f + threadIdx.x * 1e-6f; acc[i] = sb + i; }
    u64 pa[CH], pacc[CH];
    for (int i = 0; i < CH; ++i) { pa[i] = pk(a[2 * i], a[2 * i + 1]); pacc[i] = pk(acc[2 * i], acc[2 * i + 1]); }
    u64 pm = pk(sa, sa);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 2 * CH; ++i) acc[i] = __fadd_rn(__fmul_rn(acc[i], sa), a[i]);
        } else if (MODE == 1) {       // FFMA2(acc, m, -0) ; FADD2
#pragma unroll
            for (int i = 0; i < CH; ++i) pacc[i] = add2(fma2(pacc[i], pm, nz), pa[i]);
        } else if (MODE == 2) {       // FMUL2 ; FFMA2(x, 1, a)
#pragma unroll
            for (int i = 0; i < CH; ++i) pacc[i] = fma2(mul2(pacc[i], pm), one, pa[i]);
        } else if (MODE == 3) {       // fused FFMA2 (reference peak)
#pragma unroll
            for (int i = 0; i < CH; ++i) pacc[i] = fma2(pacc[i], pm, pa[i]);
        } else if (MODE == 4) {       // scalar fused FFMA (reference)
#pragma unroll
            for (int i = 0; i < 2 * CH; ++i) acc[i] = __fmaf_rn(acc[i], sa, a[i]);
        } else if (MODE == 5) {       // FMUL2 only + FADD2 only, independent (no fusion possible): mul chain and add chain
#pragma unroll
            for (int i = 0; i < CH; i += 2) { pacc[i] = mul2(pacc[i], pm); pacc[i + 1] = add2(pacc[i + 1], pa[i]); }
        }
    }
    float s = 0;
    for (int i = 0; i < 2 * CH; ++i) s += acc[i];
    for (int i = 0; i < CH; ++i) s += __uint_as_float((unsigned)pacc[i]) + __uint_as_float((unsigned)(pacc[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, double lane_ops_per_iter_per_thread)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    int blocks = sms * 8, threads = 256, iters = 4096;
    float* out; cudaMalloc(&out, blocks * threads * sizeof(float));
    u64 nz = 0x8000000080000000ull, one = 0x3f8000003f800000ull;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.5f, nz, one);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.5f, nz, one);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * lane_ops_per_iter_per_thread;
    printf("%-40s %8.3f ms  %7.2f T lane-ops/s  = %6.1f lane-ops/clk/SM (at %d MHz nominal)\n", name, ms, ops / ms / 1e9,
           ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    cudaFree(out);
}

int main()
{
    run<0>("scalar FMUL+FADD (unfused)", 4.0 * CH);
    run<1>("FFMA2(a,b,-0)+FADD2 (unfused, packed)", 4.0 * CH);
    run<2>("FMUL2+FFMA2(x,1,y) (unfused, packed)", 4.0 * CH);
    run<3>("FFMA2 fused", 4.0 * CH);
    run<4>("FFMA scalar fused", 4.0 * CH);
    run<5>("FMUL2 / FADD2 independent", 1.0 * CH);
    return 0;
}
