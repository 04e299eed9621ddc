// Microbenchmark: cost of LDS.128 on sm_100a as a function of how many DISTINCT 16-byte addresses the 32 lanes
// of a warp read (same-address lanes may be served by one broadcast wavefront).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds128 lds128.cu && ./lds128
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, int distinct, int stride16)
{
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // lane -> one of `distinct` addresses, stride16 16-byte units apart (7 = the skin kernel's 112-byte matrices)
    int idx = (distinct >= 32 ? lane : (lane % distinct)) * stride16;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float4 v = sm[(idx + u) & 4095];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx = (idx + 8 + (int)acc.x * 0) & 4095;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, sms * 4 * 1024 * sizeof(float));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int iters = 2048;
    for (int stride16 : {1, 7}) for (int distinct : {1, 2, 4, 8, 16, 32}) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<sms * 2, 1024, 65536>>>(out, iters, distinct, stride16);
        cudaEventRecord(e0);
        k<<<sms * 2, 1024, 65536>>>(out, iters, distinct, stride16);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double lds = (double)sms * 2 * 32 * iters * 8;      // warp-level LDS.128 instructions
        printf("stride %d x16B, %2d distinct addresses per warp: %.3f ms  -> %.2f cycles per LDS.128 per SM\n", stride16, distinct, ms,
               ms * 1e-3 * clk * 1e3 / (lds / sms));
    }
    return 0;
}
