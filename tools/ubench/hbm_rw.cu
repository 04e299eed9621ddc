// hbm_rw.cu — what the HBM of this B200 sustains for the three stream shapes the OptimisedUtil kernels have:
// read-only, write-only and copy.  The skinning output of the fused kernel is a WRITE-ONLY stream (24 B per vertex out,
// nothing comparable in), so its ceiling is the write-only figure, not the copy figure MEASURED_PEAKS.json holds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_rw hbm_rw.cu && ./hbm_rw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_write(float4* __restrict__ d, size_t n, float v)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) d[i] = make_float4(v, v, v, v);
}
__global__ void k_write_cs(float4* __restrict__ d, size_t n, float v)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) __stcs(d + i, make_float4(v, v, v, v));
}
__global__ void k_read(const float4* __restrict__ s, size_t n, float* out)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { const float4 v = __ldcs(s + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) *out = acc;
}
__global__ void k_copy(const float4* __restrict__ s, float4* __restrict__ d, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) __stcs(d + i, __ldcs(s + i));
}
// one 6 KB TMA bulk store per 256-thread CTA iteration from shared memory (the shape skin_pair_kernel writes with)
__global__ void k_write_tma(uint8_t* __restrict__ d, size_t ntiles, float v)
{
    __shared__ __align__(128) float tile[2][1536];
    for (int i = threadIdx.x; i < 3072; i += blockDim.x) (&tile[0][0])[i] = v;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x, s ^= 1) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + t * 6144),
                         "r"((uint32_t)__cvta_generic_to_shared(&tile[s][0])), "r"(6144) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <typename F>
static float time_ms(F f, int iters = 10)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < iters; ++i) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    const size_t bytes = (size_t)2 << 30, n = bytes / 16;
    float4 *a, *b; float* o;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&o, 4);
    cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 8, block = 512;
    float t;
    t = time_ms([&] { k_write<<<grid, block>>>(a, n, 1.f); });       printf("write-only (st.global.v4)      %8.1f GB/s\n", bytes / t / 1e6);
    t = time_ms([&] { k_write_cs<<<grid, block>>>(a, n, 1.f); });    printf("write-only (st.global.cs.v4)   %8.1f GB/s\n", bytes / t / 1e6);
    t = time_ms([&] { k_write_tma<<<sms * 4, 256>>>((uint8_t*)a, bytes / 6144, 1.f); });
    printf("write-only (TMA bulk 6 KB)     %8.1f GB/s\n", (bytes / 6144) * 6144 / t / 1e6);
    t = time_ms([&] { cudaMemsetAsync(a, 1, bytes); });              printf("write-only (cudaMemset)        %8.1f GB/s\n", bytes / t / 1e6);
    t = time_ms([&] { k_read<<<grid, block>>>(a, n, o); });          printf("read-only  (ld.global.cs.v4)   %8.1f GB/s\n", bytes / t / 1e6);
    t = time_ms([&] { k_copy<<<grid, block>>>(a, b, n); });          printf("copy       (read + write)      %8.1f GB/s\n", 2.0 * bytes / t / 1e6);
    t = time_ms([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); });
    printf("copy       (cudaMemcpy D2D)    %8.1f GB/s\n", 2.0 * bytes / t / 1e6);
    return 0;
}
