// exact_rsq.cu — exhaustive check of the branch-free IEEE sqrt / reciprocal sequences used by skin_pair.cu
// (ogre_b200/csrc/exact_math.cuh) against __fsqrt_rn / __frcp_rn over EVERY float32 bit pattern of their fast ranges.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../ogre_b200/csrc -o exact_rsq exact_rsq.cu && ./exact_rsq
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "exact_math.cuh"

__global__ void check(unsigned long long* bad_sqrt, unsigned long long* bad_rcp, unsigned long long* n_sqrt, unsigned long long* n_rcp,
                      uint32_t* first_bad)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long bs = 0, br = 0, ns = 0, nr = 0;
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < (1ull << 32); u += stride) {
        const uint32_t b = (uint32_t)u;
        const float x = __uint_as_float(b);
        if (ogrecu::em_fast_range(b)) {
            ++ns;
            const float s = ogrecu::em_sqrt_fast(x);
            if (__float_as_uint(s) != __float_as_uint(__fsqrt_rn(x))) { ++bs; atomicMin(&first_bad[0], b); }
            ++nr;
            const float r = ogrecu::em_rcp_fast(x);
            if (__float_as_uint(r) != __float_as_uint(__frcp_rn(x))) { ++br; atomicMin(&first_bad[1], b); }
        }
    }
    atomicAdd(bad_sqrt, bs); atomicAdd(bad_rcp, br); atomicAdd(n_sqrt, ns); atomicAdd(n_rcp, nr);
}

int main()
{
    unsigned long long* c; uint32_t* fb;
    cudaMallocManaged(&c, 4 * sizeof(unsigned long long)); cudaMallocManaged(&fb, 8);
    c[0] = c[1] = c[2] = c[3] = 0; fb[0] = fb[1] = 0xffffffffu;
    check<<<148 * 16, 256>>>(c, c + 1, c + 2, c + 3, fb);
    cudaDeviceSynchronize();
    printf("sqrt: %llu inputs in the fast range, %llu differ from __fsqrt_rn (first 0x%08x)\n", c[2], c[0], fb[0]);
    printf("rcp : %llu inputs in the fast range, %llu differ from __frcp_rn (first 0x%08x)\n", c[3], c[1], fb[1]);
    return (c[0] || c[1]) ? 1 : 0;
}
