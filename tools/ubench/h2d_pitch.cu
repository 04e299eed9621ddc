// h2d_pitch.cu — can the copy engine skip the constant fourth row of every Affine3 on the way up?  Flat 64-byte matrices
// vs cudaMemcpy2DAsync of the first 48 bytes of each (pitch 64 on both sides), pinned host memory, 690 000 matrices
// (10 000 instances x 69 bones = the bench's per-step palette upload).
//   nvcc -O3 -o h2d_pitch h2d_pitch.cu && ./h2d_pitch
#include <cstdio>
#include <cuda_runtime.h>
int main()
{
    const size_t n = 690000;
    void *h, *d;
    cudaHostAlloc(&h, n * 64, cudaHostAllocDefault);
    cudaMalloc(&d, n * 64);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e30f;
        for (int it = 0; it < 12; ++it) {
            cudaEventRecord(a);
            if (mode == 0) cudaMemcpyAsync(d, h, n * 64, cudaMemcpyHostToDevice);
            else if (mode == 1) cudaMemcpy2DAsync(d, 64, h, 64, 48, n, cudaMemcpyHostToDevice);
            else cudaMemcpy2DAsync(d, 48, h, 64, 48, n, cudaMemcpyHostToDevice);
            cudaEventRecord(b); cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
            if (it >= 2 && ms < best) best = ms;
        }
        printf("%s: %.3f ms (%.1f GB/s of payload)\n", mode == 0 ? "flat 64 B x n            " : mode == 1 ? "2D 48 of 64 B, dst pitch 64" : "2D 48 of 64 B, dst packed  ",
               best, (mode == 0 ? n * 64 : n * 48) / best / 1e6);
    }
    return 0;
}
