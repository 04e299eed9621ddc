// store_pattern.cu — how fast can a warp that holds TWO INSTANCES PER LANE (64 instances) push finished vertices
// (24 B each) to dst[instance][vertex] (instance stride S = nverts * 24)?  The output stage of skin_lane.cu.
// Every variant writes the same bytes (checked by a checksum of the destination), nothing else happens in the kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_pattern store_pattern.cu && ./store_pattern
// Variants:
//   0  tma2d  box 12 words x 32 rows, maps A/B over instance pairs (row stride 2S): two stores per vertex pair
//   1  tma2d  box 24 words x 32 rows (4 vertices)
//   2  tma2d  box 48 words x 32 rows (8 vertices)
//   3  staging [64 rows][96 B + 16] per warp, then cooperative 16-byte STG (six lanes per row)         (skin_wide.cu's way)
//   4  direct: every lane stores its two instances' 24 B with three STG.64 each
//   5  staging [64 rows][k x 24 B] per warp, one cp.async.bulk per row issued by lanes (2 rows per lane), k = 16 vertices
//   6  like 3 but rows of 8 vertices (192 B + 16): twelve lanes per row
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct Maps { CUtensorMap a, b, all; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int x, int y)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(x), "r"(y) : "memory");
}

constexpr int WARPS = 8;
constexpr int CHUNK = 128;   // vertices per CTA item

// value written for (instance, vertex, word): a cheap hash so that the checksum catches misplaced rows
__device__ __forceinline__ float val(uint32_t inst, uint32_t v, uint32_t w) { return __uint_as_float(0x3f800000u | ((inst * 131u + v * 7u + w) & 0x7fffffu)); }

template <int VAR>
__global__ void __launch_bounds__(256, 2) k_store(const __grid_constant__ Maps M, uint8_t* dst, uint32_t nverts, uint32_t ninst, uint32_t groups, uint32_t items)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t S = (size_t)nverts * 24;
    for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
        const uint32_t c = item / groups, g = item - c * groups;
        const uint32_t i0 = g * 64, vb = c * CHUNK + wid * (CHUNK / WARPS);          // this warp: 16 vertices
        const uint32_t ia = i0 + 2 * lane, ib = ia + 1;
        if (VAR <= 2) {
            constexpr int NV = VAR == 0 ? 2 : (VAR == 1 ? 4 : 8);
            uint8_t* st = smem + wid * (2 * 32 * NV * 24);
            for (uint32_t v = vb; v < vb + 16 && v < nverts; v += NV) {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                float* ra = reinterpret_cast<float*>(st + lane * NV * 24);
                float* rb = reinterpret_cast<float*>(st + 32 * NV * 24 + lane * NV * 24);
#pragma unroll
                for (int q = 0; q < NV * 6; q += 4) {
                    *reinterpret_cast<float4*>(ra + q) = make_float4(val(ia, v + q / 6, q % 6), val(ia, v + (q + 1) / 6, (q + 1) % 6), val(ia, v + (q + 2) / 6, (q + 2) % 6), val(ia, v + (q + 3) / 6, (q + 3) % 6));
                    *reinterpret_cast<float4*>(rb + q) = make_float4(val(ib, v + q / 6, q % 6), val(ib, v + (q + 1) / 6, (q + 1) % 6), val(ib, v + (q + 2) / 6, (q + 2) % 6), val(ib, v + (q + 3) / 6, (q + 3) % 6));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&M.a, smem_u32(st), v * 6, i0 / 2);
                    tma_store_2d(&M.b, smem_u32(st + 32 * NV * 24), v * 6, i0 / 2);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        } else if (VAR == 10 || VAR == 11) {
            constexpr int NV = VAR == 10 ? 4 : 2;
            uint8_t* st0 = smem + wid * (4 * 32 * NV * 24);
            uint32_t it = 0;
            for (uint32_t v = vb; v < vb + 16 && v < nverts; v += NV, ++it) {
                uint8_t* st = st0 + (it & 1) * (2 * 32 * NV * 24);
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                float* ra = reinterpret_cast<float*>(st + lane * NV * 24);
                float* rb = reinterpret_cast<float*>(st + 32 * NV * 24 + lane * NV * 24);
#pragma unroll
                for (int q = 0; q < NV * 6; q += 4) {
                    *reinterpret_cast<float4*>(ra + q) = make_float4(val(ia, v + q / 6, q % 6), val(ia, v + (q + 1) / 6, (q + 1) % 6), val(ia, v + (q + 2) / 6, (q + 2) % 6), val(ia, v + (q + 3) / 6, (q + 3) % 6));
                    *reinterpret_cast<float4*>(rb + q) = make_float4(val(ib, v + q / 6, q % 6), val(ib, v + (q + 1) / 6, (q + 1) % 6), val(ib, v + (q + 2) / 6, (q + 2) % 6), val(ib, v + (q + 3) / 6, (q + 3) % 6));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&M.a, smem_u32(st), v * 6, i0 / 2);
                    tma_store_2d(&M.b, smem_u32(st + 32 * NV * 24), v * 6, i0 / 2);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        } else if (VAR == 3 || VAR == 6) {
            constexpr int NV = VAR == 3 ? 4 : 8;
            constexpr int ROW = NV * 24 + 16;
            constexpr int CH = NV * 24 / 16;                                       // 16-byte pieces per row
            uint8_t* st = smem + wid * (64 * ROW);
            for (uint32_t v = vb; v < vb + 16 && v < nverts; v += NV) {
                __syncwarp();
                float* ra = reinterpret_cast<float*>(st + (2 * lane) * ROW);
                float* rb = reinterpret_cast<float*>(st + (2 * lane + 1) * ROW);
#pragma unroll
                for (int q = 0; q < NV * 6; q += 4) {
                    *reinterpret_cast<float4*>(ra + q) = make_float4(val(ia, v + q / 6, q % 6), val(ia, v + (q + 1) / 6, (q + 1) % 6), val(ia, v + (q + 2) / 6, (q + 2) % 6), val(ia, v + (q + 3) / 6, (q + 3) % 6));
                    *reinterpret_cast<float4*>(rb + q) = make_float4(val(ib, v + q / 6, q % 6), val(ib, v + (q + 1) / 6, (q + 1) % 6), val(ib, v + (q + 2) / 6, (q + 2) % 6), val(ib, v + (q + 3) / 6, (q + 3) % 6));
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 2 * CH; ++k) {
                    const uint32_t cc = lane + 32u * k, row = cc / CH, part = cc - row * CH;
                    if (i0 + row < ninst)
                        __stcs(reinterpret_cast<float4*>(dst + (size_t)(i0 + row) * S + (size_t)v * 24) + part, *reinterpret_cast<const float4*>(st + row * ROW + part * 16));
                }
            }
        } else if (VAR == 7 || VAR == 9) {
            // one map over individual instances (row stride S, needs S % 16 == 0): boxes of 64 rows; lane rows 2l, 2l+1
            constexpr int NV = VAR == 7 ? 4 : 2;
            uint8_t* st = smem + wid * (64 * NV * 24);
            for (uint32_t v = vb; v < vb + 16 && v < nverts; v += NV) {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                float* ra = reinterpret_cast<float*>(st + (2 * lane) * NV * 24);
                float* rb = reinterpret_cast<float*>(st + (2 * lane + 1) * NV * 24);
#pragma unroll
                for (int q = 0; q < NV * 6; q += 4) {
                    *reinterpret_cast<float4*>(ra + q) = make_float4(val(ia, v + q / 6, q % 6), val(ia, v + (q + 1) / 6, (q + 1) % 6), val(ia, v + (q + 2) / 6, (q + 2) % 6), val(ia, v + (q + 3) / 6, (q + 3) % 6));
                    *reinterpret_cast<float4*>(rb + q) = make_float4(val(ib, v + q / 6, q % 6), val(ib, v + (q + 1) / 6, (q + 1) % 6), val(ib, v + (q + 2) / 6, (q + 2) % 6), val(ib, v + (q + 3) / 6, (q + 3) % 6));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&M.all, smem_u32(st), v * 6, i0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        } else if (VAR == 8) {
            // hybrid: even instances through TMA (32 rows x 96 B), odd instances staged (112-byte pitch) and stored cooperatively
            uint8_t* sa = smem + wid * (32 * 96 + 32 * 112);
            uint8_t* sb = sa + 32 * 96;
            for (uint32_t v = vb; v < vb + 16 && v < nverts; v += 4) {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                float* ra = reinterpret_cast<float*>(sa + lane * 96);
                float* rb = reinterpret_cast<float*>(sb + lane * 112);
#pragma unroll
                for (int q = 0; q < 24; q += 4) {
                    *reinterpret_cast<float4*>(ra + q) = make_float4(val(ia, v + q / 6, q % 6), val(ia, v + (q + 1) / 6, (q + 1) % 6), val(ia, v + (q + 2) / 6, (q + 2) % 6), val(ia, v + (q + 3) / 6, (q + 3) % 6));
                    *reinterpret_cast<float4*>(rb + q) = make_float4(val(ib, v + q / 6, q % 6), val(ib, v + (q + 1) / 6, (q + 1) % 6), val(ib, v + (q + 2) / 6, (q + 2) % 6), val(ib, v + (q + 3) / 6, (q + 3) % 6));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&M.a, smem_u32(sa), v * 6, i0 / 2);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const uint32_t cc = lane + 32u * k, row = cc / 6, part = cc - row * 6;
                    if (i0 + 2 * row + 1 < ninst)
                        __stcs(reinterpret_cast<float4*>(dst + (size_t)(i0 + 2 * row + 1) * S + (size_t)v * 24) + part, *reinterpret_cast<const float4*>(sb + row * 112 + part * 16));
                }
            }
        } else if (VAR == 4) {
            for (uint32_t v = vb; v < vb + 16 && v < nverts; ++v) {
                float2* da = reinterpret_cast<float2*>(dst + (size_t)ia * S + (size_t)v * 24);
                float2* db = reinterpret_cast<float2*>(dst + (size_t)ib * S + (size_t)v * 24);
                if (ia < ninst) { da[0] = make_float2(val(ia, v, 0), val(ia, v, 1)); da[1] = make_float2(val(ia, v, 2), val(ia, v, 3)); da[2] = make_float2(val(ia, v, 4), val(ia, v, 5)); }
                if (ib < ninst) { db[0] = make_float2(val(ib, v, 0), val(ib, v, 1)); db[1] = make_float2(val(ib, v, 2), val(ib, v, 3)); db[2] = make_float2(val(ib, v, 4), val(ib, v, 5)); }
            }
        } else if (VAR == 5) {
            constexpr int ROW = 16 * 24 + 16;
            uint8_t* st = smem + wid * (64 * ROW);
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            for (uint32_t v = vb; v < vb + 16; v += 2) {
                float* ra = reinterpret_cast<float*>(st + (2 * lane) * ROW + (v - vb) * 24);
                float* rb = reinterpret_cast<float*>(st + (2 * lane + 1) * ROW + (v - vb) * 24);
#pragma unroll
                for (int q = 0; q < 12; q += 4) {
                    *reinterpret_cast<float4*>(ra + q) = make_float4(val(ia, v + q / 6, q % 6), val(ia, v + (q + 1) / 6, (q + 1) % 6), val(ia, v + (q + 2) / 6, (q + 2) % 6), val(ia, v + (q + 3) / 6, (q + 3) % 6));
                    *reinterpret_cast<float4*>(rb + q) = make_float4(val(ib, v + q / 6, q % 6), val(ib, v + (q + 1) / 6, (q + 1) % 6), val(ib, v + (q + 2) / 6, (q + 2) % 6), val(ib, v + (q + 3) / 6, (q + 3) % 6));
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            const uint32_t nb = min(16u, nverts - vb) * 24;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t row = 2 * lane + k;
                if (i0 + row < ninst)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)(i0 + row) * S + (size_t)vb * 24),
                                 "r"(smem_u32(st + row * ROW)), "r"(nb) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    if (VAR <= 2 || VAR == 5 || VAR >= 7) { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
}

__global__ void k_sum(const uint32_t* d, size_t n, unsigned long long* out)
{
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += d[i] * (unsigned long long)(i % 1000003 + 1);
    atomicAdd(out, acc);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int VAR>
static void run(const char* name, uint8_t* dst, uint32_t nverts, uint32_t ninst, EncodeTiledFn enc, size_t smem)
{
    const size_t S = (size_t)nverts * 24, total = S * ninst;
    Maps M;
    memset(&M, 0, sizeof M);
    if (VAR <= 2 || VAR == 10 || VAR == 11) {
        const cuuint32_t nv = VAR == 0 || VAR == 11 ? 2 : (VAR == 1 || VAR == 10 ? 4 : 8);
        const cuuint64_t dims[2] = {(cuuint64_t)nverts * 6, ninst / 2};
        const cuuint64_t strides[1] = {2 * S};
        const cuuint32_t box[2] = {nv * 6, 32}, es[2] = {1, 1};
        CUresult r1 = enc(&M.a, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, dst, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = enc(&M.b, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, dst + S, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 || r2) { printf("%s: encode failed %d %d\n", name, (int)r1, (int)r2); return; }
    }
    if (VAR == 8) {
        const cuuint64_t dims[2] = {(cuuint64_t)nverts * 6, ninst / 2};
        const cuuint64_t strides[1] = {2 * S};
        const cuuint32_t box[2] = {24, 32}, es[2] = {1, 1};
        if (enc(&M.a, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, dst, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode failed\n"); return; }
    }
    if (VAR == 7 || VAR == 9) {
        const cuuint64_t dims[2] = {(cuuint64_t)nverts * 6, ninst};
        const cuuint64_t strides[1] = {S};
        const cuuint32_t box[2] = {VAR == 7 ? 24u : 12u, 64}, es[2] = {1, 1};
        if (enc(&M.all, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, dst, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode failed\n"); return; }
    }
    const uint32_t groups = (ninst + 63) / 64, items = groups * ((nverts + CHUNK - 1) / CHUNK);
    CK(cudaFuncSetAttribute(k_store<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaMemset(dst, 0, total));
    unsigned long long* dsum; CK(cudaMalloc(&dsum, 8)); CK(cudaMemset(dsum, 0, 8));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(a);
        k_store<VAR><<<items, 256, smem>>>(M, dst, nverts, ninst, groups, items);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    k_sum<<<592, 256>>>((const uint32_t*)dst, total / 4, dsum);
    unsigned long long h; CK(cudaMemcpy(&h, dsum, 8, cudaMemcpyDeviceToHost));
    printf("%-34s %8.3f ms  %7.1f GB/s  %.2f cycles/vertex-instance/SM @1.965GHz  checksum %016llx\n", name, best, total / best / 1e6,
           best * 1e-3 * 1.965e9 * 148 / ((double)nverts * ninst), h);
    cudaFree(dsum);
}

int main(int argc, char** argv)
{
    const uint32_t nverts = argc > 1 ? atoi(argv[1]) : 1888, ninst = argc > 2 ? atoi(argv[2]) : 9984;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    uint8_t* dst; CK(cudaMalloc(&dst, (size_t)nverts * 24 * ninst + 4096));
    printf("nverts %u ninst %u (%.1f MB)\n", nverts, ninst, nverts * 24.0 * ninst / 1e6);
    run<0>("tma2d 2 vertices (48 B rows)", dst, nverts, ninst, enc, WARPS * 2 * 32 * 2 * 24);
    run<1>("tma2d 4 vertices (96 B rows)", dst, nverts, ninst, enc, WARPS * 2 * 32 * 4 * 24);
    run<2>("tma2d 8 vertices (192 B rows)", dst, nverts, ninst, enc, WARPS * 2 * 32 * 8 * 24);
    run<3>("staging 4 v + cooperative STG.128", dst, nverts, ninst, enc, WARPS * 64 * (4 * 24 + 16));
    run<6>("staging 8 v + cooperative STG.128", dst, nverts, ninst, enc, WARPS * 64 * (8 * 24 + 16));
    run<4>("direct 3 x STG.64 per instance", dst, nverts, ninst, enc, 0);
    run<5>("staging 16 v + per-row bulk copy", dst, nverts, ninst, enc, WARPS * 64 * (16 * 24 + 16));
    run<10>("tma2d 4 vertices, two buffers", dst, nverts, ninst, enc, WARPS * 4 * 32 * 4 * 24);
    run<11>("tma2d 2 vertices, two buffers", dst, nverts, ninst, enc, WARPS * 4 * 32 * 2 * 24);
    run<9>("tma2d 64-row boxes, 48 B rows", dst, nverts, ninst, enc, WARPS * 64 * 48);
    run<7>("tma2d 64-row boxes, 96 B rows", dst, nverts, ninst, enc, WARPS * 64 * 96);
    run<8>("hybrid: A tma 96 B, B cooperative", dst, nverts, ninst, enc, WARPS * (32 * 96 + 32 * 112));
    return 0;
}
