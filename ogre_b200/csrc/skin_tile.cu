// skin_tile.cu — tiled, instanced software skinning (and fused morph+skin) stream kernel.
//
// Reference work replaced: Entity::updateAnimation's per-entity
//   Mesh::softwareVertexMorph  -> OptimisedUtil::softwareVertexMorph   (OgreMain/src/OgreMesh.cpp:1899-1946)
//   Mesh::softwareVertexBlend  -> OptimisedUtil::softwareVertexSkinning (OgreMain/src/OgreMesh.cpp:1792-1897)
// for many instances of one mesh (BASELINE config 2).  Arithmetic: skin_math.cuh.
//
// Why a tile kernel: the one-thread-per-vertex kernel reads 24-byte vertices with scalar strided
// LDG/STG — ncu showed 5.6 of every 32 sector bytes used and the L1/TEX pipe at 87 % with DRAM at
// 39 % (profiles/r01a_skin_instanced_full.txt).  Here a CTA owns one 256-vertex tile of the mesh
// for a run of consecutive instances:
//   * the tile's blend indices/weights (shared by all instances) are read ONCE into registers;
//   * per instance the source tile (6 KB) arrives in shared memory by one TMA bulk copy
//     (cp.async.bulk, mbarrier completion) — or, in the fused form, is never read at all: the
//     two morph keyframes live in registers and the instance source is lerp(key1,key2,t[i]);
//   * the instance bone palette rows are gathered through bone_map with 16-byte cp.async;
//   * results are staged in shared memory and leave by one TMA bulk store per tile, so every
//     HBM transaction is a full line;
//   * source/palette of instance i+1 are in flight while instance i is computed (2 stages).
// When an address or byte count is not 16-byte aligned the same loop runs with cooperative,
// coalesced 8- or 4-byte LDG/STG instead of TMA (decided per tile, uniformly per CTA).
#include "common.cuh"
#include "skin_math.cuh"
#include "skin_tile.cuh"

using namespace ogrecu;

namespace ogrecu {

constexpr int ST_THREADS = 256;
#ifndef ST_MIN_CTAS
#define ST_MIN_CTAS 4
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void* gsrc, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* gdst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src_smem), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(gsrc) : "memory");
}

template <int NW, bool NORM, bool FUSED>
__global__ void __launch_bounds__(ST_THREADS, ST_MIN_CTAS) skin_tile_kernel(const __grid_constant__ SkinTileParams P)
{
    constexpr int VS = NORM ? 24 : 12;              // bytes per vertex
    constexpr int TILE_BYTES = ST_THREADS * VS;
    constexpr int MAXU4 = TILE_BYTES / (ST_THREADS * 4);   // 4-byte units per thread (6 or 3)

    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t npal = P.num_matrices * 3;       // float4 rows per palette
    uint8_t* s_dst = smem;                                          // [2][TILE_BYTES]
    uint8_t* s_src = s_dst + 2 * TILE_BYTES;                        // [2][TILE_BYTES]  (absent when FUSED)
    float4* s_pal = reinterpret_cast<float4*>(s_src + (FUSED ? 0 : 2 * TILE_BYTES));   // [2][npal]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_pal + 2 * npal);                   // [2]

    const uint32_t tid = threadIdx.x;
    const uint32_t tile = blockIdx.x;
    const uint32_t v0 = tile * ST_THREADS;
    const uint32_t nvalid = min((uint32_t)ST_THREADS, P.num_vertices - v0);
    const uint32_t bytes = nvalid * VS;
    const uint32_t i0 = blockIdx.y * P.inst_per_cta;
    const uint32_t cnt = min(P.inst_per_cta, P.num_instances - i0);
    const bool active = tid < nvalid;
    const size_t tile_off = (size_t)v0 * VS;
    const bool src_shared = !FUSED && P.src_inst_stride == 0;       // one bind pose for every instance
    const bool tma_ld = !FUSED && P.tma_load && (bytes & 15u) == 0;
    const bool tma_st = P.tma_store && (bytes & 15u) == 0;

    if (tma_ld && tid == 0) {
        mbar_init(smem_u32(&s_bar[0]), 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }

    // ---- per-vertex constants of this tile: blend indices + weights (shared by all instances)
    uint32_t idx4 = 0;
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    float k1[NORM ? 6 : 3], k2[NORM ? 6 : 3];
    if (active) {
        const uint8_t* bi = P.blend_index + (size_t)(v0 + tid) * P.blend_index_stride;
        const float* bw = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(P.blend_weight) + (size_t)(v0 + tid) * P.blend_weight_stride);
#pragma unroll
        for (int b = 0; b < NW; ++b) { idx4 |= (uint32_t)__ldg(bi + b) << (8 * b); w[b] = __ldg(bw + b); }
        if (FUSED) {
            const float* a = reinterpret_cast<const float*>(P.src + tile_off + (size_t)tid * VS);
            const float* b = reinterpret_cast<const float*>(P.key2 + tile_off + (size_t)tid * VS);
#pragma unroll
            for (int k = 0; k < (NORM ? 6 : 3); ++k) { k1[k] = __ldg(a + k); k2[k] = __ldg(b + k); }
        }
    }
    // palette gather plan: float4 piece p = tid + 256*q  ->  matrix p/3, row p%3
    uint32_t pal_off[3] = {0, 0, 0};   // float offset inside one instance palette
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        uint32_t p = tid + ST_THREADS * q;
        if (p < npal) {
            uint32_t m = p / 3, r = p - m * 3;
            uint32_t sm = P.bone_map ? (uint32_t)__ldg(P.bone_map + m) : m;
            pal_off[q] = sm * 16 + r * 4;
        }
    }
    __syncthreads();   // mbarrier init visible

    uint32_t pre[MAXU4];   // cooperative-load prefetch registers (viewed as uint2 when the unit is 8 bytes)

    auto issue_loads = [&](uint32_t stage, uint32_t inst, bool want_src) {
        const float* ipal = P.palette + (size_t)inst * P.palette_stride_floats;
        float4* dpal = s_pal + stage * npal;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            uint32_t p = tid + ST_THREADS * q;
            if (p < npal) cp_async16(smem_u32(dpal + p), ipal + pal_off[q]);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (FUSED || !want_src) return;
        const uint8_t* g = P.src + (size_t)inst * P.src_inst_stride + tile_off;
        if (tma_ld) {
            if (tid == 0) {
                uint32_t bar = smem_u32(&s_bar[stage]);
                mbar_expect_tx(bar, bytes);
                tma_load_1d(smem_u32(s_src + stage * TILE_BYTES), g, bytes, bar);
            }
        } else if (P.load_unit == 8) {
            const uint2* g8 = reinterpret_cast<const uint2*>(g);
            const uint32_t nu = bytes >> 3;
#pragma unroll
            for (int q = 0; q < MAXU4 / 2; ++q) {
                uint32_t u = tid + ST_THREADS * q;
                uint2 val = make_uint2(0, 0);
                if (u < nu) val = __ldcs(g8 + u);
                pre[2 * q] = val.x; pre[2 * q + 1] = val.y;
            }
        } else {
            const uint32_t* g4 = reinterpret_cast<const uint32_t*>(g);
            const uint32_t nu = bytes >> 2;
#pragma unroll
            for (int q = 0; q < MAXU4; ++q) {
                uint32_t u = tid + ST_THREADS * q;
                pre[q] = (u < nu) ? __ldcs(g4 + u) : 0u;
            }
        }
    };
    auto land_src = [&](uint32_t stage, uint32_t use) {   // make the stage's source tile readable
        if (FUSED) return;
        if (tma_ld) { mbar_wait(smem_u32(&s_bar[stage]), use & 1u); return; }
        uint8_t* d = s_src + stage * TILE_BYTES;
        if (P.load_unit == 8) {
            uint2* d8 = reinterpret_cast<uint2*>(d);
#pragma unroll
            for (int q = 0; q < MAXU4 / 2; ++q) d8[tid + ST_THREADS * q] = make_uint2(pre[2 * q], pre[2 * q + 1]);
        } else {
            uint32_t* d4 = reinterpret_cast<uint32_t*>(d);
#pragma unroll
            for (int q = 0; q < MAXU4; ++q) d4[tid + ST_THREADS * q] = pre[q];
        }
    };

    if (cnt == 0) return;
    issue_loads(0, i0, true);
    float t_next = FUSED ? __ldg(P.tarr + i0) : 0.f;

    for (uint32_t i = 0; i < cnt; ++i) {
        const uint32_t s = i & 1u;
        const uint32_t inst = i0 + i;
        const uint32_t ss = src_shared ? 0u : s;                   // source stage
        asm volatile("cp.async.wait_group 0;" ::: "memory");       // my palette pieces of stage s have landed
        if (!src_shared || i == 0) land_src(ss, src_shared ? 0u : (i >> 1));
        if (tma_st && tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // dst[s] of iteration i-2 drained
        __syncthreads();                                           // (A) stage s visible; stage s^1 free
        if (i + 1 < cnt) issue_loads(s ^ 1u, inst + 1, !src_shared);
        float t = 0.f;
        if (FUSED) { t = t_next; if (i + 1 < cnt) t_next = __ldg(P.tarr + inst + 1); }

        float o[NORM ? 6 : 3];
        if (active) {
            float x, y, z, nx = 0.f, ny = 0.f, nz = 0.f;
            if (FUSED) {
                x = morph1(k1[0], k2[0], t); y = morph1(k1[1], k2[1], t); z = morph1(k1[2], k2[2], t);
                if (NORM) {
                    nx = morph1(k1[3], k2[3], t); ny = morph1(k1[4], k2[4], t); nz = morph1(k1[5], k2[5], t);
                    normalise3(nx, ny, nz);
                }
                if (P.morph_out) {
                    float* mo = reinterpret_cast<float*>(P.morph_out + (size_t)inst * P.morph_inst_stride + tile_off + (size_t)tid * VS);
                    mo[0] = x; mo[1] = y; mo[2] = z;
                    if (NORM) { mo[3] = nx; mo[4] = ny; mo[5] = nz; }
                }
            } else {
                const uint8_t* sv = s_src + ss * TILE_BYTES + tid * VS;
                if (NORM) {
                    float2 a = reinterpret_cast<const float2*>(sv)[0], b = reinterpret_cast<const float2*>(sv)[1],
                           c = reinterpret_cast<const float2*>(sv)[2];
                    x = a.x; y = a.y; z = b.x; nx = b.y; ny = c.x; nz = c.y;
                } else {
                    const float* f = reinterpret_cast<const float*>(sv);
                    x = f[0]; y = f[1]; z = f[2];
                }
            }
            float4 r0, r1, r2;
            collapse_matrix<NW>(s_pal + s * npal, idx4, w, r0, r1, r2);
            xform_pos(r0, r1, r2, x, y, z, o[0], o[1], o[2]);
            if (NORM) xform_norm(r0, r1, r2, nx, ny, nz, o[3], o[4], o[5]);
            uint8_t* dv = s_dst + s * TILE_BYTES + tid * VS;
            if (NORM) {
                reinterpret_cast<float2*>(dv)[0] = make_float2(o[0], o[1]);
                reinterpret_cast<float2*>(dv)[1] = make_float2(o[2], o[3]);
                reinterpret_cast<float2*>(dv)[2] = make_float2(o[4], o[5]);
            } else {
                float* f = reinterpret_cast<float*>(dv);
                f[0] = o[0]; f[1] = o[1]; f[2] = o[2];
            }
        }
        if (tma_st) fence_async_smem();                            // my STS visible to the async proxy
        __syncthreads();                                           // (B) dst[s] complete
        // one bulk store per tile: per-warp 768-byte stores measured slower (profiles/r01c notes)
        uint8_t* gd = P.dst + (size_t)inst * P.dst_inst_stride + tile_off;
        if (tma_st) {
            if (tid == 0) tma_store_1d(gd, smem_u32(s_dst + s * TILE_BYTES), bytes);
        } else if (P.store_unit == 8) {
            const uint2* s8 = reinterpret_cast<const uint2*>(s_dst + s * TILE_BYTES);
            uint2* g8 = reinterpret_cast<uint2*>(gd);
            const uint32_t nu = bytes >> 3;
#pragma unroll
            for (int q = 0; q < MAXU4 / 2; ++q) { uint32_t u = tid + ST_THREADS * q; if (u < nu) __stcs(g8 + u, s8[u]); }
        } else {
            const uint32_t* s4 = reinterpret_cast<const uint32_t*>(s_dst + s * TILE_BYTES);
            uint32_t* g4 = reinterpret_cast<uint32_t*>(gd);
            const uint32_t nu = bytes >> 2;
#pragma unroll
            for (int q = 0; q < MAXU4; ++q) { uint32_t u = tid + ST_THREADS * q; if (u < nu) __stcs(g4 + u, s4[u]); }
        }
    }
    if (tma_st && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must outlive the reads
}

size_t skin_tile_smem(bool norm, bool fused, uint32_t num_matrices)
{
    const size_t tile = (size_t)ST_THREADS * (norm ? 24 : 12);
    return 2 * tile + (fused ? 0 : 2 * tile) + 2 * (size_t)num_matrices * 48 + 16;
}

template <int NW, bool NORM, bool FUSED>
static int launch_one(ogrecu_ctx* ctx, const SkinTileParams& P, dim3 grid, size_t smem)
{
    if (smem > 48 * 1024)   // only palettes of > 170 matrices in the unfused form (per device, so not cached)
        OGRECU_CUDA(cudaFuncSetAttribute(skin_tile_kernel<NW, NORM, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    skin_tile_kernel<NW, NORM, FUSED><<<grid, ST_THREADS, smem, ctx->stream>>>(P);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

template <bool NORM, bool FUSED>
static int launch_nw(ogrecu_ctx* ctx, const SkinTileParams& P, uint32_t nw, dim3 grid, size_t smem)
{
    switch (nw) {
    case 1: return launch_one<1, NORM, FUSED>(ctx, P, grid, smem);
    case 2: return launch_one<2, NORM, FUSED>(ctx, P, grid, smem);
    case 3: return launch_one<3, NORM, FUSED>(ctx, P, grid, smem);
    default: return launch_one<4, NORM, FUSED>(ctx, P, grid, smem);
    }
}

// Fills the derived fields (alignment modes, instances per CTA) and launches.
int launch_skin_tile(ogrecu_ctx* ctx, SkinTileParams P, uint32_t nw, bool norm, bool fused)
{
    const uint32_t vs = norm ? 24 : 12;
    const uint32_t tiles = div_up(P.num_vertices, ST_THREADS);
    auto al = [](const void* p, size_t stride, size_t a) { return (((uintptr_t)p | stride) & (a - 1)) == 0; };
    const int force = ctx->skin_mode;   // 0 auto, 1 never TMA (cooperative copies), 3 single-instance tile kernel; for A/B measurements
    P.tma_load = (!fused && force != 1 && al(P.src, P.src_inst_stride, 16)) ? 1u : 0u;
    P.tma_store = (force != 1 && al(P.dst, P.dst_inst_stride, 16)) ? 1u : 0u;
    P.load_unit = (vs == 24 && al(P.src, P.src_inst_stride, 8)) ? 8u : 4u;
    P.store_unit = (vs == 24 && al(P.dst, P.dst_inst_stride, 8)) ? 8u : 4u;
    // one instance, one long stream: persistent CTAs with a resident palette (skin_stream.cu), the tail below 256 vertices here
    if (force == 0 && skin_stream_eligible(P, nw, fused)) {
        uint32_t done = 0;
        OGRECU_CHECK(launch_skin_stream(ctx, P, nw, norm, &done));
        if (done == P.num_vertices) return OGRECU_OK;
        SkinTileParams T = P;
        T.src += (size_t)done * vs; T.dst += (size_t)done * vs;
        T.blend_index += (size_t)done * P.blend_index_stride;
        T.blend_weight = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(P.blend_weight) + (size_t)done * P.blend_weight_stride);
        T.num_vertices = P.num_vertices - done;
        return launch_skin_tile(ctx, T, nw, norm, fused);
    }
    // two or more instances: the pair kernel (skin_pair.cu, packed f32x2, two instances per pass)
    if (P.num_instances >= 2 && force != 3) return launch_skin_pair(ctx, P, nw, norm, fused);
    // instances per CTA: amortise the tile's blend stream while keeping >= ~4 waves of CTAs
    const uint64_t pairs = (uint64_t)tiles * P.num_instances;
    uint64_t ipc = pairs / ((uint64_t)ctx->sm_count * 24);
    if (ipc < 1) ipc = 1;
    if (ipc > 32) ipc = 32;
    if (ctx->skin_ipc > 0) ipc = (uint64_t)ctx->skin_ipc;
    if (ipc > P.num_instances) ipc = P.num_instances;
    P.inst_per_cta = (uint32_t)ipc;
    const uint32_t gy = div_up(P.num_instances, P.inst_per_cta);
    OGRECU_REQUIRE(gy <= 65535, "too many instance groups");
    dim3 grid(tiles, gy);
    const size_t smem = skin_tile_smem(norm, fused, P.num_matrices);
    if (norm) return fused ? launch_nw<true, true>(ctx, P, nw, grid, smem) : launch_nw<true, false>(ctx, P, nw, grid, smem);
    return fused ? launch_nw<false, true>(ctx, P, nw, grid, smem) : launch_nw<false, false>(ctx, P, nw, grid, smem);
}

}  // namespace ogrecu
