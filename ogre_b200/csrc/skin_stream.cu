// skin_stream.cu — software skinning of ONE long vertex stream (a single softwareVertexSkinning call over a big mesh):
// persistent CTAs, the palette resident in shared memory, a three-stage TMA pipeline over 256-vertex tiles.
//
// Reference work replaced: OptimisedUtil::softwareVertexSkinning, shared-packed (pos+normal, stride 24) and pos-only
// (stride 12) cases, OgreMain/src/OgreOptimisedUtilSSE.cpp:606-761,923-1035; arithmetic skin_math.cuh (bit-identical to SSE
// positions, exact 1/sqrt for normals).  This is the SURVEY §8d roofline case: 2^24 distinct vertices, 4 weights, 64 bones,
// 24 + 20 + 24 = 68 algorithmic bytes per vertex.
//
// Why a separate kernel: skin_tile_kernel pipelines over the INSTANCES of a tile.  With one instance a CTA loaded its blend
// indices (LDG), gathered the whole palette (3 KB for 64 bones, per 17 KB tile), fetched one source tile, computed, stored and
// exited — a serial latency chain per CTA hidden only by the four to seven CTAs an SM holds: 0.385 of HBM peak for four
// weights, 0.44 for one or two (profiles/r01q_microbench.txt).  Here
//   * a CTA gathers the palette ONCE and then walks tiles blockIdx.x, + gridDim.x, ... (grid = resident CTAs);
//   * the source tile and the tile's blend stream (UBYTE4 indices + n float weights interleaved, OGRE's layout,
//     OgreMesh.cpp:775-800) are both contiguous in memory, so each arrives by ONE cp.async.bulk with mbarrier completion,
//     issued two tiles ahead; the blend values are read from shared memory (OGRE's stride of 4 + 4n bytes is an odd number of
//     words for n = 2 and 4: conflict free) instead of by scattered LDG;
//   * results leave through a double-buffered shared tile by one bulk store per tile, as in skin_tile_kernel.
// The tail (num_vertices % 256) and anything that does not meet the layout / alignment conditions stay on skin_tile_kernel.
#include "common.cuh"
#include "skin_math.cuh"
#include "skin_tile.cuh"

using namespace ogrecu;

namespace ogrecu {

constexpr int SS_THREADS = 256;
#ifndef SS_DEFAULT_STAGES
#define SS_DEFAULT_STAGES 3
#endif

struct SkinStreamParams {
    const uint8_t* src;        // VS bytes per vertex, 16-byte aligned
    const uint8_t* blend;      // per vertex: 4 index bytes, then NW float weights; 16-byte aligned
    uint8_t* dst;              // VS bytes per vertex, 16-byte aligned
    const float* palette;      // Affine3, 16 floats each, 16-byte aligned
    const uint16_t* bone_map;  // optional
    uint32_t num_matrices, num_tiles;
    uint32_t blend_stride;     // bytes per vertex of the blend stream (>= 4 + 4 * NW, multiple of 4)
    uint32_t copy_shift;       // log2 of the palette copies in shared memory (3: one per 16-byte bank group, see s_pal)
};

__device__ __forceinline__ uint32_t ss_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ss_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}

// skin_math.cuh's collapse_matrix with the rows of matrix m at pal[(3 * m + r) << shift]: same operations, same order
// (NW == 1 ignores the weight, NW == 2 is lerp(M0, M1, w1), NW >= 3 is w0 * M0 then accumulate; OgreOptimisedUtilSSE.cpp:278-368)
template <int NW>
__device__ __forceinline__ void collapse_rows(const float4* __restrict__ pal, const uint32_t shift, const uint32_t idx4, const float* w,
                                              float4& r0, float4& r1, float4& r2)
{
    const uint32_t one = 1u << shift;
    const float4* m0 = pal + ((3u * (idx4 & 0xffu)) << shift);
    if (NW == 1) {
        r0 = m0[0]; r1 = m0[one]; r2 = m0[2 * one];
    } else if (NW == 2) {
        const float4* m1 = pal + ((3u * ((idx4 >> 8) & 0xffu)) << shift);
        r0 = lerp4(m0[0], m1[0], w[1]); r1 = lerp4(m0[one], m1[one], w[1]); r2 = lerp4(m0[2 * one], m1[2 * one], w[1]);
    } else {
        r0 = mul4(m0[0], w[0]); r1 = mul4(m0[one], w[0]); r2 = mul4(m0[2 * one], w[0]);
#pragma unroll
        for (int b = 1; b < NW; ++b) {
            const float4* mb = pal + ((3u * ((idx4 >> (8 * b)) & 0xffu)) << shift);
            r0 = madd4(mb[0], w[b], r0); r1 = madd4(mb[one], w[b], r1); r2 = madd4(mb[2 * one], w[b], r2);
        }
    }
}

template <int NW, bool NORM, int SS_STAGES>
__global__ void __launch_bounds__(SS_THREADS, 3) skin_stream_kernel(const __grid_constant__ SkinStreamParams P)
{
    constexpr int VS = NORM ? 24 : 12;              // bytes per vertex
    constexpr int TILE_BYTES = SS_THREADS * VS;
    const uint32_t BS = P.blend_stride;             // bytes per vertex of the blend stream
    const uint32_t BLEND_BYTES = SS_THREADS * BS;

    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t npal = P.num_matrices * 3;                                   // float4 rows
    uint8_t* s_dst = smem;                                                      // [2][TILE_BYTES]
    uint8_t* s_src = s_dst + 2 * TILE_BYTES;                                    // [STAGES][TILE_BYTES]
    uint8_t* s_bl = s_src + SS_STAGES * TILE_BYTES;                             // [STAGES][BLEND_BYTES]
    float4* s_pal = reinterpret_cast<float4*>(s_bl + SS_STAGES * BLEND_BYTES);  // [npal][copies]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_pal + ((size_t)npal << P.copy_shift));   // [STAGES]

    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SS_STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ss_smem_u32(&s_bar[s])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // The palette, once per CTA: rows 0..2 of every matrix, through bone_map (Mesh::prepareMatricesForVertexBlend,
    // OgreMesh.cpp:1782-1790), in 2^copy_shift interleaved COPIES: row p of copy q sits at float4 slot (p << copy_shift) + q, and
    // lane l reads copy l mod copies.  With eight copies every lane of a quarter-warp owns one 16-byte bank group, so an
    // LDS.128 of 32 unrelated bone rows takes its minimum of four wavefronts whatever the indices are — with one copy and the
    // uniformly random indices of the SURVEY §8d stream it took about ten: 0.591 of HBM peak with one copy, 0.604 / 0.670 / 0.967
    // with two / four / eight (profiles/r01q_microbench.txt).
    const uint32_t cs = P.copy_shift, ncopy = 1u << cs;
    for (uint32_t p = tid; p < npal; p += SS_THREADS) {
        const uint32_t m = p / 3, r = p - m * 3;
        const uint32_t sm = P.bone_map ? (uint32_t)__ldg(P.bone_map + m) : m;
        const float4 row = __ldg(reinterpret_cast<const float4*>(P.palette + (size_t)sm * 16 + r * 4));
        for (uint32_t q = 0; q < ncopy; ++q) s_pal[(p << cs) + q] = row;
    }
    __syncthreads();                                                            // barriers initialised, palette visible
    const float4* my_pal = s_pal + (tid & (ncopy - 1u));

    const uint32_t first = blockIdx.x, step = gridDim.x;
    const uint32_t ntiles = P.num_tiles;
    const uint32_t mine = first < ntiles ? (ntiles - first + step - 1) / step : 0;   // tiles this CTA walks

    auto issue = [&](uint32_t k) {                                              // tile number k of this CTA -> stage k % STAGES
        const uint32_t s = k % SS_STAGES;
        const size_t tile = (size_t)first + (size_t)k * step;
        const uint32_t bar = ss_smem_u32(&s_bar[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TILE_BYTES + BLEND_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(ss_smem_u32(s_src + s * TILE_BYTES)), "l"(P.src + tile * TILE_BYTES), "r"(TILE_BYTES), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(ss_smem_u32(s_bl + s * BLEND_BYTES)), "l"(P.blend + tile * BLEND_BYTES), "r"(BLEND_BYTES), "r"(bar) : "memory");
    };
    if (tid == 0)
        for (uint32_t k = 0; k < SS_STAGES - 1 && k < mine; ++k) issue(k);

    for (uint32_t k = 0; k < mine; ++k) {
        const uint32_t s = k % SS_STAGES, d = k & 1u;
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // s_dst[d] of tile k-2 has been read out
        ss_mbar_wait(ss_smem_u32(&s_bar[s]), (k / SS_STAGES) & 1u);               // source + blend of tile k have landed
        __syncthreads();                                                        // (A) everyone is past tile k-1: its stage and s_dst[d] are free
        if (tid == 0 && k + SS_STAGES - 1 < mine) issue(k + SS_STAGES - 1);     // refill the stage tile k-1 used

        const uint8_t* sv = s_src + s * TILE_BYTES + tid * VS;
        const uint8_t* bl = s_bl + s * BLEND_BYTES + tid * BS;
        const uint32_t idx4 = *reinterpret_cast<const uint32_t*>(bl);
        float w[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int b = 0; b < NW; ++b) w[b] = *reinterpret_cast<const float*>(bl + 4 + 4 * b);
        float x, y, z, nx = 0.f, ny = 0.f, nz = 0.f;
        if (NORM) {
            const float2 a = reinterpret_cast<const float2*>(sv)[0], b = reinterpret_cast<const float2*>(sv)[1],
                         c = reinterpret_cast<const float2*>(sv)[2];
            x = a.x; y = a.y; z = b.x; nx = b.y; ny = c.x; nz = c.y;
        } else {
            const float* f = reinterpret_cast<const float*>(sv);
            x = f[0]; y = f[1]; z = f[2];
        }
        float4 r0, r1, r2;
        collapse_rows<NW>(my_pal, cs, idx4, w, r0, r1, r2);
        float o[NORM ? 6 : 3];
        xform_pos(r0, r1, r2, x, y, z, o[0], o[1], o[2]);
        if (NORM) xform_norm(r0, r1, r2, nx, ny, nz, o[3], o[4], o[5]);
        uint8_t* dv = s_dst + d * TILE_BYTES + tid * VS;
        if (NORM) {
            reinterpret_cast<float2*>(dv)[0] = make_float2(o[0], o[1]);
            reinterpret_cast<float2*>(dv)[1] = make_float2(o[2], o[3]);
            reinterpret_cast<float2*>(dv)[2] = make_float2(o[4], o[5]);
        } else {
            float* f = reinterpret_cast<float*>(dv);
            f[0] = o[0]; f[1] = o[1]; f[2] = o[2];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");             // my STS visible to the async proxy
        __syncthreads();                                                        // (B) s_dst[d] complete
        if (tid == 0) {
            const size_t tile = (size_t)first + (size_t)k * step;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(P.dst + tile * TILE_BYTES), "r"(ss_smem_u32(s_dst + d * TILE_BYTES)), "r"(TILE_BYTES) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the reads
}

static size_t skin_stream_smem(uint32_t blend_stride, bool norm, uint32_t num_matrices, uint32_t copy_shift, int stages)
{
    const size_t tile = (size_t)SS_THREADS * (norm ? 24 : 12), blend = (size_t)SS_THREADS * blend_stride;
    return 2 * tile + stages * (tile + blend) + (((size_t)num_matrices * 48) << copy_shift) + stages * 8;
}

template <int NW, bool NORM, int STAGES>
static int launch_stream_st(ogrecu_ctx* ctx, const SkinStreamParams& P, bool norm)
{
    const size_t smem = skin_stream_smem(P.blend_stride, norm, P.num_matrices, P.copy_shift, STAGES);
    auto kern = skin_stream_kernel<NW, NORM, STAGES>;
    if (smem > 48 * 1024) OGRECU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    OGRECU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SS_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const uint32_t grid = std::min<uint32_t>(P.num_tiles, (uint32_t)(per_sm * ctx->sm_count));
    kern<<<grid, SS_THREADS, smem, ctx->stream>>>(P);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

template <int NW, bool NORM>
static int launch_stream_one(ogrecu_ctx* ctx, const SkinStreamParams& P, bool norm)
{
    static const int stages = getenv("OGRECU_STREAM_STAGES") ? atoi(getenv("OGRECU_STREAM_STAGES")) : SS_DEFAULT_STAGES;   // A/B
    if (stages == 2) return launch_stream_st<NW, NORM, 2>(ctx, P, norm);
    if (stages == 4) return launch_stream_st<NW, NORM, 4>(ctx, P, norm);
    return launch_stream_st<NW, NORM, 3>(ctx, P, norm);
}

// One instance, dense layout, interleaved blend stream, everything 16-byte aligned, at least 8 full tiles.
bool skin_stream_eligible(const SkinTileParams& P, uint32_t nw, bool fused)
{
    if (fused || P.num_instances != 1 || P.num_vertices < 8u * SS_THREADS || P.num_matrices > 256) return false;
    if (reinterpret_cast<const uint8_t*>(P.blend_weight) != P.blend_index + 4) return false;
    if (P.blend_weight_stride != P.blend_index_stride || P.blend_index_stride < 4 + 4 * nw || (P.blend_index_stride & 3u) ||
        P.blend_index_stride > 64) return false;
    return (((uintptr_t)P.src | (uintptr_t)P.dst | (uintptr_t)P.blend_index | (uintptr_t)P.palette) & 15u) == 0;
}

// Skins the full 256-vertex tiles of the stream; returns how many vertices were covered in *done.
int launch_skin_stream(ogrecu_ctx* ctx, const SkinTileParams& T, uint32_t nw, bool norm, uint32_t* done)
{
    SkinStreamParams P;
    P.src = T.src; P.blend = T.blend_index; P.dst = T.dst; P.palette = T.palette; P.bone_map = T.bone_map;
    P.num_matrices = T.num_matrices; P.num_tiles = T.num_vertices / SS_THREADS;
    P.blend_stride = T.blend_index_stride;
    // palette copies: eight (one per bank group) while they fit in 24 KB, then four, two, one
    static const int forced = getenv("OGRECU_STREAM_COPIES") ? atoi(getenv("OGRECU_STREAM_COPIES")) : -1;   // A/B: log2 of the copies
    P.copy_shift = T.num_matrices <= 64 ? 3u : T.num_matrices <= 128 ? 2u : 1u;
    if (nw <= 2 && P.copy_shift > 2) P.copy_shift = 2;   // few row loads per vertex: four copies and one more CTA per SM measured best
    if (forced >= 0 && forced <= 3) P.copy_shift = (uint32_t)forced;
    *done = P.num_tiles * SS_THREADS;
    switch (nw * 2 + (norm ? 1 : 0)) {
    case 2: return launch_stream_one<1, false>(ctx, P, norm);
    case 3: return launch_stream_one<1, true>(ctx, P, norm);
    case 4: return launch_stream_one<2, false>(ctx, P, norm);
    case 5: return launch_stream_one<2, true>(ctx, P, norm);
    case 6: return launch_stream_one<3, false>(ctx, P, norm);
    case 7: return launch_stream_one<3, true>(ctx, P, norm);
    case 8: return launch_stream_one<4, false>(ctx, P, norm);
    default: return launch_stream_one<4, true>(ctx, P, norm);
    }
}

}  // namespace ogrecu
