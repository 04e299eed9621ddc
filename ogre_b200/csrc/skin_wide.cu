// skin_wide.cu — instanced software skinning (and fused morph+skin) with the INSTANCES across the lanes of a warp.
//
// Reference work replaced: Mesh::softwareVertexMorph -> OptimisedUtil::softwareVertexMorph and
// Mesh::softwareVertexBlend -> OptimisedUtil::softwareVertexSkinning (OgreMain/src/OgreMesh.cpp:1792-1946,
// arithmetic OgreOptimisedUtilSSE.cpp:278-421,606-761 and OgreSIMDHelper.h:204-227) for many instances of one mesh.
//
// Why another layout.  skin_pair.cu gives a thread one vertex and walks the instances; every lane of a warp then
// addresses different bone matrices, and ncu (profiles/r01f_*) shows what that costs: the shared-memory data pipe is
// the busiest unit (a 4-weight vertex reads 192 B of palette per instance, 21 % more in bank conflicts), 27 % of the
// instructions are per-lane address and control overhead, and the CTA's two barriers per pass lock-step the warps.
// Here a lane owns one INSTANCE and a warp walks the vertices:
//   * the bone indices and weights of a vertex are warp-uniform: one broadcast load, no per-lane address arithmetic;
//   * the palettes of the CTA's 32 instances sit in shared memory as [matrix][row][instance]: every LDS.128 of a warp
//     reads 512 contiguous bytes, conflict free;
//   * a zero blend weight is zero for the whole warp, so its slot is skipped by a uniform branch — 45 % of the slots
//     of Sinbad's body.  SSE multiplies through where General skips (`if (weight)`, OgreOptimisedUtilGeneral.cpp:135);
//     the skip is bitwise exact unless the accumulated element is -0 (adding +0 would turn it into +0), which needs a -0
//     (or a negative denormal) in the slot-0 matrix, or unless the skipped matrix is not finite (SSE: 0 * inf = NaN).
//     Both are properties of the palettes: they are checked once while the palettes are copied into shared memory
//     (per matrix: some instance holds a -0; per CTA: something not finite), and such vertices take every slot;
//   * the palettes are gathered once per CTA for a few hundred vertices, not once per 128-vertex tile;
//   * warps only synchronise with themselves (the transposing output stage is per warp): no CTA barrier in the loop,
//     so the LDS-heavy and FMA-heavy phases of different warps overlap.
// Results are staged per warp as [instance][4 vertices] and leave as 16-byte stores, six consecutive lanes per
// instance row.  Arithmetic: the scalar functions of skin_math.cuh, operation for operation those of skin_pair.cu.
#include "common.cuh"
#include "skin_math.cuh"
#include "skin_tile.cuh"

#include <algorithm>

using namespace ogrecu;

namespace ogrecu {

constexpr int SW_WARPS = 16;
constexpr int SW_THREADS = SW_WARPS * 32;
constexpr int SW_GROUP = 32;                  // instances per CTA = lanes
constexpr int SW_VB = 4;                      // vertices staged per warp and round
constexpr int SW_ROW = SW_VB * 24 + 16;       // bytes per instance row of the staging (7 x 16: odd, conflict-free STS.128)
constexpr int SW_VTX = 80;                    // bytes of warp-uniform data per vertex (5 x float4)

__device__ __forceinline__ bool sw_neg_zero_or_denormal(uint32_t u) { return (u - 0x80000000u) < 0x00800000u; }
__device__ __forceinline__ bool sw_huge(uint32_t u) { return (u & 0x7fffffffu) >= 0x7f000000u; }   // >= 2^127, inf, NaN

template <int NW, bool FUSED>
__global__ void __launch_bounds__(SW_THREADS, 1) skin_wide_kernel(const __grid_constant__ SkinTileParams P, const uint32_t verts_per_cta)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t nm = P.num_matrices;
    float4* s_pal = reinterpret_cast<float4*>(smem);                                  // [nm][3][32]
    float4* s_vtx = reinterpret_cast<float4*>(smem + (size_t)nm * 1536);              // [verts_per_cta][5]
    uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_vtx) + (size_t)verts_per_cta * SW_VTX;   // [SW_WARPS][32][SW_ROW]
    uint32_t* s_dirty = reinterpret_cast<uint32_t*>(s_stage + SW_WARPS * SW_GROUP * SW_ROW);  // [nm] + [1]

    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t vbeg = blockIdx.x * verts_per_cta;
    const uint32_t vcount = min(verts_per_cta, P.num_vertices - vbeg);
    const uint32_t i0 = blockIdx.y * SW_GROUP;
    const bool live = i0 + lane < P.num_instances;

    for (uint32_t k = tid; k <= nm; k += SW_THREADS) s_dirty[k] = 0;
    __syncthreads();

    // ---- palettes of the 32 instances -> [matrix][row][instance], with the two exactness flags
    {
        const float* pal = P.palette + (size_t)(i0 + lane) * P.palette_stride_floats;
        for (uint32_t e = wid; e < nm * 3; e += SW_WARPS) {
            const uint32_t m = e / 3, r = e - m * 3;
            const uint32_t sm = P.bone_map ? (uint32_t)__ldg(P.bone_map + m) : m;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                v = __ldg(reinterpret_cast<const float4*>(pal + sm * 16 + r * 4));
                const uint32_t a = __float_as_uint(v.x), b = __float_as_uint(v.y), c = __float_as_uint(v.z), d = __float_as_uint(v.w);
                if (sw_neg_zero_or_denormal(a) || sw_neg_zero_or_denormal(b) || sw_neg_zero_or_denormal(c) || sw_neg_zero_or_denormal(d)) s_dirty[m] = 1;
                if (sw_huge(a) || sw_huge(b) || sw_huge(c) || sw_huge(d)) s_dirty[nm] = 1;
            }
            s_pal[e * 32 + lane] = v;
        }
    }
    // ---- the warp-uniform data of this CTA's vertices: bone indices, weights, zero-weight mask, keyframe(s)
    for (uint32_t j = tid; j < vcount; j += SW_THREADS) {
        const uint32_t v = vbeg + j;
        const uint8_t* bi = P.blend_index + (size_t)v * P.blend_index_stride;
        const float* bw = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(P.blend_weight) + (size_t)v * P.blend_weight_stride);
        uint32_t idx4 = 0, zmask = 0;
        float w[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int b = 0; b < NW; ++b) {
            idx4 |= (uint32_t)__ldg(bi + b) << (8 * b);
            w[b] = __ldg(bw + b);
            if (b > 0 && __float_as_uint(w[b]) == 0u) zmask |= 1u << b;
        }
        // slot 0 must leave no -0 behind: a zero, denormal or tiny first weight could (0 * negative, underflow)
        if ((__float_as_uint(w[0]) & 0x7fffffffu) < 0x35800000u || (__float_as_uint(w[0]) >> 31)) zmask = 0;   // 2^-20
        const float* a = reinterpret_cast<const float*>(P.src + (size_t)v * 24);
        float k1[6], kd[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { k1[k] = __ldg(a + k); kd[k] = 0.f; }
        if (FUSED) {
            const float* b2 = reinterpret_cast<const float*>(P.key2 + (size_t)v * 24);
#pragma unroll
            for (int k = 0; k < 6; ++k) kd[k] = fsub(__ldg(b2 + k), k1[k]);          // (b - a) is instance-invariant
        }
        float4* o = s_vtx + (size_t)j * 5;
        o[0] = make_float4(__uint_as_float(idx4), __uint_as_float(zmask), 0.f, 0.f);
        o[1] = make_float4(w[0], w[1], w[2], w[3]);
        o[2] = make_float4(k1[0], k1[1], k1[2], k1[3]);
        o[3] = make_float4(k1[4], k1[5], kd[0], kd[1]);
        o[4] = make_float4(kd[2], kd[3], kd[4], kd[5]);
    }
    __syncthreads();                                   // the only CTA barrier: from here on the warps run independently

    const uint32_t block_all = s_dirty[nm];
    const float t = (FUSED && live) ? __ldg(P.tarr + i0 + lane) : 0.f;
    uint8_t* stage = s_stage + (size_t)wid * SW_GROUP * SW_ROW;
    const float4* palL = s_pal + lane;
    const bool dst16 = P.tma_store != 0;               // destination base and instance stride 16-byte aligned

    for (uint32_t vb = wid * SW_VB; vb < vcount; vb += SW_WARPS * SW_VB) {
        const uint32_t nval = min((uint32_t)SW_VB, vcount - vb);
#pragma unroll
        for (int h = 0; h < SW_VB / 2; ++h) {
            float o[12];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const uint32_t vl = min(vb + 2 * h + u, vcount - 1);     // a clamped repeat in the tail is never written out
                const float4* q = s_vtx + (size_t)vl * 5;
                const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4];
                const uint32_t idx4 = __float_as_uint(q0.x);
                const float w[4] = {q1.x, q1.y, q1.z, q1.w};
                float x = q2.x, y = q2.y, z = q2.z, nx = q2.w, ny = q3.x, nz = q3.y;
                if (FUSED) {
                    x = fadd(x, fmul(t, q3.z)); y = fadd(y, fmul(t, q3.w)); z = fadd(z, fmul(t, q4.x));
                    nx = fadd(nx, fmul(t, q4.y)); ny = fadd(ny, fmul(t, q4.z)); nz = fadd(nz, fmul(t, q4.w));
                    normalise3(nx, ny, nz);
                }
                // slots that may be skipped for this vertex (warp-uniform): zero weights, provided slot 0's matrix holds
                // no -0 in any of the 32 instances and nothing in the CTA's palettes is non-finite
                const uint32_t m0 = idx4 & 0xffu;
                const uint32_t skip = (block_all | s_dirty[m0]) ? 0u : __float_as_uint(q0.y);
                float4 r0, r1, r2;
                {
                    const float4* M0 = palL + m0 * 96;
                    if (NW == 1) {
                        r0 = M0[0]; r1 = M0[32]; r2 = M0[64];
                    } else if (NW == 2) {
                        if (skip & 2u) { r0 = M0[0]; r1 = M0[32]; r2 = M0[64]; }       // lerp(M0, M1, +0) == M0 bit for bit here
                        else {
                            const float4* M1 = palL + ((idx4 >> 8) & 0xffu) * 96;
                            r0 = lerp4(M0[0], M1[0], w[1]); r1 = lerp4(M0[32], M1[32], w[1]); r2 = lerp4(M0[64], M1[64], w[1]);
                        }
                    } else {
                        r0 = mul4(M0[0], w[0]); r1 = mul4(M0[32], w[0]); r2 = mul4(M0[64], w[0]);
#pragma unroll
                        for (int b = 1; b < NW; ++b) {
                            if (!((skip >> b) & 1u)) {
                                const float4* Mb = palL + ((idx4 >> (8 * b)) & 0xffu) * 96;
                                r0 = madd4(Mb[0], w[b], r0); r1 = madd4(Mb[32], w[b], r1); r2 = madd4(Mb[64], w[b], r2);
                            }
                        }
                    }
                }
                xform_pos(r0, r1, r2, x, y, z, o[6 * u], o[6 * u + 1], o[6 * u + 2]);
                xform_norm(r0, r1, r2, nx, ny, nz, o[6 * u + 3], o[6 * u + 4], o[6 * u + 5]);
            }
            float4* row = reinterpret_cast<float4*>(stage + lane * SW_ROW + h * 48);
            row[0] = make_float4(o[0], o[1], o[2], o[3]);
            row[1] = make_float4(o[4], o[5], o[6], o[7]);
            row[2] = make_float4(o[8], o[9], o[10], o[11]);
        }
        __syncwarp();
        // ---- write-out: instance rows of nval * 24 contiguous bytes
        const size_t voff = (size_t)(vbeg + vb) * 24;
        if (nval == SW_VB && dst16) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const uint32_t c = lane + 32u * k, inst = c / 6u, part = c - inst * 6u;
                if (i0 + inst < P.num_instances)
                    __stcs(reinterpret_cast<float4*>(P.dst + (size_t)(i0 + inst) * P.dst_inst_stride + voff) + part,
                           *reinterpret_cast<const float4*>(stage + inst * SW_ROW + part * 16));
            }
        } else {
            const uint32_t words = nval * 6u;
            for (uint32_t c = lane; c < SW_GROUP * words; c += 32u) {
                const uint32_t inst = c / words, wd = c - inst * words;
                if (i0 + inst < P.num_instances)
                    __stcs(reinterpret_cast<uint32_t*>(P.dst + (size_t)(i0 + inst) * P.dst_inst_stride + voff) + wd,
                           *reinterpret_cast<const uint32_t*>(stage + inst * SW_ROW + wd * 4));
            }
        }
        __syncwarp();
    }
}

static size_t skin_wide_smem(uint32_t num_matrices, uint32_t verts_per_cta)
{
    return (size_t)num_matrices * 1536 + (size_t)verts_per_cta * SW_VTX + (size_t)SW_WARPS * SW_GROUP * SW_ROW + ((size_t)num_matrices + 1) * 4 + 16;
}

// Whether this call can take the wide kernel: pos+normal streams, one source for all instances (fused morph or a shared
// bind pose), no materialised morph intermediate, and the palette block of 32 instances fits shared memory.
bool skin_wide_eligible(const SkinTileParams& P, bool norm, bool fused)
{
    if (!norm || P.morph_out) return false;
    if (!fused && P.src_inst_stride != 0) return false;
    if (P.num_instances < 2 * SW_GROUP) return false;
    return skin_wide_smem(P.num_matrices, 64) <= 200 * 1024;
}

template <int NW, bool FUSED>
static int launch_wide(ogrecu_ctx* ctx, const SkinTileParams& P, dim3 grid, size_t smem, uint32_t vpc)
{
    static size_t granted[64] = {0};
    if (smem > granted[ctx->device & 63]) {
        OGRECU_CUDA(cudaFuncSetAttribute(skin_wide_kernel<NW, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        granted[ctx->device & 63] = smem;
    }
    skin_wide_kernel<NW, FUSED><<<grid, SW_THREADS, smem, ctx->stream>>>(P, vpc);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

int launch_skin_wide(ogrecu_ctx* ctx, SkinTileParams P, uint32_t nw, bool fused)
{
    const uint32_t groups = div_up(P.num_instances, SW_GROUP);
    // vertices per CTA: a multiple of 64 (one round of all warps); split the mesh so that the grid has >= ~6 waves of
    // CTAs (one CTA per SM with a large palette) while a CTA still amortises its palette copy over >= 256 vertices
    uint32_t split = 1;
    const uint32_t min_vpc = std::max(64u, 8u * P.num_matrices);       // output bytes per CTA >= 4 x its palette bytes
    while (split < 8 && (uint64_t)groups * split < (uint64_t)ctx->sm_count * 6 && div_up(P.num_vertices, split * 2) >= min_vpc) split *= 2;
    uint32_t vpc = div_up(div_up(P.num_vertices, split), 64) * 64;
    if (ctx->skin_ipc > 0) vpc = div_up((uint32_t)ctx->skin_ipc, 64) * 64;      // development knob: vertices per CTA
    while (skin_wide_smem(P.num_matrices, vpc) > 200 * 1024 && vpc > 64) vpc -= 64;
    const size_t smem = skin_wide_smem(P.num_matrices, vpc);
    dim3 grid(div_up(P.num_vertices, vpc), groups);
    OGRECU_REQUIRE(groups <= 65535, "too many instance groups");
    if (fused) {
        switch (nw) {
        case 1: return launch_wide<1, true>(ctx, P, grid, smem, vpc);
        case 2: return launch_wide<2, true>(ctx, P, grid, smem, vpc);
        case 3: return launch_wide<3, true>(ctx, P, grid, smem, vpc);
        default: return launch_wide<4, true>(ctx, P, grid, smem, vpc);
        }
    }
    switch (nw) {
    case 1: return launch_wide<1, false>(ctx, P, grid, smem, vpc);
    case 2: return launch_wide<2, false>(ctx, P, grid, smem, vpc);
    case 3: return launch_wide<3, false>(ctx, P, grid, smem, vpc);
    default: return launch_wide<4, false>(ctx, P, grid, smem, vpc);
    }
}

}  // namespace ogrecu
