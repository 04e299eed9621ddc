// skin_lane.cu — instanced software skinning (and fused morph+skin) with an INSTANCE PAIR per lane, from a
// precomputed per-mesh plan.  The default for the pos+normal streams of an instanced mesh (BASELINE config 2).
//
// Reference work replaced: Mesh::softwareVertexMorph -> OptimisedUtil::softwareVertexMorph and
// Mesh::softwareVertexBlend -> OptimisedUtil::softwareVertexSkinning (OgreMain/src/OgreMesh.cpp:1792-1946,
// arithmetic OgreOptimisedUtilSSE.cpp:278-421,606-761 and OgreSIMDHelper.h:204-227) for many instances of one mesh —
// what Entity::updateAnimation does per entity and sub-entity (OgreMain/src/OgreEntity.cpp:747-911).
//
// Why (profiles/r01q_skin_pair_fused_full.txt): skin_pair.cu gives a thread one VERTEX and walks the instances, so the
// lanes of a warp address unrelated bone matrices: 192 B of palette through LDS.128 per 4-weight vertex-instance
// (shared-memory pipe 80 % busy, a quarter of it bank conflicts), 470 instructions per thread-pass of which under 200
// are the arithmetic the reference demands, two CTA barriers per pass.  3.2 cycles per vertex-instance per SM where
// the FP32 pipe needs ~1.0 and the 24 output bytes ~0.4.  Here the roles are transposed:
//   * a lane owns TWO INSTANCES (A, B) and keeps every value as the f32x2 pair (A, B): FFMA2/FADD2 do both;
//     a warp = 64 instances walks the vertices, so bone indices, weights and keyframes are WARP-UNIFORM: one
//     broadcast LDS, no per-lane address arithmetic, no divergence;
//   * zero-weight slots (45 % of Sinbad's) and unit first weights are skipped by uniform branches; consecutive
//     single-bone vertices of the same bone keep the matrix in registers.  SSE multiplies through where General skips
//     (`if (weight)`, OgreOptimisedUtilGeneral.cpp:135); the skip is bitwise exact unless the accumulated element is
//     -0 (needs a -0 / negative denormal in the slot-0 matrix) or a skipped matrix is not finite (0 * inf = NaN).
//     Both are properties of the palettes, tested while they are gathered: such vertices take every slot;
//   * the palettes of the 64 instances sit in shared memory as [slot][16-byte chunk][lane] = (A.e, B.e, A.e+1, B.e+1):
//     every LDS.128 of a warp reads 512 contiguous bytes and returns two ready pairs.  Only the bones a vertex
//     chunk references are resident (the PLAN, built once per mesh from its static blend stream, cuts the mesh
//     into chunks of <= 128 vertices referencing <= LN_MAXB bones: Sinbad's body needs 13 of 55 on average), so two
//     CTAs fit an SM and one CTA's palette gather hides behind the other's arithmetic;
//   * results leave through per-warp staging rows [instance][2 vertices x 24 B] by ONE TMA tensor store per
//     instance half (cp.async.bulk.tensor.2d, rows = instance pairs with stride 2 x instance stride, so odd vertex
//     counts stay 16-byte legal): no read-back of the staging, no per-row address arithmetic, OOB rows/columns
//     (ragged last group, odd last vertex) clipped by the TMA unit;
//   * warps only synchronise with themselves: no CTA barrier after the prologue;
//   * all sub-meshes of a frame are ONE launch (grid = sum of chunk x group items).
// Arithmetic is operation for operation that of skin_pair.cu (bit-identical results, tests/test_skin_lane_gpu.py).
#include <cuda.h>
#include "common.cuh"
#include "skin_math.cuh"
#include "exact_math.cuh"
#include "skin_lane.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

using namespace ogrecu;

namespace ogrecu {

typedef unsigned long long u64;

__device__ __forceinline__ uint32_t ln_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u64 ln_pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void ln_upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// RN(a*b) per half: an fma with a -0.0 addend that ptxas cannot see through (skin_pair.cu header)
__device__ __forceinline__ u64 ln_mul(u64 a, u64 b, u64 nz) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(nz)); return d; }
__device__ __forceinline__ u64 ln_add(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 ln_sub(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

__device__ __forceinline__ bool ln_neg_zero_or_denormal(uint32_t u) { return (u - 0x80000000u) < 0x00800000u; }
__device__ __forceinline__ bool ln_huge(uint32_t u) { return (u & 0x7fffffffu) >= 0x7f000000u; }   // >= 2^127, inf, NaN

__device__ __forceinline__ void ln_tma_store_2d(const CUtensorMap* map, uint32_t src_smem, int32_t x, int32_t y)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src_smem), "r"(x), "r"(y) : "memory");
}

// (sqrt(x), 1 / sqrt(x)) by the library's correctly rounded functions: lengths outside the packed Newton sequences' range.  Out of
// line: inlined, its eight conditional slow-path calls sat in the middle of the walk's loop and of its register allocation.
__device__ __noinline__ float2 ln_sqrt_rcp_slow(const float x)
{
    const float s = __fsqrt_rn(x);
    return make_float2(s, __frcp_rn(s));
}

// sqrt(x) and 1/sqrt(x) of both halves, correctly rounded each (== __frcp_rn(__fsqrt_rn(x))), for the two vertices of a
// pair at once: the packed Newton sequences of exact_math.cuh when the whole warp is in their range (always, for
// unit-scale normals), written with the negated intermediates (-h, -d, -x: round-to-nearest is sign-symmetric, so every
// step has the bits of the original) so that no sign flips are spent; the library functions otherwise.  The two
// independent chains interleave in the pipeline.
__device__ __forceinline__ void ln_sqrt_rcp2(const u64 x0, const u64 x1, const u64 nz, u64& len0, u64& inv0, u64& len1, u64& inv1)
{
    float a0, b0, a1, b1;
    ln_upk(x0, a0, b0); ln_upk(x1, a1, b1);
    const bool fast = em_fast_range(__float_as_uint(a0)) && em_fast_range(__float_as_uint(b0)) &&
                      em_fast_range(__float_as_uint(a1)) && em_fast_range(__float_as_uint(b1));
    if (__all_sync(0xffffffffu, fast)) {
        const u64 half = 0x3f0000003f000000ull, nhalf = 0xbf000000bf000000ull, one = 0x3f8000003f800000ull, zero = 0ull;
        float y0a, y0b, y1a, y1b;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0a) : "f"(a0));
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0b) : "f"(b0));
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1a) : "f"(a1));
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1b) : "f"(b1));
        const u64 y0 = ln_pk(y0a, y0b), y1 = ln_pk(y1a, y1b);
        const u64 nx0 = ln_sub(zero, x0), nx1 = ln_sub(zero, x1);
        u64 g0 = em_pfma(x0, y0, nz), g1 = em_pfma(x1, y1, nz);
        u64 h0 = em_pfma(nhalf, y0, nz), h1 = em_pfma(nhalf, y1, nz);           // -h
        const u64 r0 = em_pfma(h0, g0, half), r1 = em_pfma(h1, g1, half);
        g0 = em_pfma(g0, r0, g0); g1 = em_pfma(g1, r1, g1);
        h0 = em_pfma(h0, r0, h0); h1 = em_pfma(h1, r1, h1);
        const u64 d0 = em_pfma(g0, g0, nx0), d1 = em_pfma(g1, g1, nx1);         // -(x - g*g)
        len0 = em_pfma(d0, h0, g0); len1 = em_pfma(d1, h1, g1);
        float l0a, l0b, l1a, l1b, z0a, z0b, z1a, z1b;
        ln_upk(len0, l0a, l0b); ln_upk(len1, l1a, l1b);
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(z0a) : "f"(l0a));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(z0b) : "f"(l0b));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(z1a) : "f"(l1a));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(z1b) : "f"(l1b));
        u64 z0 = ln_pk(z0a, z0b), z1 = ln_pk(z1a, z1b);
        const u64 nl0 = ln_sub(zero, len0), nl1 = ln_sub(zero, len1);
        u64 e0 = em_pfma(nl0, z0, one), e1 = em_pfma(nl1, z1, one);
        z0 = em_pfma(z0, e0, z0); z1 = em_pfma(z1, e1, z1);
        e0 = em_pfma(nl0, z0, one); e1 = em_pfma(nl1, z1, one);
        inv0 = em_pfma(z0, e0, z0); inv1 = em_pfma(z1, e1, z1);
    } else {
        const float2 r0 = ln_sqrt_rcp_slow(a0), r1 = ln_sqrt_rcp_slow(b0), r2 = ln_sqrt_rcp_slow(a1), r3 = ln_sqrt_rcp_slow(b1);
        len0 = ln_pk(r0.x, r1.x); len1 = ln_pk(r2.x, r3.x);
        inv0 = ln_pk(r0.y, r1.y); inv1 = ln_pk(r2.y, r3.y);
    }
}

// shared memory through 32-bit shared-window addresses: no generic -> shared conversion and no 64-bit pointer in the walk's
// registers.  volatile: a unit buffer is refilled between units, the loads may not be merged across the barrier waits.
__device__ __forceinline__ float4 ln_lds_f4(const uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 ln_lds_u4(const uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void ln_sts_f4(const uint32_t a, const float x, const float y, const float z, const float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
// the six LDS.128 of one palette matrix of the lane's instance pair ([chunk][lane] layout: 512 bytes between chunks)
__device__ __forceinline__ void ln_lds6(const uint32_t a, u64 (&q)[12])
{
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%12];\n\t"
                 "ld.shared.v2.b64 {%2, %3}, [%12 + 512];\n\t"
                 "ld.shared.v2.b64 {%4, %5}, [%12 + 1024];\n\t"
                 "ld.shared.v2.b64 {%6, %7}, [%12 + 1536];\n\t"
                 "ld.shared.v2.b64 {%8, %9}, [%12 + 2048];\n\t"
                 "ld.shared.v2.b64 {%10, %11}, [%12 + 2560];"
                 : "=l"(q[0]), "=l"(q[1]), "=l"(q[2]), "=l"(q[3]), "=l"(q[4]), "=l"(q[5]), "=l"(q[6]), "=l"(q[7]), "=l"(q[8]), "=l"(q[9]), "=l"(q[10]), "=l"(q[11])
                 : "r"(a));
}
// e = RN(RN(m * w) + e), the sum back into e's own registers
__device__ __forceinline__ void ln_acc(u64& e, const u64 m, const u64 w, const u64 nz)
{
    asm("{\n\t.reg .b64 t;\n\tfma.rn.f32x2 t, %1, %2, %3;\n\tadd.rn.f32x2 %0, t, %0;\n\t}" : "+l"(e) : "l"(m), "l"(w), "l"(nz));
}
// one weighted matrix added to the collapsed one
__device__ __forceinline__ void ln_add_slot(u64 (&E)[12], const uint32_t a, const u64 W, const u64 NZ)
{
    u64 q[12];
    ln_lds6(a, q);
#pragma unroll
    for (int k = 0; k < 12; ++k) ln_acc(E[k], q[k], W, NZ);
}

// The output stage.  A quad = 4 consecutive vertices = one 96-byte row per instance = one TMA box of 32 rows (48-byte rows run
// the TMA unit at a quarter of the HBM rate, 96-byte rows at all of it: tools/ubench/store_pattern.cu).  A box must start on
// a 16-byte boundary.  Even instances start on one; odd instances start 8 mod 16 when the instance stride is 8 mod 16 (an odd
// vertex count), their quads then start at vertices 4q - 1 (PHASE 1).  A warp's 64 instances are all even or all odd (a group
// = the even or the odd half of 128 consecutive instances; lane l holds instances 2l and 2l + 64 of that half as its pair), so
// the whole warp walks the same quads: vertices [4q - PHASE, 4q - PHASE + 4).  Vertices no whole in-range quad covers (the
// first three of a shifted row, the last ones of the mesh) are stored directly, 24 bytes per lane.
struct LaneWalk {
    uint32_t pal_lane;              // shared address of the palette window + lane * 16
    uint32_t rec;                   // [1 + chunk vertices][2 x 16 B]; entry 0 = the vertex before the chunk
    uint32_t key;                   // [1 + chunk vertices][48 B (k1 6, k2 - k1 6) or, without morph, 32 B (k1 6, pad)]
    uint32_t stage_lane;            // this warp's staging rows ([2 halves][32][96 B]) + lane * 96
    uint32_t job;                   // index into LaneParams::job
    uint32_t odd;                   // the group holds the odd instances of its 128
    int32_t v_begin, nverts, nv_chunk;
    int32_t p0, p1;                 // chunk-local vertex pairs of this warp: vertices (2p - PHASE, 2p - PHASE + 1)
    int32_t row0;                   // the group's first row in the tensor map of its parity (= first instance / 2)
    uint32_t lane;
};

// vertex pair (ga, ga + 1) of both halves: fa / fb = the 12 floats of the lane's first / second instance.  qs = first vertex of
// the pair's quad, second = the pair is the quad's second one.
template <int PHASE>
__device__ __forceinline__ void ln_emit_pair(const LaneParams& P, const LaneWalk& K, const int32_t ga, const int32_t qs, const bool second,
                                             const float (&fa)[12], const float (&fb)[12])
{
    if (qs >= 0 && qs + 4 <= K.nverts) {
        if (!second) {                              // the buffers' previous boxes have been read (issued a quad ago)
            if (K.lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
        }
        const uint32_t r = K.stage_lane + (second ? 48u : 0u);
        ln_sts_f4(r, fa[0], fa[1], fa[2], fa[3]); ln_sts_f4(r + 16u, fa[4], fa[5], fa[6], fa[7]); ln_sts_f4(r + 32u, fa[8], fa[9], fa[10], fa[11]);
        ln_sts_f4(r + LN_BOX_BYTES, fb[0], fb[1], fb[2], fb[3]); ln_sts_f4(r + LN_BOX_BYTES + 16u, fb[4], fb[5], fb[6], fb[7]);
        ln_sts_f4(r + LN_BOX_BYTES + 32u, fb[8], fb[9], fb[10], fb[11]);
        if (second) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // my STS visible to the async proxy
            __syncwarp();
#ifndef LN_NO_STORE   // (timing experiments only: the arithmetic without the output stage)
            if (K.lane == 0) {
                const LaneJob& J = P.job[K.job];
                const CUtensorMap* map = K.odd ? &J.mapB : &J.mapA;
                const int32_t x = qs * 6 + 2 * PHASE;                           // shifted: map B starts 8 bytes before the odd instance
                ln_tma_store_2d(map, K.stage_lane, x, K.row0);
                ln_tma_store_2d(map, K.stage_lane + LN_BOX_BYTES, x, K.row0 + 32);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
#endif
        }
    } else {                                        // head / tail of the row: 24 bytes per vertex, directly
        const LaneJob& J = P.job[K.job];
        const uint32_t ia = 2u * (uint32_t)K.row0 + K.odd + 2u * K.lane, ib = ia + 64u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t inst = h ? ib : ia;
            if (inst >= P.ninst) continue;
            uint8_t* dst = J.dst + (size_t)inst * J.dst_stride;
            const float (&f)[12] = h ? fb : fa;
            if (ga >= 0 && ga < K.nverts) {
                float2* d = reinterpret_cast<float2*>(dst + (size_t)ga * 24);
                d[0] = make_float2(f[0], f[1]); d[1] = make_float2(f[2], f[3]); d[2] = make_float2(f[4], f[5]);
            }
            if (ga + 1 >= 0 && ga + 1 < K.nverts) {
                float2* d = reinterpret_cast<float2*>(dst + (size_t)(ga + 1) * 24);
                d[0] = make_float2(f[6], f[7]); d[1] = make_float2(f[8], f[9]); d[2] = make_float2(f[10], f[11]);
            }
        }
    }
}

// The chunk walk of one warp: vertex pairs p0..p1-1 in order, every value as the (A, B) pair of the lane's two instances.
template <int NW, bool MORPH, int PHASE>
__device__ __forceinline__ void lane_walk(const LaneParams& P, const LaneWalk& K, const u64 T, const u64 NZ)
{
    u64 E[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) E[k] = 0;
    uint32_t cur = 0xffffffffu;                       // slot offset whose matrix E is a plain copy of, or none
#pragma unroll 1
    for (int32_t p = K.p0; p < K.p1; ++p) {
        u64 V[2][6], N[2][3], L2[2];
        float fa[12], fb[12];                                            // the two vertices' results of the lane's first / second instance
        // ---- both vertices: keyframes -> animated source
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t vl = (uint32_t)min(2 * p + u + 1 - PHASE, K.nv_chunk);     // entry of the vertex; past the last one: repeat it (never stored)
            const uint32_t ky = K.key + vl * (MORPH ? 48u : 32u);
            const float4 a = ln_lds_f4(ky), b = ln_lds_f4(ky + 16u);
            V[u][0] = ln_pk(a.x, a.x); V[u][1] = ln_pk(a.y, a.y); V[u][2] = ln_pk(a.z, a.z);     // warp-uniform scalars: ptxas folds the
            V[u][3] = ln_pk(a.w, a.w); V[u][4] = ln_pk(b.x, b.x); V[u][5] = ln_pk(b.y, b.y);     // duplication into .F32 broadcast operands
            if (MORPH) {
                // dst = a + t*(b-a) (OgreOptimisedUtilGeneral.cpp:249-256); (b - a) is instance-invariant
                const float4 c = ln_lds_f4(ky + 32u);
                V[u][0] = ln_add(V[u][0], ln_mul(T, ln_pk(b.z, b.z), NZ)); V[u][1] = ln_add(V[u][1], ln_mul(T, ln_pk(b.w, b.w), NZ));
                V[u][2] = ln_add(V[u][2], ln_mul(T, ln_pk(c.x, c.x), NZ)); V[u][3] = ln_add(V[u][3], ln_mul(T, ln_pk(c.y, c.y), NZ));
                V[u][4] = ln_add(V[u][4], ln_mul(T, ln_pk(c.z, c.z), NZ)); V[u][5] = ln_add(V[u][5], ln_mul(T, ln_pk(c.w, c.w), NZ));
                L2[u] = ln_add(ln_add(ln_mul(V[u][3], V[u][3], NZ), ln_mul(V[u][4], V[u][4], NZ)), ln_mul(V[u][5], V[u][5], NZ));
            }
        }
        if (MORPH) {
            // Vector3::normalise (OgreVector.h:494-508): scale by 1/len only when len > 0
            u64 len[2], inv[2];
            ln_sqrt_rcp2(L2[0], L2[1], NZ, len[0], inv[0], len[1], inv[1]);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float la, lb, ia, ib;
                ln_upk(len[u], la, lb); ln_upk(inv[u], ia, ib);
                const u64 I = ln_pk(la > 0.0f ? ia : 1.0f, lb > 0.0f ? ib : 1.0f);     // x * 1.0f == x
                V[u][3] = ln_mul(V[u][3], I, NZ); V[u][4] = ln_mul(V[u][4], I, NZ); V[u][5] = ln_mul(V[u][5], I, NZ);
            }
        }
        // ---- per vertex: collapse the weighted matrix, transform
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t vl = (uint32_t)min(2 * p + u + 1 - PHASE, K.nv_chunk);
            const uint4 r0 = ln_lds_u4(K.rec + vl * 32u);
            // E[4*r + c] = (M_A[r][c], M_B[r][c]); control flow is warp-uniform (plan flags)
            const uint32_t fl = r0.x & 15u, off0 = r0.x & ~15u;
            const bool single = NW == 1 || (NW == 2 ? !(fl & 2u) : fl == 1u);
            // (kept across vertices only by single-weight meshes, whose consecutive vertices mostly share the bone: with more weights
            // one vertex in eight found its matrix there, and E stayed live through the keyframe stage for it)
            if (NW == 1) {
                if (off0 != cur) ln_lds6(K.pal_lane + off0, E);
                cur = off0;
            } else ln_lds6(K.pal_lane + off0, E);
            if (NW >= 2 && !single) {
                const float4 w = ln_lds_f4(K.rec + vl * 32u + 16u);
                if (NW == 2) {
                    u64 q[12];
                    ln_lds6(K.pal_lane + r0.y, q);
                    const u64 W1 = ln_pk(w.y, w.y);
#pragma unroll
                    for (int k = 0; k < 12; ++k) E[k] = ln_add(ln_mul(ln_sub(q[k], E[k]), W1, NZ), E[k]);     // lerp(M0, M1, w1) = (M1 - M0) * w1 + M0
                } else {
                    if (!(fl & 1u)) {                                   // w0 == 1.0f: 1 * x == x, nothing to do
                        const u64 W0 = ln_pk(w.x, w.x);
#pragma unroll
                        for (int k = 0; k < 12; ++k) E[k] = ln_mul(E[k], W0, NZ);
                    }
                    if (fl & 2u) ln_add_slot(E, K.pal_lane + r0.y, ln_pk(w.y, w.y), NZ);
                    if (fl & 4u) ln_add_slot(E, K.pal_lane + r0.z, ln_pk(w.z, w.z), NZ);
                    if (NW == 4 && (fl & 8u)) ln_add_slot(E, K.pal_lane + r0.w, ln_pk(w.w, w.w), NZ);
                }
            }
            // position: (r.x*x + r.y*y) + (r.z*z + r.w)   (__MM_DOT4x3_PS, OgreSIMDHelper.h:204-217).  The last addition per instance:
            // its results land in the order the staging rows want them instead of as (A, B) pairs that need re-sorting.
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const u64 s0 = ln_add(ln_mul(E[4 * r], V[u][0], NZ), ln_mul(E[4 * r + 1], V[u][1], NZ)), s1 = ln_add(ln_mul(E[4 * r + 2], V[u][2], NZ), E[4 * r + 3]);
                float s0a, s0b, s1a, s1b;
                ln_upk(s0, s0a, s0b); ln_upk(s1, s1a, s1b);
                fa[6 * u + r] = __fadd_rn(s0a, s1a); fb[6 * u + r] = __fadd_rn(s0b, s1b);
            }
            // normal: (r.x*x + r.y*y) + r.z*z (OgreSIMDHelper.h:220-227); its exact 1/sqrt(len2) for both vertices below
#pragma unroll
            for (int r = 0; r < 3; ++r)
                N[u][r] = ln_add(ln_add(ln_mul(E[4 * r], V[u][3], NZ), ln_mul(E[4 * r + 1], V[u][4], NZ)), ln_mul(E[4 * r + 2], V[u][5], NZ));
            L2[u] = ln_add(ln_add(ln_mul(N[u][0], N[u][0], NZ), ln_mul(N[u][1], N[u][1], NZ)), ln_mul(N[u][2], N[u][2], NZ));
        }
        {
            u64 len[2], R[2];
            ln_sqrt_rcp2(L2[0], L2[1], NZ, len[0], R[0], len[1], R[1]);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float ra, rb;
                ln_upk(R[u], ra, rb);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    float na, nb;
                    ln_upk(N[u][r], na, nb);
                    fa[6 * u + 3 + r] = __fmul_rn(na, ra); fb[6 * u + 3 + r] = __fmul_rn(nb, rb);
                }
            }
        }
        // ---- output
        const int32_t ga = K.v_begin + 2 * p - PHASE;                   // mesh vertex of the pair's first
        ln_emit_pair<PHASE>(P, K, ga, K.v_begin + 4 * (p >> 1) - PHASE, (p & 1) != 0, fa, fb);
    }
}

__device__ __forceinline__ void ln_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void ln_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void ln_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}

#ifndef LN_CONSUMER_SLEEP
#define LN_CONSUMER_SLEEP 0
#endif
__device__ __forceinline__ void ln_mbar_wait_consumer(uint32_t bar, uint32_t parity)
{
#if LN_CONSUMER_SLEEP > 0
    uint32_t ok;
    for (;;) {          // 15 warps polling in a tight loop take issue slots from the one warp they are waiting for
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(LN_CONSUMER_SLEEP);
    }
#else
    ln_mbar_wait(bar, parity);
#endif
}

__device__ __forceinline__ void ln_mbar_wait_relaxed(uint32_t bar, uint32_t parity)      // off the critical path: back off between polls
{
    uint32_t ok;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(32);
    }
}

// Palettes -> the pair layout the lane kernel keeps in shared memory, once per frame: for every group of 64 instances and
// every palette matrix, 3 KB = [6 chunks][32 lanes] x 16 bytes (A.e, B.e, A.e+1, B.e+1) with A, B = the lane's two instances,
// so that the lane kernel's producer fetches a chunk's matrices with one bulk copy each and nobody transposes per unit.
// Also the two exactness hazards per (group, matrix): bit 0 = some instance holds a -0 or a negative denormal (a zero
// weight may not be skipped when that is the slot-0 matrix), bit 1 = something is not finite or >= 2^127.
// One warp per (group, matrix); lane = instance pair (2l and 2l + 64 of the group's even or odd half of 128 instances).
__global__ void __launch_bounds__(256) palette_repack_kernel(const float* __restrict__ palette, size_t pal_stride_floats, uint32_t ninst, uint32_t nmat,
                                                              uint32_t groups, uint8_t* __restrict__ ipal, uint32_t* __restrict__ iflags)
{
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (w >= groups * nmat) return;
    const uint32_t g = w / nmat, m = w - g * nmat;
    const uint32_t i0 = (g >> 1) * 128u + (g & 1u) + 2u * lane;                                 // group = even or odd instances of 128
    const uint32_t ia = min(i0, ninst - 1u), ib = min(i0 + 64u, ninst - 1u);                    // past the end: repeat the last instance
    const float4* pa = reinterpret_cast<const float4*>(palette + (size_t)ia * pal_stride_floats + m * 16u);
    const float4* pb = reinterpret_cast<const float4*>(palette + (size_t)ib * pal_stride_floats + m * 16u);
    float4 A[3], B[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) { A[r] = __ldg(pa + r); B[r] = __ldg(pb + r); }
    bool dz = false, hg = false;
    float4* d = reinterpret_cast<float4*>(ipal + (size_t)w * LN_SLOT_BYTES) + lane;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const uint32_t u[8] = {__float_as_uint(A[r].x), __float_as_uint(A[r].y), __float_as_uint(A[r].z), __float_as_uint(A[r].w),
                               __float_as_uint(B[r].x), __float_as_uint(B[r].y), __float_as_uint(B[r].z), __float_as_uint(B[r].w)};
#pragma unroll
        for (int k = 0; k < 8; ++k) { dz |= ln_neg_zero_or_denormal(u[k]); hg |= ln_huge(u[k]); }
        d[(2 * r) * 32] = make_float4(A[r].x, B[r].x, A[r].y, B[r].y);
        d[(2 * r + 1) * 32] = make_float4(A[r].z, B[r].z, A[r].w, B[r].w);
    }
    const uint32_t f = (__any_sync(0xffffffffu, dz) ? 1u : 0u) | (__any_sync(0xffffffffu, hg) ? 2u : 0u);
    if (lane == 0) iflags[w] = f;
}

// what the producer leaves in a unit buffer besides palettes, records and keyframes
struct LaneUnitMeta { uint32_t done, job, v_begin, nquads, group, pad[3]; };     // 32 bytes, followed by the 32 lanes' (tA, tB)

// One persistent CTA per SM: 16 consumer warps walk work units (one vertex chunk x one group of 64 instances), one PRODUCER
// warp prepares the next unit while they do: it takes the unit from a global counter, gathers the 64 instances' palettes
// (the chunk's bones only) into the [slot][chunk][lane] pair layout, tests them for the two exactness hazards, and copies
// the chunk's plan records (with the zero-skip flags withdrawn where a hazard was found) and keyframes.  Two unit buffers,
// full / empty mbarriers: no CTA barrier after start-up, the consumers never wait for global memory.
// Development only (-DLN_PROFILE, tools/build_variant.sh): cycles CTA 0 spends in the phases of its producer warp and of consumer
// warp 0, accumulated behind the unit counter and read by ogrecu_debug_lane_profile.
#ifdef LN_PROFILE
#define LN_PROF_DECL long long prof_t0 = clock64(); const bool prof_on = blockIdx.x == 0 && (threadIdx.x & 31) == 0
#define LN_PROF(slot) do { if (prof_on) { const long long t1 = clock64(); atomicAdd(reinterpret_cast<unsigned long long*>(P.counter) + 8 + (slot), (unsigned long long)(t1 - prof_t0)); prof_t0 = t1; } } while (0)
#else
#define LN_PROF_DECL
#define LN_PROF(slot)
#endif

template <bool MORPH>
__global__ void __launch_bounds__(LN_THREADS, 1) skin_lane_kernel(const __grid_constant__ LaneParams P, const u64 NZ)
{
    constexpr int KEY_BYTES = MORPH ? 48 : 32;
    constexpr uint32_t RAW_BYTES = ((LN_CHUNK_VERTS + 1) * 4 + 15) & ~15;
    constexpr uint32_t UNIT_BYTES = LN_MAXB * LN_SLOT_BYTES + (LN_CHUNK_VERTS + 1) * LN_REC_BYTES + (LN_CHUNK_VERTS + 1) * KEY_BYTES + RAW_BYTES + 32 + 256;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_stage = smem;                                                       // [consumer warps][A, B][32 x 96 B], 128-byte aligned boxes
    uint8_t* s_unit = s_stage + LN_WARPS * 2 * LN_BOX_BYTES;                       // [LN_NBUF] x {palette, records, keys, meta}
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_unit + LN_NBUF * UNIT_BYTES);  // full[LN_NBUF], empty[LN_NBUF]

    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    if (tid == 0) {
        for (int k = 0; k < LN_NBUF; ++k) {
            ln_mbar_init(ln_smem_u32(&s_bar[k]), 2);                       // full: the producer's expect_tx arrive + its final arrive
            ln_mbar_init(ln_smem_u32(&s_bar[LN_NBUF + k]), LN_WARPS);      // empty: every consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (wid == LN_WARPS) {
        // ================================================================ producer warp
        uint32_t held[LN_NBUF];                                             // which chunk's records / keys each buffer holds
#pragma unroll
        for (int k = 0; k < LN_NBUF; ++k) held[k] = 0xffffffffu;
        // Units are taken from the global counter in PAIRS of consecutive numbers: units are numbered chunk-major, so a pair shares
        // its chunk, and the second unit copies the chunk's records and keyframes — 121 entries of ten loads each, the producer's
        // longest job — from the first unit's buffer instead of fetching them again.  (Taken one by one, the 157 groups of a chunk
        // went to 148 different SMs and every unit refilled them from global memory.  Batches of 4 / 8 / 16 / 32 units: 0.468 /
        // 0.482 / 0.518 / 0.672 ms against 0.463 — the tail of the launch grows with the batch; profiles/r02_notes.md.)
        uint32_t batch_next = 0, cur = 0, cur_end = 0;
        if (lane == 0) batch_next = atomicAdd(P.counter, (uint32_t)LN_UNIT_BATCH);
        LN_PROF_DECL;
        for (uint32_t it = 0;; ++it) {
            const uint32_t b = it % LN_NBUF, use = it / LN_NBUF;           // buffer, and how many times it has been used before
            uint8_t* U = s_unit + b * UNIT_BYTES;
            uint8_t* u_pal = U;
            uint4* u_rec = reinterpret_cast<uint4*>(U + LN_MAXB * LN_SLOT_BYTES);
            uint8_t* u_key = reinterpret_cast<uint8_t*>(u_rec) + (LN_CHUNK_VERTS + 1) * LN_REC_BYTES;
            uint32_t* u_raw = reinterpret_cast<uint32_t*>(u_key + (LN_CHUNK_VERTS + 1) * KEY_BYTES);
            LaneUnitMeta* u_meta = reinterpret_cast<LaneUnitMeta*>(reinterpret_cast<uint8_t*>(u_raw) + RAW_BYTES);
            const uint32_t full = ln_smem_u32(&s_bar[b]);
            // ---- everything that does not touch the buffer first: the unit, its chunk, its hazard flags
            bool new_batch = false;
            if (cur == cur_end) {
                cur = __shfl_sync(0xffffffffu, batch_next, 0);
                cur_end = min(cur + (uint32_t)LN_UNIT_BATCH, P.total_units);
                if (cur_end < cur) cur_end = cur;                       // past the end: `done` below
                new_batch = true;
            }
            const uint32_t unit = cur;
            const bool done = unit >= P.total_units;
            if (!done) ++cur;
            uint32_t j = 0, c = 0, g = 0;
            uint4 ch = make_uint4(0, 0, 0, 0);
            uint32_t gm_mine = 0, fl_mine = 0;
            float2 t_mine = make_float2(0.f, 0.f);
            if (!done) {
#pragma unroll 1
                while (j + 1 < P.njobs && unit >= P.job[j + 1].first_item) ++j;
                const uint32_t local = unit - P.job[j].first_item;
                c = local / P.groups; g = local - c * P.groups;
                const uint4* chrec = P.job[j].chunks + (size_t)c * 3;      // {v_begin, nquads, nslots, -}, 16 x u16 palette matrix of each slot
                ch = __ldg(chrec);
                if (lane < LN_MAXB) gm_mine = (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(chrec + 1) + lane);
                if (lane < ch.z) fl_mine = __ldg(P.iflags + (size_t)g * P.nmat + gm_mine);
                if (new_batch) {
                    batch_next = 0;
                    if (lane == 0) batch_next = atomicAdd(P.counter, (uint32_t)LN_UNIT_BATCH);      // the next batch's number travels while this one is worked on
                }
                if (MORPH) {                                               // the instances' morph weights, as (A, B) per lane
                    const uint32_t i0 = (g >> 1) * 128u + (g & 1u) + 2u * lane;
                    const uint32_t ia = min(i0, P.ninst - 1u), ib = min(i0 + 64u, P.ninst - 1u);
                    t_mine = make_float2(__ldg(P.t + ia), __ldg(P.t + ib));
                }
            }
            LN_PROF(0);                                                    // prep: unit, chunk record, flags, weights
            if (use) ln_mbar_wait_relaxed(ln_smem_u32(&s_bar[LN_NBUF + b]), (use - 1u) & 1u);            // the consumers are done with this buffer
            LN_PROF(1);                                                    // waiting for the buffer
            if (done) {
                if (lane == 0) { u_meta->done = 1; ln_mbar_arrive(full); ln_mbar_arrive(full); }
                break;
            }
            const LaneJob& J = P.job[j];
            const uint32_t v_begin = ch.x, nquads = ch.y, nslots = ch.z;
            const uint32_t nv_chunk = min(4u * nquads, J.nverts - v_begin);
            // ---- palettes: one 3 KB bulk copy per slot from the repacked palettes (already in the [chunk][lane] pair layout),
            // completion counted on the buffer's full barrier
            if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(nslots * (uint32_t)LN_SLOT_BYTES) : "memory");
            }
            __syncwarp();
            if (lane < nslots) {
                const uint8_t* src = P.ipal + ((size_t)g * P.nmat + gm_mine) * LN_SLOT_BYTES;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(ln_smem_u32(u_pal + lane * LN_SLOT_BYTES)), "l"(src), "r"((uint32_t)LN_SLOT_BYTES), "r"(full) : "memory");
            }
            // ---- plan records and keyframes of the chunk (entry 0: the vertex before it); units arrive chunk-major, so the
            // buffer usually holds them already.  The flag word is kept raw on the side: it is re-derived per unit below.
            LN_PROF(2);                                                    // issuing the palette copies
            const uint32_t chunk_id = (j << 24) | c;
            bool have = false;
            int other = -1;                                                 // another buffer that holds this chunk's records
#pragma unroll
            for (int k = 0; k < LN_NBUF; ++k) {
                if ((uint32_t)k == b) { have = held[k] == chunk_id; held[k] = chunk_id; }
                else if (held[k] == chunk_id) other = k;
            }
            if (!have && other >= 0) {
                // the consumers may be reading that buffer: so does this copy, nobody writes it.  (Its flag words were patched for
                // another group of instances; they are re-derived from the raw copy below.)
                const uint32_t n16 = ((nv_chunk + 1u) * (LN_REC_BYTES + KEY_BYTES)) / 16u, r16 = ((nv_chunk + 1u) * 4u + 15u) / 16u;
                const uint8_t* O = s_unit + (uint32_t)other * UNIT_BYTES;
                const uint4* o_rec = reinterpret_cast<const uint4*>(O + LN_MAXB * LN_SLOT_BYTES);
                const uint4* o_key = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(o_rec) + (LN_CHUNK_VERTS + 1) * LN_REC_BYTES);
                const uint4* o_raw = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(o_key) + (LN_CHUNK_VERTS + 1) * KEY_BYTES);
                const uint32_t nrec16 = (nv_chunk + 1u) * (LN_REC_BYTES / 16u);
                for (uint32_t q = lane; q < n16; q += 32u) {
                    if (q < nrec16) u_rec[q] = o_rec[q];
                    else reinterpret_cast<uint4*>(u_key)[q - nrec16] = o_key[q - nrec16];
                }
                for (uint32_t q = lane; q < r16; q += 32u) reinterpret_cast<uint4*>(u_raw)[q] = o_raw[q];
                __syncwarp();
            } else if (!have) {
                // from global memory: all of the lane's loads first (up to four entries of ten loads), then the stores — in one loop
                // the shared-memory stores kept the iterations' loads apart and the fill cost four L2 round trips
                constexpr int NIT = (LN_CHUNK_VERTS + 1 + 31) / 32;
                uint4 rr0[NIT], rr1[NIT];
                float2 ka[NIT][3], kb[NIT][3];
#pragma unroll
                for (int q = 0; q < NIT; ++q) {
                    const uint32_t e = lane + 32u * (uint32_t)q;
                    if (e > nv_chunk) continue;
                    const bool head = e == 0 && v_begin == 0;               // no vertex before the mesh: the shifted walk computes entry 0 all the same
                    const uint32_t gv = head ? 0u : v_begin + e - 1u;       // (and stores nothing for it), so it gets vertex 0's record
                    const uint4* r = head ? J.recs : e == 0 ? J.pre_recs + (size_t)c * 2 : J.recs + (size_t)gv * 2;
                    rr0[q] = __ldg(r); rr1[q] = __ldg(r + 1);
                    const float2* a = reinterpret_cast<const float2*>(J.key1 + (size_t)gv * 24);
                    ka[q][0] = __ldg(a); ka[q][1] = __ldg(a + 1); ka[q][2] = __ldg(a + 2);
                    if (MORPH) {
                        const float2* k2 = reinterpret_cast<const float2*>(J.key2 + (size_t)gv * 24);
                        kb[q][0] = __ldg(k2); kb[q][1] = __ldg(k2 + 1); kb[q][2] = __ldg(k2 + 2);
                    }
                }
#pragma unroll
                for (int q = 0; q < NIT; ++q) {
                    const uint32_t e = lane + 32u * (uint32_t)q;
                    if (e > nv_chunk) continue;
                    float* ky = reinterpret_cast<float*>(u_key + e * KEY_BYTES);
                    ky[0] = ka[q][0].x; ky[1] = ka[q][0].y; ky[2] = ka[q][1].x; ky[3] = ka[q][1].y; ky[4] = ka[q][2].x; ky[5] = ka[q][2].y;
                    if (MORPH) {                                            // (b - a) is instance-invariant
                        ky[6] = fsub(kb[q][0].x, ka[q][0].x); ky[7] = fsub(kb[q][0].y, ka[q][0].y); ky[8] = fsub(kb[q][1].x, ka[q][1].x);
                        ky[9] = fsub(kb[q][1].y, ka[q][1].y); ky[10] = fsub(kb[q][2].x, ka[q][2].x); ky[11] = fsub(kb[q][2].y, ka[q][2].y);
                    }
                    u_rec[e * 2u] = rr0[q]; u_rec[e * 2u + 1u] = rr1[q];
                    u_raw[e] = rr0[q].x;
                }
                __syncwarp();
            }
            LN_PROF(3);                                                    // records / keys
            if (MORPH) reinterpret_cast<float2*>(u_meta + 1)[lane] = t_mine;
            // a vertex whose slot-0 matrix holds a -0 in some instance, or any vertex when something is not finite, takes every slot
            {
                const uint32_t dirty_mask = __ballot_sync(0xffffffffu, (fl_mine & 1u) != 0u);
                const uint32_t huge_any = __any_sync(0xffffffffu, (fl_mine & 2u) != 0u) ? 1u : 0u;
                const uint32_t fullmask = J.nw >= 3 ? (J.nw == 3 ? 6u : 14u) : 2u;
                for (uint32_t e = lane; e <= nv_chunk; e += 32u) {
                    const uint32_t x = u_raw[e];
                    const uint32_t slot0 = (x & ~15u) / LN_SLOT_BYTES;
                    reinterpret_cast<uint32_t*>(&u_rec[e * 2u])[0] = (huge_any | ((dirty_mask >> slot0) & 1u)) ? ((x & ~15u) | fullmask) : x;
                }
            }
            if (lane == 0) { u_meta->done = 0; u_meta->job = j; u_meta->v_begin = v_begin; u_meta->nquads = nquads; u_meta->group = g; }
            __syncwarp();
            if (lane == 0) ln_mbar_arrive(full);                            // release: everything above is visible to who acquires
            LN_PROF(4);                                                    // flag patch, meta, arrive
        }
        return;
    }

    // ==================================================================== consumer warps
    uint8_t* my_stage = s_stage + wid * (2 * LN_BOX_BYTES);
#ifdef LN_PROFILE
    long long prof_t0 = clock64(); const bool prof_on = blockIdx.x == 0 && tid == 0;
#endif
    for (uint32_t it = 0;; ++it) {
        const uint32_t b = it % LN_NBUF;
        const uint8_t* U = s_unit + b * UNIT_BYTES;
        const uint4* u_rec = reinterpret_cast<const uint4*>(U + LN_MAXB * LN_SLOT_BYTES);
        const uint8_t* u_key = reinterpret_cast<const uint8_t*>(u_rec) + (LN_CHUNK_VERTS + 1) * LN_REC_BYTES;
        const LaneUnitMeta* u_meta = reinterpret_cast<const LaneUnitMeta*>(u_key + (LN_CHUNK_VERTS + 1) * KEY_BYTES + RAW_BYTES);
        LN_PROF(5);                                                        // consumer: arrive + loop overhead
        ln_mbar_wait_consumer(ln_smem_u32(&s_bar[b]), (it / LN_NBUF) & 1u);
        LN_PROF(6);                                                        // consumer: waiting for the unit
        if (u_meta->done) break;
        const LaneJob& J = P.job[u_meta->job];
        const uint32_t v_begin = u_meta->v_begin, nquads = u_meta->nquads, group = u_meta->group;
        const uint32_t nv_chunk = min(4u * nquads, J.nverts - v_begin);
        // contiguous runs of quads per warp (consecutive vertices share bones: the matrix stays in registers).  (Handing the
        // quads out on demand through a shared-memory ticket evens the warps out but loses more in per-hand-out overhead:
        // 0.504 -> 0.527 ms on the bench step.)
        // (the run a warp takes rotates from unit to unit: a chunk with fewer runs than warps leaves a different warp idle every time,
        // and the idle one is already a unit ahead when its turn comes)
        const uint32_t per = (nquads + LN_WARPS - 1u) / LN_WARPS;
        const uint32_t run = (wid + it * 4u) % LN_WARPS;
        const uint32_t q0 = min(run * per, nquads), q1 = min(q0 + per, nquads);
        if (q0 < q1) {
            u64 T = 0;
            if (MORPH) T = reinterpret_cast<const u64*>(u_meta + 1)[lane];
            // the odd instances of a mesh whose instance stride is 8 mod 16 walk quads that start one vertex early (PHASE 1): chunk
            // vertices [4 q0 - 1, 4 q1 - 1) — the chunk's last vertex is the next chunk's entry 0 —, up to the end in the mesh's last run
            const uint32_t odd = group & 1u, phase = J.shift & odd;
            const bool mesh_end = q1 == nquads && v_begin + nv_chunk == J.nverts;
            const uint32_t v_end = mesh_end ? nv_chunk : 4u * q1 - phase;
            LaneWalk K;
            K.pal_lane = ln_smem_u32(U) + lane * 16u;
            K.rec = ln_smem_u32(u_rec); K.key = ln_smem_u32(u_key);
            K.stage_lane = ln_smem_u32(my_stage) + lane * 96u;
            K.job = u_meta->job;
            K.odd = odd;
            K.v_begin = (int32_t)v_begin; K.nverts = (int32_t)J.nverts; K.nv_chunk = (int32_t)nv_chunk;
            K.p0 = (int32_t)(2u * q0); K.p1 = (int32_t)((v_end + phase + 1u) >> 1);
            K.row0 = (int32_t)((group >> 1) * 64u);
            K.lane = lane;
            if (phase) {
                switch (J.nw) {
                case 1: lane_walk<1, MORPH, 1>(P, K, T, NZ); break;
                case 2: lane_walk<2, MORPH, 1>(P, K, T, NZ); break;
                case 3: lane_walk<3, MORPH, 1>(P, K, T, NZ); break;
                default: lane_walk<4, MORPH, 1>(P, K, T, NZ); break;
                }
            } else {
                switch (J.nw) {
                case 1: lane_walk<1, MORPH, 0>(P, K, T, NZ); break;
                case 2: lane_walk<2, MORPH, 0>(P, K, T, NZ); break;
                case 3: lane_walk<3, MORPH, 0>(P, K, T, NZ); break;
                default: lane_walk<4, MORPH, 0>(P, K, T, NZ); break;
                }
            }
        }
        __syncwarp();
        LN_PROF(7);                                                        // consumer: walking its run
        if (lane == 0) ln_mbar_arrive(ln_smem_u32(&s_bar[LN_NBUF + b]));     // this warp no longer reads the unit buffer
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // shared memory must outlive the TMA reads
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------ host side

static size_t lane_smem(bool morph)
{
    const size_t unit = (size_t)LN_MAXB * LN_SLOT_BYTES + (size_t)(LN_CHUNK_VERTS + 1) * LN_REC_BYTES + (size_t)(LN_CHUNK_VERTS + 1) * (morph ? 48 : 32) +
                        (((LN_CHUNK_VERTS + 1) * 4 + 15) & ~15) + 32 + 256;
    return (size_t)LN_WARPS * 2 * LN_BOX_BYTES + LN_NBUF * unit + 2 * LN_NBUF * 8;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// rows = instance pairs (stride 2 x instance stride), columns = 4-byte words of the row; box = 4 vertices x 32 rows
static int make_map(CUtensorMap* m, void* base, uint64_t words, uint64_t rows, uint64_t row_stride_bytes)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return OGRECU_ERR_CUDA; }
    const cuuint64_t dims[2] = {words, rows};
    const cuuint64_t strides[1] = {row_stride_bytes};
    const cuuint32_t box[2] = {24, 32};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return OGRECU_ERR_CUDA; }
    return OGRECU_OK;
}

bool skin_lane_eligible(const ogrecu_skin_plan* plan, const ogrecu_anim_job& J, uint32_t num_instances, size_t palette_stride_floats)
{
    const ogrecu_skin_job& j = J.skin;
    if (!plan || J.morph_out || num_instances < 2) return false;
    const bool packed24 = j.src_norm && (const uint8_t*)j.src_norm == (const uint8_t*)j.src_pos + 12 &&
                          (uint8_t*)j.dst_norm == (uint8_t*)j.dst_pos + 12 && j.src_pos_stride == 24 && j.src_norm_stride == 24 &&
                          j.dst_pos_stride == 24 && j.dst_norm_stride == 24;
    if (!packed24) return false;
    if (((uintptr_t)j.dst_pos & 15) || (J.dst_instance_stride_bytes & 7) || J.dst_instance_stride_bytes < (size_t)j.num_vertices * 24) return false;
    if (((uintptr_t)j.src_pos & 3) || (J.key2 && ((uintptr_t)J.key2 & 3))) return false;
    if ((palette_stride_floats & 3) || ((uintptr_t)j.palette & 15)) return false;
    if ((uint64_t)J.dst_instance_stride_bytes * 2 >= (1ull << 40)) return false;
    return plan->num_vertices == j.num_vertices && plan->num_weights == j.num_weights && plan->num_matrices == j.num_matrices;
}

int repack_palettes(ogrecu_ctx* ctx, const float* palette, size_t palette_stride_floats, uint32_t num_instances, uint32_t nmat)
{
    const uint32_t groups = 2 * div_up(num_instances, 128);          // the even and the odd instances of every 128
    OGRECU_REQUIRE((size_t)nmat * 16 <= palette_stride_floats, "palette stride smaller than the matrices the plans reference");
    const size_t ipal_bytes = (size_t)groups * nmat * LN_SLOT_BYTES, iflag_bytes = (size_t)groups * nmat * 4;
    OGRECU_CHECK(scratch_reserve(ctx, ctx->lane_ipal, ipal_bytes + iflag_bytes));
    palette_repack_kernel<<<div_up((size_t)groups * nmat, 8), 256, 0, ctx->stream>>>(palette, palette_stride_floats, num_instances, nmat, groups,
                                                                                 (uint8_t*)ctx->lane_ipal.ptr, (uint32_t*)((uint8_t*)ctx->lane_ipal.ptr + ipal_bytes));
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

int launch_skin_lane(ogrecu_ctx* ctx, const ogrecu_anim_job* jobs, ogrecu_skin_plan* const* plans, uint32_t num_jobs, const float* t_dev,
                     uint32_t num_instances, size_t palette_stride_floats, bool morph, uint32_t nmat)
{
    OGRECU_REQUIRE(num_jobs >= 1 && num_jobs <= LN_MAX_JOBS, "1..8 jobs per lane launch");
    LaneParams P;
    memset(&P, 0, sizeof P);
    P.njobs = num_jobs;
    P.ninst = num_instances;
    P.groups = 2 * div_up(num_instances, 128);
    P.palette = jobs[0].skin.palette;
    P.pal_stride_floats = palette_stride_floats;
    P.t = t_dev;
    uint64_t items = 0;
    for (uint32_t k = 0; k < num_jobs; ++k) {
        const ogrecu_anim_job& J = jobs[k];
        const ogrecu_skin_plan* pl = plans[k];
        OGRECU_REQUIRE(J.skin.palette == P.palette, "the jobs of one frame share the instances' palettes");
        LaneJob& L = P.job[k];
        const uint64_t S = J.dst_instance_stride_bytes;
        const uint64_t words = (uint64_t)J.skin.num_vertices * 6;
        // odd instances: 16-byte aligned when S is a multiple of 16; otherwise (S = 8 mod 16) their map starts 8 bytes early
        // and their groups walk quads of vertices 4q-1 .. 4q+2 so that every box still starts on a 16-byte boundary
        const uint32_t shift = (S & 8) ? 1u : 0u;
        OGRECU_CHECK(make_map(&L.mapA, J.skin.dst_pos, words, (num_instances + 1) / 2, 2 * S));
        OGRECU_CHECK(make_map(&L.mapB, (uint8_t*)J.skin.dst_pos + S - 8 * shift, words + 2 * shift, num_instances / 2, 2 * S));
        L.shift = shift;
        L.dst = (uint8_t*)J.skin.dst_pos;
        L.dst_stride = S;
        L.pre_recs = (const uint4*)pl->d_pre_recs;
        L.key1 = (const uint8_t*)J.skin.src_pos;
        L.key2 = (const uint8_t*)J.key2;
        L.recs = (const uint4*)pl->d_recs;
        L.chunks = (const uint4*)pl->d_chunks;
        L.nw = J.skin.num_weights;
        L.nverts = J.skin.num_vertices;
        L.nchunks = pl->num_chunks;
        L.first_item = (uint32_t)items;
        items += (uint64_t)pl->num_chunks * P.groups;
        OGRECU_REQUIRE(items < 0x7fffffffull, "too many work items");
    }
    const size_t smem = lane_smem(morph);
    static bool granted[2][64] = {};
    if (!granted[morph][ctx->device & 63]) {
        if (morph) OGRECU_CUDA(cudaFuncSetAttribute(skin_lane_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else OGRECU_CUDA(cudaFuncSetAttribute(skin_lane_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        granted[morph][ctx->device & 63] = true;
    }
    // the frame's palettes in the pair layout (repack_palettes, once per frame), then the unit counter of the persistent CTAs
    P.ipal = (const uint8_t*)ctx->lane_ipal.ptr;
    P.iflags = (const uint32_t*)((uint8_t*)ctx->lane_ipal.ptr + (size_t)P.groups * nmat * LN_SLOT_BYTES);
    P.nmat = nmat;
    OGRECU_CHECK(scratch_reserve(ctx, ctx->lane_counter, 256));
#ifdef LN_PROFILE
    OGRECU_CUDA(cudaMemsetAsync(ctx->lane_counter.ptr, 0, 256, ctx->stream));
#else
    OGRECU_CUDA(cudaMemsetAsync(ctx->lane_counter.ptr, 0, 4, ctx->stream));
#endif
    P.counter = (uint32_t*)ctx->lane_counter.ptr;
    P.total_units = (uint32_t)items;
    const unsigned grid = (unsigned)std::min<uint64_t>(items, (uint64_t)ctx->sm_count);
    if (morph) skin_lane_kernel<true><<<grid, LN_THREADS, smem, ctx->stream>>>(P, 0x8000000080000000ull);
    else skin_lane_kernel<false><<<grid, LN_THREADS, smem, ctx->stream>>>(P, 0x8000000080000000ull);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

// ---- the plan: chunks of <= LN_CHUNK_QUADS vertex quads referencing <= LN_MAXB palette entries (the vertex before the chunk
// included: the shifted walk starts there); per vertex the shared-memory offsets of its slots' matrices, which slots to
// take, and the weights.  OGRECU_ERR_UNSUPPORTED when a single quad needs more entries than fit (no bone locality at all).
int build_skin_plan(ogrecu_ctx* ctx, const uint8_t* blend_index, size_t index_stride, const uint8_t* blend_weight, size_t weight_stride,
                    const uint16_t* bone_map, uint32_t nw, uint32_t num_matrices, uint32_t nverts, ogrecu_skin_plan* plan)
{
    struct Chunk { uint32_t v_begin, nquads, nslots, pad; uint16_t bones[16]; };
    static_assert(sizeof(Chunk) == 48 && LN_MAXB <= 16, "chunk record");
    std::vector<Chunk> chunks;
    std::vector<uint32_t> recs((size_t)nverts * 8, 0), pre_recs;
    OGRECU_REQUIRE(num_matrices >= 1 && num_matrices <= 256 && nw >= 1 && nw <= 4, "plan: 1..256 matrices, 1..4 weights");
    for (uint32_t u = 0; u < nverts; ++u)
        for (uint32_t b = 0; b < nw; ++b)
            if (blend_index[(size_t)u * index_stride + b] >= num_matrices) {
                set_error("blend index %u out of range (%u matrices)", (unsigned)blend_index[(size_t)u * index_stride + b], num_matrices);
                return OGRECU_ERR_INVALID;
            }
    uint32_t v = 0;
    uint32_t quad_limit = LN_CHUNK_QUADS;
    while (v < nverts) {
        const uint32_t v_chunk = v;
        bool cut_by_bones = false;
        Chunk c;
        memset(&c, 0, sizeof c);
        c.v_begin = v;
        int slot_of[256];
        for (uint32_t m = 0; m < num_matrices; ++m) slot_of[m] = -1;
        uint32_t nlocal = 0;
        auto add_bone = [&](uint32_t m) {
            if (slot_of[m] < 0) { slot_of[m] = (int)nlocal; c.bones[nlocal++] = (uint16_t)(bone_map ? bone_map[m] : m); }
        };
        // record of vertex `u` in this chunk's slot numbering
        auto make_rec = [&](uint32_t u, uint32_t* r) {
            const uint8_t* bi = blend_index + (size_t)u * index_stride;
            const uint8_t* bw = blend_weight + (size_t)u * weight_stride;
            uint32_t w[4] = {0, 0, 0, 0};
            for (uint32_t q = 0; q < 8; ++q) r[q] = 0;
            for (uint32_t b = 0; b < nw; ++b) {
                memcpy(&w[b], bw + 4 * b, 4);
                r[b] = (uint32_t)slot_of[bi[b]] * LN_SLOT_BYTES;
                r[4 + b] = w[b];
            }
            // slot 0 must leave no -0 behind: a zero, negative, denormal or tiny first weight could (0 * negative, underflow)
            const bool w0_safe = !((w[0] & 0x7fffffffu) < 0x35800000u || (w[0] >> 31));       // >= 2^-20
            uint32_t fl = 0;
            if (nw == 2) fl = (w[1] & 0x7fffffffu) != 0u ? 2u : 0u;                           // SSE's lerp(M0, M1, w1) reads w1 only
            else if (nw >= 3) {
                for (uint32_t b = 1; b < nw; ++b) if (!w0_safe || (w[b] & 0x7fffffffu) != 0u) fl |= 1u << b;
                if (w[0] == 0x3f800000u) fl |= 1u;
            }
            r[0] |= fl;
        };
        if (v > 0) {                                    // the vertex before the chunk
            const uint8_t* bi = blend_index + (size_t)(v - 1) * index_stride;
            for (uint32_t b = 0; b < nw; ++b) add_bone(bi[b]);
        }
        while (c.nquads < quad_limit && v < nverts) {
            uint32_t add[16], nadd = 0;                 // bones the next quad brings in
            for (uint32_t u = v; u < v + 4 && u < nverts; ++u) {
                const uint8_t* bi = blend_index + (size_t)u * index_stride;
                for (uint32_t b = 0; b < nw; ++b) {
                    const uint32_t m = bi[b];
                    bool seen = slot_of[m] >= 0;
                    for (uint32_t q = 0; q < nadd; ++q) seen |= add[q] == m;
                    if (!seen) add[nadd++] = m;
                }
            }
            if (nlocal + nadd > LN_MAXB) {
                if (c.nquads == 0) { set_error("skin plan: one vertex quad references more than %d palette entries", LN_MAXB); return OGRECU_ERR_UNSUPPORTED; }
                cut_by_bones = true;
                break;
            }
            for (uint32_t q = 0; q < nadd; ++q) add_bone(add[q]);
            for (uint32_t u = v; u < v + 4 && u < nverts; ++u) make_rec(u, &recs[(size_t)u * 8]);
            v = std::min(v + 4, nverts);
            c.nquads++;
        }
        // The consumer warps split a chunk's quads evenly (LN_WARPS runs): a chunk the palette limit cut at, say, 22 quads would leave
        // 8 of 15 warps with half the work of the others.  Cut it at the multiple of LN_WARPS below instead and start over there.
        if (cut_by_bones && c.nquads > LN_WARPS && c.nquads % LN_WARPS != 0) {
            quad_limit = c.nquads / LN_WARPS * LN_WARPS;
            v = v_chunk;
            continue;
        }
        quad_limit = LN_CHUNK_QUADS;
        c.nslots = nlocal;
        pre_recs.resize(pre_recs.size() + 8, 0);
        if (c.v_begin > 0) make_rec(c.v_begin - 1, &pre_recs[pre_recs.size() - 8]);
        chunks.push_back(c);
    }
    plan->num_chunks = (uint32_t)chunks.size();
    plan->max_matrix = 0;
    for (const Chunk& c : chunks) for (uint32_t q = 0; q < c.nslots; ++q) plan->max_matrix = std::max<uint32_t>(plan->max_matrix, c.bones[q]);
    const size_t rb = recs.size() * 4, cb = chunks.size() * sizeof(Chunk), pb = pre_recs.size() * 4;
    OGRECU_CUDA(cudaMalloc(&plan->d_recs, std::max<size_t>(rb, 16)));
    OGRECU_CUDA(cudaMalloc(&plan->d_chunks, std::max<size_t>(cb, 16)));
    OGRECU_CUDA(cudaMalloc(&plan->d_pre_recs, std::max<size_t>(pb, 16)));
    OGRECU_CUDA(cudaMemcpy(plan->d_recs, recs.data(), rb, cudaMemcpyHostToDevice));
    OGRECU_CUDA(cudaMemcpy(plan->d_chunks, chunks.data(), cb, cudaMemcpyHostToDevice));
    OGRECU_CUDA(cudaMemcpy(plan->d_pre_recs, pre_recs.data(), pb, cudaMemcpyHostToDevice));
    (void)ctx;
    return OGRECU_OK;
}

}  // namespace ogrecu

// ------------------------------------------------------------------------------------------------ C ABI

extern "C" {

int ogrecu_skin_plan_create(ogrecu_ctx* ctx, const ogrecu_skin_job* job, ogrecu_skin_plan** out)
{
    OGRECU_REQUIRE(ctx && job && out, "NULL argument");
    *out = nullptr;
    const ogrecu_skin_job& j = *job;
    OGRECU_REQUIRE(j.num_weights >= 1 && j.num_weights <= 4, "num_weights must be 1..4");
    OGRECU_REQUIRE(j.num_matrices >= 1 && j.num_matrices <= 256, "num_matrices must be 1..256");
    OGRECU_REQUIRE(j.num_vertices >= 1 && j.blend_index && j.blend_weight, "empty blend stream");
    OGRECU_REQUIRE(j.blend_index_stride >= j.num_weights && j.blend_weight_stride >= 4 * j.num_weights, "blend strides smaller than the data");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    // the blend streams and the bone map are static mesh data: read them back once
    const size_t n = j.num_vertices;
    const size_t ib = (n - 1) * j.blend_index_stride + j.num_weights, wb = (n - 1) * j.blend_weight_stride + 4 * (size_t)j.num_weights;
    std::vector<uint8_t> hidx(ib), hw(wb);
    std::vector<uint16_t> hmap;
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    OGRECU_CUDA(cudaMemcpy(hidx.data(), j.blend_index, ib, cudaMemcpyDeviceToHost));
    OGRECU_CUDA(cudaMemcpy(hw.data(), j.blend_weight, wb, cudaMemcpyDeviceToHost));
    if (j.bone_map) {
        hmap.resize(j.num_matrices);
        OGRECU_CUDA(cudaMemcpy(hmap.data(), j.bone_map, (size_t)j.num_matrices * 2, cudaMemcpyDeviceToHost));
    }
    ogrecu_skin_plan* p = new ogrecu_skin_plan();
    memset(p, 0, sizeof *p);
    p->ctx = ctx;
    p->num_vertices = j.num_vertices; p->num_weights = j.num_weights; p->num_matrices = j.num_matrices;
    const int rc = build_skin_plan(ctx, hidx.data(), j.blend_index_stride, hw.data(), j.blend_weight_stride, j.bone_map ? hmap.data() : nullptr,
                                   j.num_weights, j.num_matrices, j.num_vertices, p);
    if (rc != OGRECU_OK) { ogrecu_skin_plan_destroy(p); return rc; }
    *out = p;
    return OGRECU_OK;
}

int ogrecu_skin_plan_destroy(ogrecu_skin_plan* plan)
{
    if (!plan) return OGRECU_OK;
    cudaSetDevice(plan->ctx->device);
    cudaStreamSynchronize(plan->ctx->stream);
    if (plan->d_recs) cudaFree(plan->d_recs);
    if (plan->d_chunks) cudaFree(plan->d_chunks);
    if (plan->d_pre_recs) cudaFree(plan->d_pre_recs);
    delete plan;
    return OGRECU_OK;
}

int ogrecu_skin_plan_info(const ogrecu_skin_plan* plan, uint32_t* num_chunks)
{
    OGRECU_REQUIRE(plan && num_chunks, "NULL argument");
    *num_chunks = plan->num_chunks;
    return OGRECU_OK;
}

int ogrecu_anim_frame_dev(ogrecu_ctx* ctx, const ogrecu_anim_job* jobs, ogrecu_skin_plan* const* plans, uint32_t num_jobs, const float* t_dev,
                          uint32_t num_instances, size_t palette_stride_floats)
{
    OGRECU_REQUIRE(ctx && (num_jobs == 0 || jobs), "NULL argument");
    if (num_jobs == 0 || num_instances == 0) return OGRECU_OK;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    const bool morph = t_dev != nullptr;
    // jobs the lane kernel takes go into launches of up to LN_MAX_JOBS, largest first; the rest take the tile kernels
    std::vector<uint32_t> lane, other;
    uint32_t lane_nmat = 0;
    for (uint32_t k = 0; k < num_jobs; ++k) {
        OGRECU_REQUIRE(morph ? jobs[k].key2 != nullptr : jobs[k].key2 == nullptr, "key2 and t_dev go together (both or neither)");
        if (jobs[k].skin.num_vertices == 0) continue;
        const bool ok = ctx->skin_mode == 0 && plans && skin_lane_eligible(plans[k], jobs[k], num_instances, palette_stride_floats) &&
                        (lane.empty() || jobs[k].skin.palette == jobs[lane[0]].skin.palette);
        (ok ? lane : other).push_back(k);
    }
    std::stable_sort(lane.begin(), lane.end(), [&](uint32_t a, uint32_t b) {
        return (uint64_t)jobs[a].skin.num_vertices * jobs[a].skin.num_weights > (uint64_t)jobs[b].skin.num_vertices * jobs[b].skin.num_weights;
    });
    if (!lane.empty()) {
        uint32_t nmat = 0;
        for (uint32_t k : lane) nmat = std::max(nmat, plans[k]->max_matrix + 1);
        if (!(ctx->lane_reuse_repack && ctx->lane_ipal.ptr))
            OGRECU_CHECK(repack_palettes(ctx, jobs[lane[0]].skin.palette, palette_stride_floats, num_instances, nmat));
        lane_nmat = nmat;
    }
    for (size_t at = 0; at < lane.size(); at += LN_MAX_JOBS) {
        const uint32_t cnt = (uint32_t)std::min<size_t>(LN_MAX_JOBS, lane.size() - at);
        ogrecu_anim_job jj[LN_MAX_JOBS];
        ogrecu_skin_plan* pp[LN_MAX_JOBS];
        for (uint32_t k = 0; k < cnt; ++k) { jj[k] = jobs[lane[at + k]]; pp[k] = plans[lane[at + k]]; }
        OGRECU_CHECK(launch_skin_lane(ctx, jj, pp, cnt, t_dev, num_instances, palette_stride_floats, morph, lane_nmat));
    }
    for (uint32_t k : other) {
        const ogrecu_anim_job& J = jobs[k];
        if (morph)
            OGRECU_CHECK(ogrecu_anim_instanced_dev(ctx, &J.skin, J.key2, t_dev, J.morph_out, J.morph_instance_stride_bytes, num_instances,
                                                   palette_stride_floats, J.dst_instance_stride_bytes));
        else
            OGRECU_CHECK(ogrecu_skin_instanced_dev(ctx, &J.skin, num_instances, palette_stride_floats, 0, J.dst_instance_stride_bytes));
    }
    return OGRECU_OK;
}

#ifdef LN_PROFILE
// development builds only: the phase cycle counts of the last lane launch (see LN_PROF)
__attribute__((visibility("default"))) int ogrecu_debug_lane_profile(ogrecu_ctx* ctx, unsigned long long* out16)
{
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    OGRECU_CUDA(cudaMemcpy(out16, (const uint8_t*)ctx->lane_counter.ptr + 64, 128, cudaMemcpyDeviceToHost));
    return OGRECU_OK;
}
#endif

}  // extern "C"
