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

// sqrt then 1/x of both halves, correctly rounded (== __fsqrt_rn / __frcp_rn); the packed fast sequences when the whole
// warp is in their range (always, for unit-scale normals), the library ones otherwise
__device__ __forceinline__ void ln_sqrt_rcp(u64 l2, u64 nz, u64& len, u64& inv)
{
    float a, b;
    ln_upk(l2, a, b);
    const bool fast = em_fast_range(__float_as_uint(a)) && em_fast_range(__float_as_uint(b));
    if (__all_sync(0xffffffffu, fast)) {
        len = em_sqrt_fast2(l2, nz);
        inv = em_rcp_fast2(len);
    } else {
        const float la = __fsqrt_rn(a), lb = __fsqrt_rn(b);
        len = ln_pk(la, lb);
        inv = ln_pk(__frcp_rn(la), __frcp_rn(lb));
    }
}

// One half (even or odd instances) of the output stage.  A quad = 4 consecutive vertices = one 96-byte row per instance =
// one TMA box of 32 rows (48-byte rows run the TMA unit at a quarter of the HBM rate, 96-byte rows at all of it:
// tools/ubench/store_pattern.cu).  A box must start on a 16-byte boundary: quads of the even instances start at
// vertices 4q; the odd instances start 8 mod 16 when the instance stride is 8 mod 16 (an odd vertex count), their quads
// then start at vertices 4q - 1 (phi = 1).  Vertices no whole in-range quad covers (the first three of a phi = 1 row, the
// last n mod 4) are stored directly, 24 bytes per lane.
struct LaneHalf {
    uint8_t* buf;                   // this warp's staging rows of the half: [32][96 B]
    const CUtensorMap* map;
    uint8_t* dst;                   // this lane's instance of the half (NULL = past the last instance)
    int32_t v0, v1;                 // chunk-local vertices this warp stores for the half
    int32_t phi, xoff;              // quad phase; word offset of vertex 0 in the map's rows
};

template <bool HI>
__device__ __forceinline__ void ln_emit(const LaneHalf& H, const int32_t v, const int32_t gv, const int32_t nverts, const int32_t row0,
                                        const uint32_t lane, const u64 (&Oprev)[6], const u64 (&O)[6])
{
    if (v < H.v0 || v >= H.v1) return;
    const int32_t j = (gv + H.phi) & 3, qs = gv - j;
    if (qs >= 0 && qs + 4 <= nverts) {
        if (j & 1) {
            if (j == 1) {                           // the buffer's previous box has been read (issued a quad ago)
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
            float f[12];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                float lo, hi;
                ln_upk(Oprev[k], lo, hi); f[k] = HI ? hi : lo;
                ln_upk(O[k], lo, hi); f[6 + k] = HI ? hi : lo;
            }
            float4* r = reinterpret_cast<float4*>(H.buf + lane * 96u + (uint32_t)(j >> 1) * 48u);
            r[0] = make_float4(f[0], f[1], f[2], f[3]); r[1] = make_float4(f[4], f[5], f[6], f[7]); r[2] = make_float4(f[8], f[9], f[10], f[11]);
            if (j == 3) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // my STS visible to the async proxy
                __syncwarp();
#ifndef LN_NO_STORE   // (timing experiments only: the arithmetic without the output stage)
                if (lane == 0) {
                    ln_tma_store_2d(H.map, ln_smem_u32(H.buf), qs * 6 + H.xoff, row0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
#endif
            }
        }
    } else if (H.dst) {
        float2* d = reinterpret_cast<float2*>(H.dst + (size_t)gv * 24);
        float f[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { float lo, hi; ln_upk(O[k], lo, hi); f[k] = HI ? hi : lo; }
        d[0] = make_float2(f[0], f[1]); d[1] = make_float2(f[2], f[3]); d[2] = make_float2(f[4], f[5]);
    }
}

struct LaneWalk {
    const uint8_t* s_pal_lane;      // palette base + lane * 16
    const uint4* s_rec;             // [1 + chunk vertices][2]; entry 0 = the vertex before the chunk
    const float4* s_key;            // [1 + chunk vertices][3 (k1 6, k2 - k1 6) or, without morph, 1.5 -> 2 (k1 6, pad)]
    LaneHalf A, B;
    int32_t v_begin, nverts;
    int32_t vfirst, vend;           // chunk-local vertex range walked (vfirst may be -1: the vertex before the chunk)
    int32_t row0;
    uint32_t lane;
};

// The chunk walk of one warp: vertices vfirst..vend-1 in order, every value as the (A, B) pair of the lane's two instances.
template <int NW, bool MORPH>
__device__ __forceinline__ void lane_walk(const LaneWalk& K, const u64 T, const u64 NZ)
{
    u64 E[12], Oprev[6];
#pragma unroll
    for (int k = 0; k < 12; ++k) E[k] = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) Oprev[k] = 0;
    uint32_t cur = 0xffffffffu;                       // slot offset whose matrix E is a plain copy of, or none
    const uint32_t lane = K.lane;
#pragma unroll 1
    for (int32_t v = K.vfirst; v < K.vend; ++v) {
        const int32_t gv = K.v_begin + v;                               // mesh vertex (>= 0: chunk 0 never walks its -1)
        u64 O[6];
        const uint32_t vl = (uint32_t)(v + 1);
        const uint4 r0 = K.s_rec[vl * 2u];
        const float4* ky = K.s_key + vl * (MORPH ? 3u : 2u);
        u64 V[6];
        {
            const float4 a = ky[0], b = ky[1];
            V[0] = ln_pk(a.x, a.x); V[1] = ln_pk(a.y, a.y); V[2] = ln_pk(a.z, a.z);     // warp-uniform scalars: ptxas folds the
            V[3] = ln_pk(a.w, a.w); V[4] = ln_pk(b.x, b.x); V[5] = ln_pk(b.y, b.y);     // duplication into .F32 broadcast operands
            if (MORPH) {
                // dst = a + t*(b-a) (OgreOptimisedUtilGeneral.cpp:249-256); (b - a) is instance-invariant
                const float4 c = ky[2];
                V[0] = ln_add(V[0], ln_mul(T, ln_pk(b.z, b.z), NZ)); V[1] = ln_add(V[1], ln_mul(T, ln_pk(b.w, b.w), NZ));
                V[2] = ln_add(V[2], ln_mul(T, ln_pk(c.x, c.x), NZ)); V[3] = ln_add(V[3], ln_mul(T, ln_pk(c.y, c.y), NZ));
                V[4] = ln_add(V[4], ln_mul(T, ln_pk(c.z, c.z), NZ)); V[5] = ln_add(V[5], ln_mul(T, ln_pk(c.w, c.w), NZ));
                // Vector3::normalise (OgreVector.h:494-508): scale by 1/len only when len > 0
                const u64 l2 = ln_add(ln_add(ln_mul(V[3], V[3], NZ), ln_mul(V[4], V[4], NZ)), ln_mul(V[5], V[5], NZ));
                u64 len, inv;
                ln_sqrt_rcp(l2, NZ, len, inv);
                float la, lb, ia, ib;
                ln_upk(len, la, lb); ln_upk(inv, ia, ib);
                const u64 I = ln_pk(la > 0.0f ? ia : 1.0f, lb > 0.0f ? ib : 1.0f);     // x * 1.0f == x
                V[3] = ln_mul(V[3], I, NZ); V[4] = ln_mul(V[4], I, NZ); V[5] = ln_mul(V[5], I, NZ);
            }
        }
        // ---- collapse: E[4*r + c] = (M_A[r][c], M_B[r][c]); control flow is warp-uniform (plan flags).
        // One load site for slot 0 (E keeps its registers across the paths); a single-matrix vertex whose matrix E
        // already holds loads nothing.
        const uint32_t fl = r0.x & 15u, off0 = r0.x & ~15u;
        const bool single = NW == 1 || (NW == 2 ? !(fl & 2u) : fl == 1u);
        if (!(single && off0 == cur)) {
            const ulonglong2* m0 = reinterpret_cast<const ulonglong2*>(K.s_pal_lane + off0);
#pragma unroll
            for (int k = 0; k < 6; ++k) { const ulonglong2 q = m0[k * 32]; E[2 * k] = q.x; E[2 * k + 1] = q.y; }
        }
        cur = single ? off0 : 0xffffffffu;
        if (NW >= 2 && !single) {
            const float4 w = *reinterpret_cast<const float4*>(&K.s_rec[vl * 2u + 1u]);
            if (NW == 2) {
                const ulonglong2* m1 = reinterpret_cast<const ulonglong2*>(K.s_pal_lane + r0.y);
                const u64 W1 = ln_pk(w.y, w.y);
#pragma unroll
                for (int k = 0; k < 6; ++k) {                           // lerp(M0, M1, w1) = (M1 - M0) * w1 + M0
                    const ulonglong2 b = m1[k * 32];
                    E[2 * k] = ln_add(ln_mul(ln_sub(b.x, E[2 * k]), W1, NZ), E[2 * k]);
                    E[2 * k + 1] = ln_add(ln_mul(ln_sub(b.y, E[2 * k + 1]), W1, NZ), E[2 * k + 1]);
                }
            } else {
                if (!(fl & 1u)) {                                       // w0 == 1.0f: 1 * x == x, nothing to do
                    const u64 W0 = ln_pk(w.x, w.x);
#pragma unroll
                    for (int k = 0; k < 12; ++k) E[k] = ln_mul(E[k], W0, NZ);
                }
                if (fl & 2u) {
                    const ulonglong2* mb = reinterpret_cast<const ulonglong2*>(K.s_pal_lane + r0.y);
                    const u64 Wb = ln_pk(w.y, w.y);
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const ulonglong2 q = mb[k * 32];
                        E[2 * k] = ln_add(ln_mul(q.x, Wb, NZ), E[2 * k]);
                        E[2 * k + 1] = ln_add(ln_mul(q.y, Wb, NZ), E[2 * k + 1]);
                    }
                }
                if (fl & 4u) {
                    const ulonglong2* mb = reinterpret_cast<const ulonglong2*>(K.s_pal_lane + r0.z);
                    const u64 Wb = ln_pk(w.z, w.z);
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const ulonglong2 q = mb[k * 32];
                        E[2 * k] = ln_add(ln_mul(q.x, Wb, NZ), E[2 * k]);
                        E[2 * k + 1] = ln_add(ln_mul(q.y, Wb, NZ), E[2 * k + 1]);
                    }
                }
                if (NW == 4 && (fl & 8u)) {
                    const ulonglong2* mb = reinterpret_cast<const ulonglong2*>(K.s_pal_lane + r0.w);
                    const u64 Wb = ln_pk(w.w, w.w);
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const ulonglong2 q = mb[k * 32];
                        E[2 * k] = ln_add(ln_mul(q.x, Wb, NZ), E[2 * k]);
                        E[2 * k + 1] = ln_add(ln_mul(q.y, Wb, NZ), E[2 * k + 1]);
                    }
                }
            }
        }
        // ---- position: (r.x*x + r.y*y) + (r.z*z + r.w)   (__MM_DOT4x3_PS, OgreSIMDHelper.h:204-217)
#pragma unroll
        for (int r = 0; r < 3; ++r)
            O[r] = ln_add(ln_add(ln_mul(E[4 * r], V[0], NZ), ln_mul(E[4 * r + 1], V[1], NZ)), ln_add(ln_mul(E[4 * r + 2], V[2], NZ), E[4 * r + 3]));
        // ---- normal: (r.x*x + r.y*y) + r.z*z (OgreSIMDHelper.h:220-227), then exact 1/sqrt(len2)
        u64 N[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            N[r] = ln_add(ln_add(ln_mul(E[4 * r], V[3], NZ), ln_mul(E[4 * r + 1], V[4], NZ)), ln_mul(E[4 * r + 2], V[5], NZ));
        const u64 l2 = ln_add(ln_add(ln_mul(N[0], N[0], NZ), ln_mul(N[1], N[1], NZ)), ln_mul(N[2], N[2], NZ));
        u64 len, R;
        ln_sqrt_rcp(l2, NZ, len, R);
        O[3] = ln_mul(N[0], R, NZ); O[4] = ln_mul(N[1], R, NZ); O[5] = ln_mul(N[2], R, NZ);
        // ---- output
        ln_emit<false>(K.A, v, gv, K.nverts, K.row0, lane, Oprev, O);
        ln_emit<true>(K.B, v, gv, K.nverts, K.row0, lane, Oprev, O);
#pragma unroll
        for (int k = 0; k < 6; ++k) Oprev[k] = O[k];
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // shared memory must outlive the reads
    __syncwarp();
}

template <bool MORPH>
__global__ void __launch_bounds__(LN_THREADS, 2) skin_lane_kernel(const __grid_constant__ LaneParams P, const u64 NZ)
{
    constexpr int KEY_BYTES = MORPH ? 48 : 32;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_stage = smem;                                                       // [warps][A, B][32 x 96 B], 128-byte aligned boxes
    uint8_t* s_pal = s_stage + LN_WARPS * 2 * LN_BOX_BYTES;                        // [LN_MAXB][6][32] x 16 B
    uint4* s_rec = reinterpret_cast<uint4*>(s_pal + LN_MAXB * LN_SLOT_BYTES);      // [LN_CHUNK_VERTS + 1][2]
    float4* s_key = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(s_rec) + (LN_CHUNK_VERTS + 1) * LN_REC_BYTES);
    uint32_t* s_dirty = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(s_key) + (LN_CHUNK_VERTS + 1) * KEY_BYTES);   // [LN_MAXB + 1]

    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    // ---- which (job, chunk, instance group) this CTA is
    uint32_t j = 0;
#pragma unroll 1
    while (j + 1 < P.njobs && blockIdx.x >= P.job[j + 1].first_item) ++j;
    const LaneJob& J = P.job[j];
    const uint32_t local = blockIdx.x - J.first_item;
    const uint32_t c = local / P.groups, g = local - c * P.groups;
    const uint4* chrec = J.chunks + (size_t)c * 3;         // {v_begin, nquads, nslots, -}, 16 x u16 palette matrix of each slot
    const uint4 ch = __ldg(chrec);
    const uint32_t v_begin = ch.x, nquads = ch.y, nslots = ch.z;
    const uint32_t nv_chunk = min(4u * nquads, J.nverts - v_begin);
    const uint32_t i0 = g * 64u;

    // ---- palettes of the 64 instances, the chunk's bones only -> [slot][chunk][lane] (A.e, B.e, A.e+1, B.e+1).
    // Task e = (slot, instance pair, row): three threads read one matrix's 48 bytes.  All loads of a thread first.
    constexpr int PG = (LN_MAXB * 96 + LN_THREADS - 1) / LN_THREADS;
    float4 PA[PG], PB[PG];
    {
        const uint16_t* slots = reinterpret_cast<const uint16_t*>(chrec + 1);
#pragma unroll
        for (int k = 0; k < PG; ++k) {
            const uint32_t e = tid + LN_THREADS * k;
            if (e < nslots * 96u) {
                const uint32_t s = e / 96u, q = e - s * 96u, l = q / 3u, r = q - l * 3u;
                const uint32_t gm = (uint32_t)__ldg(slots + s);
                const uint32_t ia = min(i0 + 2u * l, P.ninst - 1u), ib = min(i0 + 2u * l + 1u, P.ninst - 1u);     // lanes past the end repeat the last instance
                PA[k] = __ldg(reinterpret_cast<const float4*>(P.palette + (size_t)ia * P.pal_stride_floats + gm * 16u + r * 4u));
                PB[k] = __ldg(reinterpret_cast<const float4*>(P.palette + (size_t)ib * P.pal_stride_floats + gm * 16u + r * 4u));
            }
        }
    }
    if (tid <= LN_MAXB) s_dirty[tid] = 0;
    // raw plan record and keyframes of my vertex (thread 0: the vertex before the chunk, thread 1 + i: chunk vertex i)
    uint4 rec0 = make_uint4(0, 0, 0, 0), rec1 = rec0;
    float k1[6], k2[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { k1[k] = 0.f; k2[k] = 0.f; }
    const bool has_v = tid == 0 ? v_begin > 0 : tid <= nv_chunk;
    if (has_v) {
        const uint32_t gv = v_begin + tid - 1u;
        const uint4* r = tid == 0 ? J.pre_recs + (size_t)c * 2 : J.recs + (size_t)gv * 2;
        rec0 = __ldg(r); rec1 = __ldg(r + 1);
        const float* a = reinterpret_cast<const float*>(J.key1 + (size_t)gv * 24);
#pragma unroll
        for (int k = 0; k < 6; ++k) k1[k] = __ldg(a + k);
        if (MORPH) {
            const float* b = reinterpret_cast<const float*>(J.key2 + (size_t)gv * 24);
#pragma unroll
            for (int k = 0; k < 6; ++k) k2[k] = __ldg(b + k);
        }
    }
    __syncthreads();                                       // s_dirty cleared
    {
#pragma unroll
        for (int k = 0; k < PG; ++k) {
            const uint32_t e = tid + LN_THREADS * k;
            if (e < nslots * 96u) {
                const uint32_t s = e / 96u, q = e - s * 96u, l = q / 3u, r = q - l * 3u;
                const float4 A = PA[k], B = PB[k];
                const uint32_t u0 = __float_as_uint(A.x), u1 = __float_as_uint(A.y), u2 = __float_as_uint(A.z), u3 = __float_as_uint(A.w);
                const uint32_t u4 = __float_as_uint(B.x), u5 = __float_as_uint(B.y), u6 = __float_as_uint(B.z), u7 = __float_as_uint(B.w);
                if (ln_neg_zero_or_denormal(u0) || ln_neg_zero_or_denormal(u1) || ln_neg_zero_or_denormal(u2) || ln_neg_zero_or_denormal(u3) ||
                    ln_neg_zero_or_denormal(u4) || ln_neg_zero_or_denormal(u5) || ln_neg_zero_or_denormal(u6) || ln_neg_zero_or_denormal(u7)) s_dirty[s] = 1;
                if (ln_huge(u0) || ln_huge(u1) || ln_huge(u2) || ln_huge(u3) || ln_huge(u4) || ln_huge(u5) || ln_huge(u6) || ln_huge(u7)) s_dirty[LN_MAXB] = 1;
                float4* d = reinterpret_cast<float4*>(s_pal + s * LN_SLOT_BYTES + (2u * r) * 512u + l * 16u);
                d[0] = make_float4(A.x, B.x, A.y, B.y);
                d[32] = make_float4(A.z, B.z, A.w, B.w);
            }
        }
    }
    __syncthreads();                                       // palettes and exactness flags complete
    if (has_v) {
        // a vertex whose slot-0 matrix holds a -0 in some instance, or any vertex when something is not finite, takes every slot
        const uint32_t slot0 = (rec0.x & ~15u) / LN_SLOT_BYTES;
        if (s_dirty[LN_MAXB] | s_dirty[slot0]) rec0.x = (rec0.x & ~15u) | (J.nw >= 3 ? (J.nw == 3 ? 6u : 14u) : 2u);
        s_rec[tid * 2u] = rec0; s_rec[tid * 2u + 1u] = rec1;
        float* ky = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_key) + tid * KEY_BYTES);
#pragma unroll
        for (int k = 0; k < 6; ++k) ky[k] = k1[k];
        if (MORPH) {
#pragma unroll
            for (int k = 0; k < 6; ++k) ky[6 + k] = fsub(k2[k], k1[k]);
        }
    }
    __syncthreads();                                       // the last CTA barrier: from here on the warps run independently

    u64 T = 0;
    if (MORPH) {
        const uint32_t ia = min(i0 + 2u * lane, P.ninst - 1u), ib = min(i0 + 2u * lane + 1u, P.ninst - 1u);
        T = ln_pk(__ldg(P.t + ia), __ldg(P.t + ib));
    }
    // contiguous runs of quads per warp (consecutive vertices share bones: the matrix stays in registers)
    const uint32_t per = (nquads + LN_WARPS - 1u) / LN_WARPS;
    const uint32_t q0 = min(wid * per, nquads), q1 = min(q0 + per, nquads);
    if (q0 >= q1) return;
    const int32_t r0 = (int32_t)(4u * q0), r1 = (int32_t)min(4u * q1, nv_chunk);
    const bool last_run = v_begin + (uint32_t)r1 == J.nverts;
    const int32_t shift = (int32_t)J.shift;
    LaneWalk K;
    K.s_pal_lane = s_pal + lane * 16u;
    K.s_rec = s_rec; K.s_key = s_key;
    K.A.buf = s_stage + wid * (2 * LN_BOX_BYTES);
    K.B.buf = K.A.buf + LN_BOX_BYTES;
    K.A.map = &J.mapA; K.B.map = &J.mapB;
    K.A.dst = i0 + 2u * lane < P.ninst ? J.dst + (size_t)(i0 + 2u * lane) * J.dst_stride : nullptr;
    K.B.dst = i0 + 2u * lane + 1u < P.ninst ? J.dst + (size_t)(i0 + 2u * lane + 1u) * J.dst_stride : nullptr;
    K.A.phi = 0; K.A.xoff = 0;
    K.B.phi = shift; K.B.xoff = 2 * shift;                 // shifted: map B starts 8 bytes before the odd instance
    K.A.v0 = r0; K.A.v1 = r1;
    // the shifted half stores the quads that START in this run: vertices r0-1 .. r1-2, and the mesh's last run its last vertex too
    K.B.v0 = (v_begin == 0 && r0 == 0) ? 0 : r0 - shift;
    K.B.v1 = last_run ? r1 : r1 - shift;
    K.v_begin = (int32_t)v_begin; K.nverts = (int32_t)J.nverts;
    K.vfirst = K.B.v0 < r0 ? K.B.v0 : r0;
    K.vend = r1;
    K.row0 = (int32_t)(i0 >> 1);
    K.lane = lane;
    switch (J.nw) {
    case 1: lane_walk<1, MORPH>(K, T, NZ); break;
    case 2: lane_walk<2, MORPH>(K, T, NZ); break;
    case 3: lane_walk<3, MORPH>(K, T, NZ); break;
    default: lane_walk<4, MORPH>(K, T, NZ); break;
    }
}

// ------------------------------------------------------------------------------------------------ host side

static size_t lane_smem(bool morph)
{
    return (size_t)LN_WARPS * 2 * LN_BOX_BYTES + (size_t)LN_MAXB * LN_SLOT_BYTES + (size_t)(LN_CHUNK_VERTS + 1) * LN_REC_BYTES +
           (size_t)(LN_CHUNK_VERTS + 1) * (morph ? 48 : 32) + (LN_MAXB + 1) * 4 + 16;
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

int launch_skin_lane(ogrecu_ctx* ctx, const ogrecu_anim_job* jobs, ogrecu_skin_plan* const* plans, uint32_t num_jobs, const float* t_dev,
                     uint32_t num_instances, size_t palette_stride_floats, bool morph)
{
    OGRECU_REQUIRE(num_jobs >= 1 && num_jobs <= LN_MAX_JOBS, "1..8 jobs per lane launch");
    LaneParams P;
    memset(&P, 0, sizeof P);
    P.njobs = num_jobs;
    P.ninst = num_instances;
    P.groups = div_up(num_instances, 64);
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
        // and the kernel pairs their vertices (2p-1, 2p) so that every box still starts on a 16-byte boundary
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
    if (morph) skin_lane_kernel<true><<<(unsigned)items, LN_THREADS, smem, ctx->stream>>>(P, 0x8000000080000000ull);
    else skin_lane_kernel<false><<<(unsigned)items, LN_THREADS, smem, ctx->stream>>>(P, 0x8000000080000000ull);
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
    while (v < nverts) {
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
        while (c.nquads < LN_CHUNK_QUADS && v < nverts) {
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
                break;
            }
            for (uint32_t q = 0; q < nadd; ++q) add_bone(add[q]);
            for (uint32_t u = v; u < v + 4 && u < nverts; ++u) make_rec(u, &recs[(size_t)u * 8]);
            v = std::min(v + 4, nverts);
            c.nquads++;
        }
        c.nslots = nlocal;
        pre_recs.resize(pre_recs.size() + 8, 0);
        if (c.v_begin > 0) make_rec(c.v_begin - 1, &pre_recs[pre_recs.size() - 8]);
        chunks.push_back(c);
    }
    plan->num_chunks = (uint32_t)chunks.size();
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
    for (size_t at = 0; at < lane.size(); at += LN_MAX_JOBS) {
        const uint32_t cnt = (uint32_t)std::min<size_t>(LN_MAX_JOBS, lane.size() - at);
        ogrecu_anim_job jj[LN_MAX_JOBS];
        ogrecu_skin_plan* pp[LN_MAX_JOBS];
        for (uint32_t k = 0; k < cnt; ++k) { jj[k] = jobs[lane[at + k]]; pp[k] = plans[lane[at + k]]; }
        OGRECU_CHECK(launch_skin_lane(ctx, jj, pp, cnt, t_dev, num_instances, palette_stride_floats, morph));
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

}  // extern "C"
