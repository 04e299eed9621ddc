// skin_pair.cu — tiled, instanced software skinning (and fused morph+skin), two instances per pass with
// packed-f32x2 arithmetic.  Successor of skin_tile_kernel (skin_tile.cu) for calls with >= 2 instances.
//
// Reference work replaced: Mesh::softwareVertexMorph -> OptimisedUtil::softwareVertexMorph and
// Mesh::softwareVertexBlend -> OptimisedUtil::softwareVertexSkinning (OgreMain/src/OgreMesh.cpp:1792-1946,
// arithmetic OgreOptimisedUtilSSE.cpp:278-421,606-761 and OgreSIMDHelper.h:204-227) for many instances of a mesh.
//
// Why: ncu on skin_tile_kernel (profiles/r01b_*) showed the issue slots 72 % busy with DRAM at 22-60 %: ~315
// instructions per vertex, of which ~160 are the un-fused FMUL/FADD the bit-exact replay of SSE needs (no FFMA)
// and the rest is per-iteration overhead (barriers, copy issue, address arithmetic).  Blackwell has packed
// FP32 instructions (FMUL2/FADD2/FFMA2, two lanes of a 64-bit register pair per issue slot; measured in
// tools/ubench/fp32x2.cu: same lane throughput, half the issue slots).  Here a thread owns ONE vertex of the
// tile for TWO instances A and B and keeps every value as the pair (A, B):
//   * the palettes of the two instances are gathered into shared memory INTERLEAVED element by element
//     (4-byte cp.async), so one LDS.128 returns two ready-made (A, B) pairs and no register shuffling is needed;
//   * collapse / transform / normalise run on pairs: one instruction does the work for both instances;
//   * barriers, copy issue and address arithmetic are paid once per two instances.
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false, which would change the
// rounding; the packed multiply is therefore issued as fma.rn.f32x2(a, b, -0.0) with the -0.0 pair coming from
// a kernel argument (opaque to ptxas): a*b + (-0) rounds once and equals RN(a*b) in every case including
// signed zeros, so the result is bit-identical to the scalar __fmul_rn / __fadd_rn sequence of skin_math.cuh.
#include "common.cuh"
#include "skin_math.cuh"
#include "skin_tile.cuh"

using namespace ogrecu;

namespace ogrecu {

// 128-thread CTAs, six per SM: the same 24 warps per SM as 256 x 3 (registers bound both), but the vertex tiles
// quantise Sinbad's sub-meshes better (94 % of the lanes carry a vertex instead of 86 %) and a barrier waits for four
// warps instead of eight: 87.6 -> 90.8 Gvert/s on the bench step.
#ifndef SP_THREADS_N
#define SP_THREADS_N 128
#endif
constexpr int SP_THREADS = SP_THREADS_N;
constexpr int SP_PQ = 768 / SP_THREADS;       // palette elements per thread kept as a register plan (covers 64 matrices)
constexpr int SP_MSTRIDE = 28;        // floats per matrix pair in shared memory: 2 x 12 used + 4 pad (112 B: the
                                      // 16-byte chunks of 8 different matrices fall into 8 different bank groups)
#ifndef SP_MIN_CTAS
#define SP_MIN_CTAS 6
#endif

typedef unsigned long long u64;

__device__ __forceinline__ uint32_t sp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// RN(a*b) per lane, see the header comment for why it is an fma with a -0.0 addend
__device__ __forceinline__ u64 pmul(u64 a, u64 b, u64 nz) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(nz)); return d; }
__device__ __forceinline__ u64 padd(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 psub(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

__device__ __forceinline__ void sp_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void sp_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sp_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void sp_tma_load(uint32_t dst_smem, const void* gsrc, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void sp_tma_store(void* gdst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(gsrc) : "memory");
}

// exact 1/sqrt(len2) as skin_math.cuh::xform_norm takes it: correctly rounded sqrt, then correctly rounded 1/x
__device__ __forceinline__ float rsq_exact(float len2) { return __frcp_rn(__fsqrt_rn(len2)); }

template <int NW, bool NORM, bool FUSED>
__global__ void __launch_bounds__(SP_THREADS, SP_MIN_CTAS) skin_pair_kernel(const __grid_constant__ SkinTileParams P, const u64 NZ)
{
    constexpr int VS = NORM ? 24 : 12;              // bytes per vertex
    constexpr int NC = NORM ? 6 : 3;                // floats per vertex
    constexpr int TILE_BYTES = SP_THREADS * VS;
    constexpr int MAXU4 = TILE_BYTES / (SP_THREADS * 4);   // 4-byte units per thread and tile (6 or 3)

    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t nelem = P.num_matrices * 12;                         // palette floats used per instance
    const uint32_t pal_stage_bytes = P.num_matrices * (SP_MSTRIDE * 4);
    uint8_t* s_dst = smem;                                              // [2 stages][2 instances][TILE_BYTES]
    uint8_t* s_src = s_dst + 4 * TILE_BYTES;                            // [2][2][TILE_BYTES]  (absent when FUSED)
    uint8_t* s_pal = s_src + (FUSED ? 0 : 4 * TILE_BYTES);              // [2][num_matrices][28 floats]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_pal + 2 * pal_stage_bytes);   // [2]
    float* s_t = reinterpret_cast<float*>(s_bar + 2);                   // [inst_per_cta] morph weights of this CTA's instances (FUSED)

    const uint32_t tid = threadIdx.x;
    const uint32_t v0 = blockIdx.x * SP_THREADS;
    const uint32_t nvalid = min((uint32_t)SP_THREADS, P.num_vertices - v0);
    const uint32_t bytes = nvalid * VS;
    const uint32_t i0 = blockIdx.y * P.inst_per_cta;
    const uint32_t cnt = min(P.inst_per_cta, P.num_instances - i0);
    const uint32_t npairs = (cnt + 1) >> 1;
    const bool active = tid < nvalid;
    const size_t tile_off = (size_t)v0 * VS;
    const bool src_shared = !FUSED && P.src_inst_stride == 0;           // one bind pose for every instance
    const bool tma_ld = !FUSED && P.tma_load && (bytes & 15u) == 0;
    const bool tma_st = P.tma_store && (bytes & 15u) == 0;

    if (tma_ld && tid == 0) {
        sp_mbar_init(sp_smem_u32(&s_bar[0]), 1);
        sp_mbar_init(sp_smem_u32(&s_bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }

    // ---- per-vertex constants of this tile (shared by all instances): bone row offsets, weight pairs, keyframes
    uint32_t boff[NW];
    u64 W[NW];
    float k1[NC], kd[NC];
#pragma unroll
    for (int b = 0; b < NW; ++b) { boff[b] = 0; W[b] = 0; }
#pragma unroll
    for (int k = 0; k < NC; ++k) { k1[k] = 0.f; kd[k] = 0.f; }
    if (active) {
        const uint8_t* bi = P.blend_index + (size_t)(v0 + tid) * P.blend_index_stride;
        const float* bw = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(P.blend_weight) + (size_t)(v0 + tid) * P.blend_weight_stride);
#pragma unroll
        for (int b = 0; b < NW; ++b) {
            boff[b] = (uint32_t)__ldg(bi + b) * (SP_MSTRIDE * 4);
            const float w = __ldg(bw + b);
            W[b] = pk(w, w);
        }
        if (FUSED) {
            const float* a = reinterpret_cast<const float*>(P.src + tile_off + (size_t)tid * VS);
            const float* b = reinterpret_cast<const float*>(P.key2 + tile_off + (size_t)tid * VS);
#pragma unroll
            for (int k = 0; k < NC; ++k) { k1[k] = __ldg(a + k); kd[k] = fsub(__ldg(b + k), k1[k]); }   // (b - a) is instance-invariant
        }
    }
    // palette gather plan: element e = tid + THREADS*q of the instance palette (matrix e/12, float e%12 of its first
    // three rows, contiguous in Affine3) -> global byte offset | shared byte offset << 16; 0xffffffff = no element.
    // (Measured and rejected: lane pairs copying the A and the B float of one element — conflict-free shared-memory
    // writes, but twice the plan registers; 7 % slower.)
    uint32_t plan[SP_PQ];
#pragma unroll
    for (int q = 0; q < SP_PQ; ++q) {
        const uint32_t e = tid + SP_THREADS * q;
        plan[q] = 0xffffffffu;
        if (e < nelem) {
            const uint32_t m = e / 12, k = e - m * 12;
            const uint32_t sm = P.bone_map ? (uint32_t)__ldg(P.bone_map + m) : m;
            plan[q] = ((sm * 16 + k) * 4) | ((m * (SP_MSTRIDE * 4) + 8 * k) << 16);
        }
    }
    const size_t pal_stride_bytes = P.palette_stride_floats * 4;
    __syncthreads();   // mbarrier init visible

    auto issue_loads = [&](uint32_t stage, uint32_t instA, bool hasB, bool want_src) {
        const uint8_t* palA = reinterpret_cast<const uint8_t*>(P.palette) + (size_t)instA * pal_stride_bytes;
        const uint8_t* palB = hasB ? palA + pal_stride_bytes : palA;
        const uint32_t dpal = sp_smem_u32(s_pal + stage * pal_stage_bytes);
#pragma unroll
        for (int q = 0; q < SP_PQ; ++q) {
            if (plan[q] != 0xffffffffu) {
                const uint32_t g = plan[q] & 0xffffu, d = dpal + (plan[q] >> 16);
                cp_async4(d, palA + g);
                cp_async4(d + 4, palB + g);
            }
        }
        if (nelem > SP_PQ * SP_THREADS) {                                               // palettes of more than 64 matrices
            for (uint32_t e = tid + SP_THREADS * SP_PQ; e < nelem; e += SP_THREADS) {
                const uint32_t m = e / 12, k = e - m * 12;
                const uint32_t sm = P.bone_map ? (uint32_t)__ldg(P.bone_map + m) : m;
                const uint32_t g = (sm * 16 + k) * 4, d = dpal + m * (SP_MSTRIDE * 4) + 8 * k;
                cp_async4(d, palA + g);
                cp_async4(d + 4, palB + g);
            }
        }
        if (!FUSED && want_src) {
            const uint8_t* gA = P.src + (size_t)instA * P.src_inst_stride + tile_off;
            const uint8_t* gB = hasB ? gA + P.src_inst_stride : gA;
            const uint32_t dA = sp_smem_u32(s_src + (stage * 2) * TILE_BYTES), dB = dA + TILE_BYTES;
            if (tma_ld) {
                if (tid == 0) {
                    const uint32_t bar = sp_smem_u32(&s_bar[stage]);
                    sp_mbar_expect_tx(bar, hasB ? 2 * bytes : bytes);
                    sp_tma_load(dA, gA, bytes, bar);
                    if (hasB) sp_tma_load(dB, gB, bytes, bar);
                }
            } else if (P.load_unit == 8) {
                const uint32_t nu = bytes >> 3;
#pragma unroll
                for (int q = 0; q < MAXU4 / 2; ++q) {
                    const uint32_t u = tid + SP_THREADS * q;
                    if (u < nu) { cp_async8(dA + u * 8, gA + (size_t)u * 8); if (hasB) cp_async8(dB + u * 8, gB + (size_t)u * 8); }
                }
            } else {
                const uint32_t nu = bytes >> 2;
#pragma unroll
                for (int q = 0; q < MAXU4; ++q) {
                    const uint32_t u = tid + SP_THREADS * q;
                    if (u < nu) { cp_async4(dA + u * 4, gA + (size_t)u * 4); if (hasB) cp_async4(dB + u * 4, gB + (size_t)u * 4); }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if (cnt == 0) return;
    issue_loads(0, i0, cnt > 1, true);
    // the morph weights of all this CTA's instances, staged once (a per-pass __ldg of the next pair's weights showed up
    // as 7 % long-scoreboard stalls in ncu although it was issued a whole pass ahead)
    if (FUSED) {
        for (uint32_t k = tid; k < cnt; k += SP_THREADS) s_t[k] = __ldg(P.tarr + i0 + k);
    }

    for (uint32_t j = 0; j < npairs; ++j) {
        const uint32_t s = j & 1u;
        const uint32_t instA = i0 + 2 * j;
        const bool hasB = 2 * j + 1 < cnt;
        const uint32_t ss = src_shared ? 0u : s;                      // source stage
        asm volatile("cp.async.wait_group 0;" ::: "memory");         // my cp.async pieces of stage s have landed
        if (tma_ld && (!src_shared || j == 0)) sp_mbar_wait(sp_smem_u32(&s_bar[ss]), src_shared ? 0u : ((j >> 1) & 1u));
        if (tma_st && tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // dst[s] of pass j-2 drained
        __syncthreads();                                              // (A) stage s visible; stage s^1 free
        if (j + 1 < npairs) issue_loads(s ^ 1u, instA + 2, 2 * j + 3 < cnt, !src_shared);
        float tA = 0.f, tB = 0.f;
        if (FUSED) {
            tA = s_t[2 * j]; tB = hasB ? s_t[2 * j + 1] : tA;           // visible since barrier (A) of pass 0
        }

        if (active) {
            u64 V[NC];                                                // vertex components as (A, B) pairs
            if (FUSED) {
                float a[NC], b[NC];
#pragma unroll
                for (int k = 0; k < NC; ++k) { a[k] = fadd(k1[k], fmul(tA, kd[k])); b[k] = fadd(k1[k], fmul(tB, kd[k])); }
#pragma unroll
                for (int k = 0; k < NC; ++k) V[k] = pk(a[k], b[k]);
                if (NORM) {
                    // Vector3::normalise (OgreVector.h:494-508): len = sqrt(x*x + y*y + z*z); scale by 1/len only when len > 0
                    const u64 l2 = padd(padd(pmul(V[3], V[3], NZ), pmul(V[4], V[4], NZ)), pmul(V[5], V[5], NZ));
                    float lA, lB;
                    upk(l2, lA, lB);
                    lA = fsqrt(lA); lB = fsqrt(lB);
                    const float iA = lA > 0.0f ? __frcp_rn(lA) : 1.0f, iB = lB > 0.0f ? __frcp_rn(lB) : 1.0f;   // x * 1.0f == x
                    const u64 I = pk(iA, iB);
                    V[3] = pmul(V[3], I, NZ); V[4] = pmul(V[4], I, NZ); V[5] = pmul(V[5], I, NZ);
                }
                if (P.morph_out) {
                    float* mo = reinterpret_cast<float*>(P.morph_out + (size_t)instA * P.morph_inst_stride + tile_off + (size_t)tid * VS);
                    float* mb = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(mo) + P.morph_inst_stride);
#pragma unroll
                    for (int k = 0; k < NC; ++k) { float x, y; upk(V[k], x, y); mo[k] = x; if (hasB) mb[k] = y; }
                }
            } else {
                const float* sa = reinterpret_cast<const float*>(s_src + (ss * 2) * TILE_BYTES + tid * VS);
                const float* sb = (src_shared || !hasB) ? sa : reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(sa) + TILE_BYTES);
#pragma unroll
                for (int k = 0; k < NC; ++k) V[k] = pk(sa[k], sb[k]);
            }

            // ---- collapse the weighted bone matrix of both instances: E[4*r + c] = (M_A[r][c], M_B[r][c])
            u64 E[12];
            const uint8_t* pal = s_pal + s * pal_stage_bytes;
            {
                const ulonglong2* m0 = reinterpret_cast<const ulonglong2*>(pal + boff[0]);
                if (NW == 1) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) { const ulonglong2 v = m0[k]; E[2 * k] = v.x; E[2 * k + 1] = v.y; }
                } else if (NW == 2) {
                    const ulonglong2* m1 = reinterpret_cast<const ulonglong2*>(pal + boff[1]);
#pragma unroll
                    for (int k = 0; k < 6; ++k) {                       // lerp(M0, M1, w1) = (M1 - M0) * w1 + M0
                        const ulonglong2 a = m0[k], b = m1[k];
                        E[2 * k] = padd(pmul(psub(b.x, a.x), W[1], NZ), a.x);
                        E[2 * k + 1] = padd(pmul(psub(b.y, a.y), W[1], NZ), a.y);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 6; ++k) { const ulonglong2 v = m0[k]; E[2 * k] = pmul(v.x, W[0], NZ); E[2 * k + 1] = pmul(v.y, W[0], NZ); }
#pragma unroll
                    for (int b = 1; b < NW; ++b) {
                        const ulonglong2* mb = reinterpret_cast<const ulonglong2*>(pal + boff[b]);
#pragma unroll
                        for (int k = 0; k < 6; ++k) {
                            const ulonglong2 v = mb[k];
                            E[2 * k] = padd(pmul(v.x, W[b], NZ), E[2 * k]);
                            E[2 * k + 1] = padd(pmul(v.y, W[b], NZ), E[2 * k + 1]);
                        }
                    }
                }
            }
            // ---- position: (r.x*x + r.y*y) + (r.z*z + r.w)
            u64 O[NC];
#pragma unroll
            for (int r = 0; r < 3; ++r)
                O[r] = padd(padd(pmul(E[4 * r], V[0], NZ), pmul(E[4 * r + 1], V[1], NZ)), padd(pmul(E[4 * r + 2], V[2], NZ), E[4 * r + 3]));
            if (NORM) {
                // ---- normal: (r.x*x + r.y*y) + r.z*z, then exact 1/sqrt(len2)
                u64 N[3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    N[r] = padd(padd(pmul(E[4 * r], V[3], NZ), pmul(E[4 * r + 1], V[4], NZ)), pmul(E[4 * r + 2], V[5], NZ));
                const u64 l2 = padd(padd(pmul(N[0], N[0], NZ), pmul(N[1], N[1], NZ)), pmul(N[2], N[2], NZ));
                float lA, lB;
                upk(l2, lA, lB);
                const u64 R = pk(rsq_exact(lA), rsq_exact(lB));
                O[3] = pmul(N[0], R, NZ); O[4] = pmul(N[1], R, NZ); O[5] = pmul(N[2], R, NZ);
            }
            float* da = reinterpret_cast<float*>(s_dst + (s * 2) * TILE_BYTES + tid * VS);
            float* db = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(da) + TILE_BYTES);
#pragma unroll
            for (int k = 0; k < NC; ++k) { float x, y; upk(O[k], x, y); da[k] = x; db[k] = y; }
        }
        if (tma_st) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my STS visible to the async proxy
        __syncthreads();                                              // (B) dst[s] complete
        uint8_t* gA = P.dst + (size_t)instA * P.dst_inst_stride + tile_off;
        uint8_t* gB = gA + P.dst_inst_stride;
        const uint8_t* sA = s_dst + (s * 2) * TILE_BYTES;
        if (tma_st) {
            if (tid == 0) {
                sp_tma_store(gA, sp_smem_u32(sA), bytes);
                if (hasB) sp_tma_store(gB, sp_smem_u32(sA + TILE_BYTES), bytes);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else if (P.store_unit == 8) {
            const uint32_t nu = bytes >> 3;
#pragma unroll
            for (int q = 0; q < MAXU4 / 2; ++q) {
                const uint32_t u = tid + SP_THREADS * q;
                if (u < nu) {
                    __stcs(reinterpret_cast<uint2*>(gA) + u, reinterpret_cast<const uint2*>(sA)[u]);
                    if (hasB) __stcs(reinterpret_cast<uint2*>(gB) + u, reinterpret_cast<const uint2*>(sA + TILE_BYTES)[u]);
                }
            }
        } else {
            const uint32_t nu = bytes >> 2;
#pragma unroll
            for (int q = 0; q < MAXU4; ++q) {
                const uint32_t u = tid + SP_THREADS * q;
                if (u < nu) {
                    __stcs(reinterpret_cast<uint32_t*>(gA) + u, reinterpret_cast<const uint32_t*>(sA)[u]);
                    if (hasB) __stcs(reinterpret_cast<uint32_t*>(gB) + u, reinterpret_cast<const uint32_t*>(sA + TILE_BYTES)[u]);
                }
            }
        }
    }
    if (tma_st && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must outlive the reads
}

static size_t skin_pair_smem(bool norm, bool fused, uint32_t num_matrices)
{
    const size_t tile = (size_t)SP_THREADS * (norm ? 24 : 12);
    return 4 * tile + (fused ? 0 : 4 * tile) + 2 * (size_t)num_matrices * (SP_MSTRIDE * 4) + 16;   // + morph weights, see the launch
}

template <int NW, bool NORM, bool FUSED>
static int launch_one(ogrecu_ctx* ctx, const SkinTileParams& P, dim3 grid, size_t smem)
{
    static size_t granted[64] = {0};    // per device: the opt-in is sticky, raise it only when a larger palette arrives
    if (smem > 48 * 1024 && smem > granted[ctx->device & 63]) {
        OGRECU_CUDA(cudaFuncSetAttribute(skin_pair_kernel<NW, NORM, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        granted[ctx->device & 63] = smem;
    }
    skin_pair_kernel<NW, NORM, FUSED><<<grid, SP_THREADS, smem, ctx->stream>>>(P, 0x8000000080000000ull);
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

// P arrives with the alignment-derived fields filled by launch_skin_tile; picks instances per CTA (even) and launches.
int launch_skin_pair(ogrecu_ctx* ctx, SkinTileParams P, uint32_t nw, bool norm, bool fused)
{
    const uint32_t tiles = div_up(P.num_vertices, SP_THREADS);
    // instances per CTA: amortise the tile's constants while keeping >= ~4 waves of CTAs; even, so only a CTA's
    // last pass can be half empty
    const uint64_t work = (uint64_t)tiles * P.num_instances;
    uint64_t ipc = work / ((uint64_t)ctx->sm_count * SP_MIN_CTAS * 8);
    if (ipc < 2) ipc = 2;
    if (ipc > 64) ipc = 64;
    if (ctx->skin_ipc > 0) ipc = (uint64_t)ctx->skin_ipc;
    ipc = (ipc + 1) & ~1ull;
    P.inst_per_cta = (uint32_t)ipc;
    const uint32_t gy = div_up(P.num_instances, P.inst_per_cta);
    OGRECU_REQUIRE(gy <= 65535, "too many instance groups");
    const size_t smem = skin_pair_smem(norm, fused, P.num_matrices) + (fused ? (size_t)P.inst_per_cta * sizeof(float) : 0);
    dim3 grid(tiles, gy);
    if (norm) return fused ? launch_nw<true, true>(ctx, P, nw, grid, smem) : launch_nw<true, false>(ctx, P, nw, grid, smem);
    return fused ? launch_nw<false, true>(ctx, P, nw, grid, smem) : launch_nw<false, false>(ctx, P, nw, grid, smem);
}

}  // namespace ogrecu
