// optutil.cu — CUDA back end of Ogre::OptimisedUtil (OgreMain/include/OgreOptimisedUtil.h:99-206).
//
// Six bandwidth-bound streams.  Arithmetic replays the SSE implementation's operation
// order (OgreMain/src/OgreOptimisedUtilSSE.cpp) with explicit __fmul_rn/__fadd_rn so that
// nothing is contracted into FMA; where SSE uses the ~12-bit rsqrt estimate (normals,
// extrusion) or is wrong (morph normals) the exact General formulation
// (OgreMain/src/OgreOptimisedUtilGeneral.cpp) is followed — see DESIGN.md.
#include "common.cuh"

#include <algorithm>
#include <vector>
#include "skin_math.cuh"
#include "skin_tile.cuh"
#include "exact_math.cuh"

#include <cstring>

using namespace ogrecu;

namespace {

template <typename T>
__device__ __forceinline__ const T* byte_off(const T* p, size_t off) { return (const T*)((const uint8_t*)p + off); }
template <typename T>
__device__ __forceinline__ T* byte_off(T* p, size_t off) { return (T*)((uint8_t*)p + off); }

/* ------------------------------------------------------------------ skinning (K6) --- */

constexpr int SKIN_THREADS = 256;
constexpr int SKIN_MAX_MATRICES = 256;

// Collapse + transform one vertex; SSE formulation:
//   collapse  OgreOptimisedUtilSSE.cpp:278-368   transform  OgreSIMDHelper.h:204-227
template <int NW, bool NORM>
__device__ __forceinline__ void skin_vertex(const ogrecu_skin_job& j, const float4* __restrict__ pal, float* dst_pos_base,
                                            float* dst_norm_base, uint32_t v)
{
    const float* sp = byte_off(j.src_pos, (size_t)v * j.src_pos_stride);
    const float* bw = byte_off(j.blend_weight, (size_t)v * j.blend_weight_stride);
    const uint8_t* bi = j.blend_index + (size_t)v * j.blend_index_stride;

    float x = sp[0], y = sp[1], z = sp[2];
    float4 r0, r1, r2;
    {
        const float4* m0 = pal + 3 * (int)bi[0];
        if (NW == 1) {
            r0 = m0[0]; r1 = m0[1]; r2 = m0[2];
        } else if (NW == 2) {
            const float4* m1 = pal + 3 * (int)bi[1];
            float w1 = bw[1];
            r0 = lerp4(m0[0], m1[0], w1); r1 = lerp4(m0[1], m1[1], w1); r2 = lerp4(m0[2], m1[2], w1);
        } else {
            float w0 = bw[0];
            r0 = mul4(m0[0], w0); r1 = mul4(m0[1], w0); r2 = mul4(m0[2], w0);
#pragma unroll
            for (int b = 1; b < NW; ++b) {
                const float4* mb = pal + 3 * (int)bi[b];
                float wb = bw[b];
                r0 = madd4(mb[0], wb, r0); r1 = madd4(mb[1], wb, r1); r2 = madd4(mb[2], wb, r2);
            }
        }
    }
    float* dp = byte_off(dst_pos_base, (size_t)v * j.dst_pos_stride);
    dp[0] = fadd(fadd(fmul(r0.x, x), fmul(r0.y, y)), fadd(fmul(r0.z, z), r0.w));
    dp[1] = fadd(fadd(fmul(r1.x, x), fmul(r1.y, y)), fadd(fmul(r1.z, z), r1.w));
    dp[2] = fadd(fadd(fmul(r2.x, x), fmul(r2.y, y)), fadd(fmul(r2.z, z), r2.w));
    if (NORM) {
        const float* sn = byte_off(j.src_norm, (size_t)v * j.src_norm_stride);
        float nx = sn[0], ny = sn[1], nz = sn[2];
        float a = fadd(fadd(fmul(r0.x, nx), fmul(r0.y, ny)), fmul(r0.z, nz));
        float b = fadd(fadd(fmul(r1.x, nx), fmul(r1.y, ny)), fmul(r1.z, nz));
        float c = fadd(fadd(fmul(r2.x, nx), fmul(r2.y, ny)), fmul(r2.z, nz));
        float len2 = fadd(fadd(fmul(a, a), fmul(b, b)), fmul(c, c));
        float rs = fdiv(1.0f, fsqrt(len2));   // exact where SSE takes _mm_rsqrt_ps (:590,:695)
        float* dn = byte_off(dst_norm_base, (size_t)v * j.dst_norm_stride);
        dn[0] = fmul(a, rs); dn[1] = fmul(b, rs); dn[2] = fmul(c, rs);
    }
}

__device__ __forceinline__ void load_palette(float4* pal, const float* palette, const uint16_t* bone_map, uint32_t n)
{
    // rows 0..2 of each Affine3 (row 3 is 0,0,0,1 and never used by skinning)
    for (uint32_t i = threadIdx.x; i < n * 3; i += blockDim.x) {
        uint32_t m = i / 3, r = i - m * 3;
        uint32_t src = bone_map ? (uint32_t)bone_map[m] : m;
        pal[i] = __ldg(reinterpret_cast<const float4*>(palette + (size_t)src * 16 + r * 4));
    }
}

__device__ __forceinline__ void skin_block(const ogrecu_skin_job& j, const float4* pal, float* dst_pos, float* dst_norm,
                                           uint32_t v)
{
    if (v >= j.num_vertices) return;
    const bool norm = j.src_norm != nullptr;
    switch (j.num_weights) {
    case 1: norm ? skin_vertex<1, true>(j, pal, dst_pos, dst_norm, v) : skin_vertex<1, false>(j, pal, dst_pos, dst_norm, v); break;
    case 2: norm ? skin_vertex<2, true>(j, pal, dst_pos, dst_norm, v) : skin_vertex<2, false>(j, pal, dst_pos, dst_norm, v); break;
    case 3: norm ? skin_vertex<3, true>(j, pal, dst_pos, dst_norm, v) : skin_vertex<3, false>(j, pal, dst_pos, dst_norm, v); break;
    default: norm ? skin_vertex<4, true>(j, pal, dst_pos, dst_norm, v) : skin_vertex<4, false>(j, pal, dst_pos, dst_norm, v); break;
    }
}

// batched: one launch covers a descriptor list; block_prefix[k] = first block of job k
__global__ void __launch_bounds__(SKIN_THREADS) skin_batched_kernel(const ogrecu_skin_job* __restrict__ jobs,
                                                                    const uint32_t* __restrict__ block_prefix, uint32_t njobs)
{
    __shared__ float4 pal[SKIN_MAX_MATRICES * 3];
    __shared__ ogrecu_skin_job sj;
    __shared__ uint32_t s_first;
    if (threadIdx.x == 0) {
        uint32_t lo = 0, hi = njobs;   // last k with block_prefix[k] <= blockIdx.x
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (block_prefix[mid] <= blockIdx.x) lo = mid; else hi = mid;
        }
        sj = jobs[lo];
        s_first = block_prefix[lo];
    }
    __syncthreads();
    load_palette(pal, sj.palette, sj.bone_map, sj.num_matrices);
    __syncthreads();
    uint32_t v = (blockIdx.x - s_first) * SKIN_THREADS + threadIdx.x;
    skin_block(sj, pal, sj.dst_pos, sj.dst_norm, v);
}

// instanced: blockIdx.y = instance; shared source/blend streams, per-instance palette + destination
__global__ void __launch_bounds__(SKIN_THREADS) skin_instanced_kernel(const __grid_constant__ ogrecu_skin_job j0,
                                                                      size_t palette_stride_floats, size_t src_instance_stride,
                                                                      size_t dst_instance_stride)
{
    __shared__ float4 pal[SKIN_MAX_MATRICES * 3];
    const uint32_t inst = blockIdx.y;
    load_palette(pal, j0.palette + (size_t)inst * palette_stride_floats, j0.bone_map, j0.num_matrices);
    __syncthreads();
    uint32_t v = blockIdx.x * SKIN_THREADS + threadIdx.x;
    ogrecu_skin_job j = j0;
    j.src_pos = byte_off(j0.src_pos, (size_t)inst * src_instance_stride);
    if (j0.src_norm) j.src_norm = byte_off(j0.src_norm, (size_t)inst * src_instance_stride);
    float* dp = byte_off(j0.dst_pos, (size_t)inst * dst_instance_stride);
    float* dn = j0.dst_norm ? byte_off(j0.dst_norm, (size_t)inst * dst_instance_stride) : nullptr;
    skin_block(j, pal, dp, dn, v);
}

/* --------------------------------------------------------------------- morph (K7) --- */

// generic strides: one thread per vertex.  dst = a + t*(b-a)  (OgreOptimisedUtilGeneral.cpp:249-256)
template <bool NORMALS>
__global__ void morph_strided_kernel(float t, const float* __restrict__ s1, const float* __restrict__ s2, float* __restrict__ d,
                                     size_t st1, size_t st2, size_t std_, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = byte_off(s1, i * st1);
    const float* b = byte_off(s2, i * st2);
    float* o = byte_off(d, i * std_);
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] = fadd(a[k], fmul(t, fsub(b[k], a[k])));
    if (NORMALS) {
        float n0 = fadd(a[3], fmul(t, fsub(b[3], a[3])));
        float n1 = fadd(a[4], fmul(t, fsub(b[4], a[4])));
        float n2 = fadd(a[5], fmul(t, fsub(b[5], a[5])));
        // Vector3::normalise (OgreVector.h:494-508): len = sqrt(((0+x*x)+y*y)+z*z); if len>0 scale by 1/len
        float len = fsqrt(fadd(fadd(fadd(0.0f, fmul(n0, n0)), fmul(n1, n1)), fmul(n2, n2)));
        if (len > 0.0f) { float inv = fdiv(1.0f, len); n0 = fmul(n0, inv); n1 = fmul(n1, inv); n2 = fmul(n2, inv); }
        o[3] = n0; o[4] = n1; o[5] = n2;
    }
}

// dense positions (all strides 12, 16-byte aligned): pure elementwise float4 stream
__global__ void morph_dense_kernel(float t, const float4* __restrict__ s1, const float4* __restrict__ s2, float4* __restrict__ d, size_t n4)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        float4 a = __ldcs(s1 + i), b = __ldcs(s2 + i);
        float4 o = make_float4(fadd(a.x, fmul(t, fsub(b.x, a.x))), fadd(a.y, fmul(t, fsub(b.y, a.y))),
                               fadd(a.z, fmul(t, fsub(b.z, a.z))), fadd(a.w, fmul(t, fsub(b.w, a.w))));
        __stcs(d + i, o);
    }
}

// float tail (< 4 elements) of the dense stream
__global__ void morph_tail_kernel(float t, const float* __restrict__ s1, const float* __restrict__ s2, float* __restrict__ d, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = fadd(s1[i], fmul(t, fsub(s2[i], s1[i])));
}

// dense pos+normal (all strides 24, 16-byte aligned): two vertices = three float4 per thread
__global__ void morph_dense_normals_kernel(float t, const float4* __restrict__ s1, const float4* __restrict__ s2,
                                           float4* __restrict__ d, size_t npairs)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    float a[12], b[12], o[12];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float4 va = __ldcs(s1 + 3 * i + k), vb = __ldcs(s2 + 3 * i + k);
        a[4 * k] = va.x; a[4 * k + 1] = va.y; a[4 * k + 2] = va.z; a[4 * k + 3] = va.w;
        b[4 * k] = vb.x; b[4 * k + 1] = vb.y; b[4 * k + 2] = vb.z; b[4 * k + 3] = vb.w;
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) o[k] = fadd(a[k], fmul(t, fsub(b[k], a[k])));
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        float n0 = o[6 * v + 3], n1 = o[6 * v + 4], n2 = o[6 * v + 5];
        float len = fsqrt(fadd(fadd(fadd(0.0f, fmul(n0, n0)), fmul(n1, n1)), fmul(n2, n2)));
        if (len > 0.0f) { float inv = fdiv(1.0f, len); o[6 * v + 3] = fmul(n0, inv); o[6 * v + 4] = fmul(n1, inv); o[6 * v + 5] = fmul(n2, inv); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) __stcs(d + 3 * i + k, make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
}

// instanced morph from two shared keyframes; blockIdx.y = instance, t read per instance.
// Dense pos+normal (stride 24, 16-byte aligned): two vertices = three float4 per thread.
template <bool NORMALS, bool DENSE>
__global__ void morph_instanced_kernel(const float* __restrict__ tarr, const float* __restrict__ s1, const float* __restrict__ s2,
                                       float* __restrict__ d0, size_t st1, size_t st2, size_t std_, size_t n, size_t dst_inst_stride)
{
    const float t = __ldg(tarr + blockIdx.y);
    float* d = byte_off(d0, (size_t)blockIdx.y * dst_inst_stride);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (DENSE) {
        // unit = 12 floats (pos-only: 4 vertices; pos+normal: 2 vertices)
        const size_t units = NORMALS ? n / 2 : n / 4;
        if (i < units) {
            const float4* a4 = reinterpret_cast<const float4*>(s1) + 3 * i;
            const float4* b4 = reinterpret_cast<const float4*>(s2) + 3 * i;
            float a[12], b[12], o[12];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float4 va = __ldg(a4 + k), vb = __ldg(b4 + k);
                a[4 * k] = va.x; a[4 * k + 1] = va.y; a[4 * k + 2] = va.z; a[4 * k + 3] = va.w;
                b[4 * k] = vb.x; b[4 * k + 1] = vb.y; b[4 * k + 2] = vb.z; b[4 * k + 3] = vb.w;
            }
#pragma unroll
            for (int k = 0; k < 12; ++k) o[k] = fadd(a[k], fmul(t, fsub(b[k], a[k])));
            if (NORMALS) {
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    float n0 = o[6 * v + 3], n1 = o[6 * v + 4], n2 = o[6 * v + 5];
                    float len = fsqrt(fadd(fadd(fadd(0.0f, fmul(n0, n0)), fmul(n1, n1)), fmul(n2, n2)));
                    if (len > 0.0f) { float inv = fdiv(1.0f, len); o[6 * v + 3] = fmul(n0, inv); o[6 * v + 4] = fmul(n1, inv); o[6 * v + 5] = fmul(n2, inv); }
                }
            }
            float4* o4 = reinterpret_cast<float4*>(d) + 3 * i;
#pragma unroll
            for (int k = 0; k < 3; ++k) __stcs(o4 + k, make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
            return;
        }
        // leftover vertices (n % 2 or n % 4) fall through to the strided form on the last threads
        const size_t done = units * (NORMALS ? 2 : 4);
        i = done + (i - units);
        if (i >= n) return;
    } else if (i >= n) return;
    const float* a = byte_off(s1, i * st1);
    const float* b = byte_off(s2, i * st2);
    float* o = byte_off(d, i * std_);
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] = fadd(a[k], fmul(t, fsub(b[k], a[k])));
    if (NORMALS) {
        float n0 = fadd(a[3], fmul(t, fsub(b[3], a[3])));
        float n1 = fadd(a[4], fmul(t, fsub(b[4], a[4])));
        float n2 = fadd(a[5], fmul(t, fsub(b[5], a[5])));
        float len = fsqrt(fadd(fadd(fadd(0.0f, fmul(n0, n0)), fmul(n1, n1)), fmul(n2, n2)));
        if (len > 0.0f) { float inv = fdiv(1.0f, len); n0 = fmul(n0, inv); n1 = fmul(n1, inv); n2 = fmul(n2, inv); }
        o[3] = n0; o[4] = n1; o[5] = n2;
    }
}

/* -------------------------------------------------------------- concatenate (K8) --- */

struct Mat4 { float m[16]; };

// dst[i] = base * src[i], SSE association (t0+t1)+(t2+t3) with t3 = base_row * base_row3
// (OgreOptimisedUtilSSE.cpp:1545-1580).  One thread per matrix (n <= 256 in OGRE): all three
// source rows are read before anything is written, so dst may alias src like the reference.
__global__ void concat_affine_kernel(const __grid_constant__ Mat4 base, const float4* src, float4* dst, size_t n)
{
    size_t mi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (mi >= n) return;
    float4 s0 = src[mi * 4 + 0], s1 = src[mi * 4 + 1], s2 = src[mi * 4 + 2];
    float4 o[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float b0 = base.m[4 * r + 0], b1 = base.m[4 * r + 1], b2 = base.m[4 * r + 2], b3 = base.m[4 * r + 3];
        o[r].x = fadd(fadd(fmul(b0, s0.x), fmul(b1, s1.x)), fadd(fmul(b2, s2.x), fmul(b0, base.m[12])));
        o[r].y = fadd(fadd(fmul(b0, s0.y), fmul(b1, s1.y)), fadd(fmul(b2, s2.y), fmul(b1, base.m[13])));
        o[r].z = fadd(fadd(fmul(b0, s0.z), fmul(b1, s1.z)), fadd(fmul(b2, s2.z), fmul(b2, base.m[14])));
        o[r].w = fadd(fadd(fmul(b0, s0.w), fmul(b1, s1.w)), fadd(fmul(b2, s2.w), fmul(b3, base.m[15])));
    }
    dst[mi * 4 + 0] = o[0]; dst[mi * 4 + 1] = o[1]; dst[mi * 4 + 2] = o[2];
    dst[mi * 4 + 3] = make_float4(0.f, 0.f, 0.f, 1.f);
}

// The same for many entities at once — Entity::cacheBoneMatrices (OgreMain/src/OgreEntity.cpp:1307-1330) concatenates every
// entity's parent-node transform with its bone matrices into mBoneWorldMatrices: instance i uses bases[i] (16 floats, row
// major) on its n matrices at src + i * stride.  One thread per (instance, matrix); same operation order as above.
__global__ void concat_affine_instanced_kernel(const float4* __restrict__ bases, const float4* src, float4* dst, uint32_t n, uint32_t ninst,
                                               size_t stride_f4)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * ninst) return;
    const uint32_t inst = (uint32_t)(t / n), mi = (uint32_t)(t - (size_t)inst * n);
    const float4* b = bases + (size_t)inst * 4;
    const float4 B[4] = {__ldg(b), __ldg(b + 1), __ldg(b + 2), __ldg(b + 3)};
    const float4* sp = src + (size_t)inst * stride_f4 + (size_t)mi * 4;
    float4* dp = dst + (size_t)inst * stride_f4 + (size_t)mi * 4;
    const float4 s0 = sp[0], s1 = sp[1], s2 = sp[2];
    float4 o[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float b0 = B[r].x, b1 = B[r].y, b2 = B[r].z, b3 = B[r].w;
        o[r].x = fadd(fadd(fmul(b0, s0.x), fmul(b1, s1.x)), fadd(fmul(b2, s2.x), fmul(b0, B[3].x)));
        o[r].y = fadd(fadd(fmul(b0, s0.y), fmul(b1, s1.y)), fadd(fmul(b2, s2.y), fmul(b1, B[3].y)));
        o[r].z = fadd(fadd(fmul(b0, s0.z), fmul(b1, s1.z)), fadd(fmul(b2, s2.z), fmul(b2, B[3].z)));
        o[r].w = fadd(fadd(fmul(b0, s0.w), fmul(b1, s1.w)), fadd(fmul(b2, s2.w), fmul(b3, B[3].w)));
    }
    dp[0] = o[0]; dp[1] = o[1]; dp[2] = o[2];
    dp[3] = make_float4(0.f, 0.f, 0.f, 1.f);
}

/* ------------------------------------------------------------- face normals (K9) --- */

// n = (v1-v0)x(v2-v0), w = -(n.v0); EdgeData::Triangle is 32 bytes with vertIndex[3] at byte 8
// (OgreEdgeListBuilder.h:59-72).  OgreOptimisedUtilSSE.cpp:1659-1676 / General :285-307.
__global__ void face_normals_kernel(const float* __restrict__ pos, const uint32_t* __restrict__ tris, float4* __restrict__ out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* t = tris + i * 8;
    uint32_t i0 = __ldcs(t + 2), i1 = __ldcs(t + 3), i2 = __ldcs(t + 4);
    const float* p0 = pos + (size_t)i0 * 3;
    const float* p1 = pos + (size_t)i1 * 3;
    const float* p2 = pos + (size_t)i2 * 3;
    float x0 = __ldg(p0), y0 = __ldg(p0 + 1), z0 = __ldg(p0 + 2);
    float ax = fsub(__ldg(p1), x0), ay = fsub(__ldg(p1 + 1), y0), az = fsub(__ldg(p1 + 2), z0);
    float bx = fsub(__ldg(p2), x0), by = fsub(__ldg(p2 + 1), y0), bz = fsub(__ldg(p2 + 2), z0);
    float nx = fsub(fmul(ay, bz), fmul(az, by));
    float ny = fsub(fmul(az, bx), fmul(ax, bz));
    float nz = fsub(fmul(ax, by), fmul(ay, bx));
    float d = fadd(fadd(fmul(nx, x0), fmul(ny, y0)), fmul(nz, z0));
    __stcs(out + i, make_float4(nx, ny, nz, -d));
}

/* ------------------------------------------------------------ light facing (K10) --- */

struct Lights { float l[8][4]; };

// facing = !(((x+z)+(y+w)) <= 0)   (OgreOptimisedUtilSSE.cpp:1776-1793).  One thread per 4
// faces per light: 4 float4 loads, one 32-bit store.
__global__ void light_facing_kernel(const __grid_constant__ Lights L, uint32_t nlights, const float4* __restrict__ fn,
                                    char* __restrict__ out, size_t out_stride, size_t n)
{
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // quad index
    size_t i = q * 4;
    if (i >= n) return;
    float4 f[4];
    int cnt = (n - i) >= 4 ? 4 : (int)(n - i);
#pragma unroll
    for (int k = 0; k < 4; ++k) f[k] = (k < cnt) ? __ldcs(fn + i + k) : make_float4(0, 0, 0, 0);
    for (uint32_t li = 0; li < nlights; ++li) {
        float lx = L.l[li][0], ly = L.l[li][1], lz = L.l[li][2], lw = L.l[li][3];
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float dp = fadd(fadd(fmul(f[k].x, lx), fmul(f[k].z, lz)), fadd(fmul(f[k].y, ly), fmul(f[k].w, lw)));
            packed |= (!(dp <= 0.0f) ? 1u : 0u) << (8 * k);
        }
        char* o = out + (size_t)li * out_stride + i;
        if (cnt == 4 && (((uintptr_t)o) & 3) == 0) *reinterpret_cast<uint32_t*>(o) = packed;
        else for (int k = 0; k < cnt; ++k) o[k] = (char)((packed >> (8 * k)) & 0xff);
    }
}

/* ----------------------------------------------------------------- extrude (K11) --- */

// Vector3::normalise then *= dist then src + dir (OgreOptimisedUtilGeneral.cpp:321-365)
__device__ __forceinline__ void extrude_dir(float dx, float dy, float dz, float dist, float& ox, float& oy, float& oz)
{
    float len = fsqrt(fadd(fadd(fadd(0.0f, fmul(dx, dx)), fmul(dy, dy)), fmul(dz, dz)));
    if (len > 0.0f) { float inv = fdiv(1.0f, len); dx = fmul(dx, inv); dy = fmul(dy, inv); dz = fmul(dz, inv); }
    ox = fmul(dx, dist); oy = fmul(dy, dist); oz = fmul(dz, dist);
}

// The four vertices of a thread extruded away from a point light.  sqrt and 1 / len by the branch-free exact sequences of
// exact_math.cuh when all four squared lengths are in their range (every vertex that is neither on the light nor 2^20 units
// from it), the four chains interleaved; __fsqrt_rn / __fdiv_rn otherwise.  Same bits either way: both are the correctly
// rounded sqrt and reciprocal (tools/ubench/exact_rsq.cu) — the two library calls were two thirds of the kernel's instructions.
__device__ __forceinline__ void extrude_point4(const float (&s)[12], int cnt, float lx, float ly, float lz, float dist, float (&o)[12])
{
    float dx[4], dy[4], dz[4], l2[4];
    bool fast = true;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        dx[v] = fsub(s[3 * v], lx); dy[v] = fsub(s[3 * v + 1], ly); dz[v] = fsub(s[3 * v + 2], lz);
        l2[v] = fadd(fadd(fadd(0.0f, fmul(dx[v], dx[v])), fmul(dy[v], dy[v])), fmul(dz[v], dz[v]));
        if (v < cnt) fast = fast && em_fast_range(__float_as_uint(l2[v]));
    }
    if (fast) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            if (v < cnt) {
                const float inv = em_rcp_fast(em_sqrt_fast(l2[v]));            // len > 0 in the fast range
                o[3 * v] = fadd(s[3 * v], fmul(fmul(dx[v], inv), dist));
                o[3 * v + 1] = fadd(s[3 * v + 1], fmul(fmul(dy[v], inv), dist));
                o[3 * v + 2] = fadd(s[3 * v + 2], fmul(fmul(dz[v], inv), dist));
            }
        }
    } else {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            if (v < cnt) {
                float ex, ey, ez;
                extrude_dir(dx[v], dy[v], dz[v], dist, ex, ey, ez);
                o[3 * v] = fadd(s[3 * v], ex); o[3 * v + 1] = fadd(s[3 * v + 1], ey); o[3 * v + 2] = fadd(s[3 * v + 2], ez);
            }
        }
    }
}

// packed xyz; each thread handles 4 vertices = 12 floats (three float4 when aligned)
template <bool POINT, bool ALIGNED>
__global__ void extrude_kernel(float lx, float ly, float lz, float dist, const float* __restrict__ src, float* __restrict__ dst, size_t n)
{
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t v0 = g * 4;
    if (v0 >= n) return;
    int cnt = (n - v0) >= 4 ? 4 : (int)(n - v0);
    float s[12], o[12];
    if (ALIGNED && cnt == 4) {
        const float4* p = reinterpret_cast<const float4*>(src + v0 * 3);
#pragma unroll
        for (int k = 0; k < 3; ++k) { float4 v = __ldcs(p + k); s[4 * k] = v.x; s[4 * k + 1] = v.y; s[4 * k + 2] = v.z; s[4 * k + 3] = v.w; }
    } else {
        for (int k = 0; k < cnt * 3; ++k) s[k] = src[v0 * 3 + k];
    }
    if (POINT) extrude_point4(s, cnt, lx, ly, lz, dist, o);
    else {
        float ex, ey, ez;
        extrude_dir(-lx, -ly, -lz, dist, ex, ey, ez);
#pragma unroll
        for (int v = 0; v < 4; ++v)
            if (v < cnt) { o[3 * v] = fadd(s[3 * v], ex); o[3 * v + 1] = fadd(s[3 * v + 1], ey); o[3 * v + 2] = fadd(s[3 * v + 2], ez); }
    }
    if (ALIGNED && cnt == 4) {
        float4* q = reinterpret_cast<float4*>(dst + v0 * 3);
#pragma unroll
        for (int k = 0; k < 3; ++k) __stcs(q + k, make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
    } else {
        for (int k = 0; k < cnt * 3; ++k) dst[v0 * 3 + k] = o[k];
    }
}

// 1..8 lights per launch: the positions are read once and extruded away from (or along) every light; light k
// writes dst + k*dst_light_stride floats.  Per light the arithmetic is extrude_kernel's.  A shadow-casting frame
// extrudes the same vertices for each light (ShadowCaster::extrudeVertices, OgreShadowCaster.cpp:535), so 8 lights move
// 12 + 8*12 B per vertex in one launch instead of 8 x 24 B in eight 24 MB launches whose latency dominates.
__global__ void extrude_multi_kernel(const __grid_constant__ Lights L, uint32_t nlights, float dist, const float* __restrict__ src,
                                     float* __restrict__ dst, size_t dst_light_stride, size_t n, int aligned)
{
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t v0 = g * 4;
    if (v0 >= n) return;
    int cnt = (n - v0) >= 4 ? 4 : (int)(n - v0);
    float s[12], o[12];
    const bool vec = aligned && cnt == 4;
    if (vec) {
        const float4* p = reinterpret_cast<const float4*>(src + v0 * 3);
#pragma unroll
        for (int k = 0; k < 3; ++k) { float4 v = __ldg(p + k); s[4 * k] = v.x; s[4 * k + 1] = v.y; s[4 * k + 2] = v.z; s[4 * k + 3] = v.w; }
    } else {
        for (int k = 0; k < cnt * 3; ++k) s[k] = src[v0 * 3 + k];
    }
    {
        // blockIdx.y = light: the eight passes over the positions run concurrently and re-read them from L2 (12 MB per
        // million vertices), so DRAM still sees them once; one thread per light keeps 8x the threads in flight than a
        // per-thread loop over the lights (measured: 34.8 us for 8 lights x 1 M vertices with the loop)
        const uint32_t li = blockIdx.y;
        const float lx = L.l[li][0], ly = L.l[li][1], lz = L.l[li][2];
        const bool point = !(L.l[li][3] == 0.0f);
        if (point) extrude_point4(s, cnt, lx, ly, lz, dist, o);
        else {
            float ex, ey, ez;
            extrude_dir(-lx, -ly, -lz, dist, ex, ey, ez);
#pragma unroll
            for (int v = 0; v < 4; ++v)
                if (v < cnt) { o[3 * v] = fadd(s[3 * v], ex); o[3 * v + 1] = fadd(s[3 * v + 1], ey); o[3 * v + 2] = fadd(s[3 * v + 2], ez); }
        }
        float* d = dst + li * dst_light_stride + v0 * 3;
        if (vec) {
            float4* q = reinterpret_cast<float4*>(d);
#pragma unroll
            for (int k = 0; k < 3; ++k) __stcs(q + k, make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
        } else {
            for (int k = 0; k < cnt * 3; ++k) d[k] = o[k];
        }
    }
}

/* ------------------------------------------------------------- pose blending ------- */

// Mesh::softwareVertexPoseBlend (OgreMesh.cpp:1948-2007) for one element (positions or normals) of many instances:
// thread (entry e, instance i): dst_i[idx[e]] = dst_i[idx[e]] + (offset[e] * weight_i), nothing when weight_i == 0 (the
// reference returns before touching the buffer).  A pose names a vertex at most once (std::map keys), so the writes of
// one launch never collide; successive poses are successive launches, as the reference applies them one after the other.
__global__ void pose_blend_kernel(const float* __restrict__ weights, float weight0, const uint32_t* __restrict__ idx,
                                  const float* __restrict__ off3, uint32_t n, uint8_t* __restrict__ dst, uint32_t stride_bytes,
                                  size_t inst_stride_bytes)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const float w = weights ? __ldg(weights + blockIdx.y) : weight0;
    if (w == 0.0f) return;
    float* p = reinterpret_cast<float*>(dst + (size_t)blockIdx.y * inst_stride_bytes + (size_t)__ldg(idx + e) * stride_bytes);
    const float ox = __ldg(off3 + 3 * (size_t)e), oy = __ldg(off3 + 3 * (size_t)e + 1), oz = __ldg(off3 + 3 * (size_t)e + 2);
    p[0] = fadd(p[0], fmul(ox, w));
    p[1] = fadd(p[1], fmul(oy, w));
    p[2] = fadd(p[2], fmul(oz, w));
}

// Entity::finalisePoseNormals (OgreEntity.cpp:1207-1261) for many instances: where the accumulated pose normal is shorter
// than 1 (- 1e-4) the base mesh normal fills in with weight 1 - length, then Vector3::normalise.  `zero_first` instead
// performs initialisePoseVertexData's reset of the normals (:1180-1204) so that a frame needs no separate memset.
__global__ void pose_normals_kernel(uint8_t* __restrict__ dst_norm, const uint8_t* __restrict__ src_norm, uint32_t n, uint32_t dst_stride,
                                    uint32_t src_stride, size_t dst_inst_stride, int zero_first)
{
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    float* d = reinterpret_cast<float*>(dst_norm + (size_t)blockIdx.y * dst_inst_stride + (size_t)v * dst_stride);
    if (zero_first) { d[0] = 0.0f; d[1] = 0.0f; d[2] = 0.0f; return; }
    const float* s = reinterpret_cast<const float*>(src_norm + (size_t)v * src_stride);
    float x = d[0], y = d[1], z = d[2];
    const float len = fsqrt(fadd(fadd(fadd(0.0f, fmul(x, x)), fmul(y, y)), fmul(z, z)));
    if (fadd(len, 1e-4f) < 1.0f) {
        const float bw = fsub(1.0f, len);
        x = fadd(x, fmul(__ldg(s), bw)); y = fadd(y, fmul(__ldg(s + 1), bw)); z = fadd(z, fmul(__ldg(s + 2), bw));
    }
    const float l2 = fsqrt(fadd(fadd(fadd(0.0f, fmul(x, x)), fmul(y, y)), fmul(z, z)));
    if (l2 > 0.0f) { const float inv = fdiv(1.0f, l2); x = fmul(x, inv); y = fmul(y, inv); z = fmul(z, inv); }
    d[0] = x; d[1] = y; d[2] = z;
}

inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

int validate_skin_job(const ogrecu_skin_job& j)
{
    OGRECU_REQUIRE(j.num_weights >= 1 && j.num_weights <= 4, "num_weights must be 1..4");
    OGRECU_REQUIRE(j.num_matrices >= 1 && j.num_matrices <= SKIN_MAX_MATRICES, "num_matrices must be 1..256");
    if (j.num_vertices == 0) return OGRECU_OK;
    OGRECU_REQUIRE(j.src_pos && j.dst_pos && j.blend_weight && j.blend_index && j.palette, "NULL stream pointer");
    OGRECU_REQUIRE(!j.src_norm || j.dst_norm, "src_norm without dst_norm");
    OGRECU_REQUIRE(aligned16(j.palette), "palette must be 16-byte aligned");
    OGRECU_REQUIRE((((uintptr_t)j.src_pos | (uintptr_t)j.dst_pos | (uintptr_t)j.blend_weight | j.src_pos_stride |
                     j.dst_pos_stride | j.blend_weight_stride) & 3) == 0, "float streams must be 4-byte aligned");
    OGRECU_REQUIRE(!j.src_norm || ((((uintptr_t)j.src_norm | (uintptr_t)j.dst_norm | j.src_norm_stride | j.dst_norm_stride) & 3) == 0),
                   "normal streams must be 4-byte aligned");
    return OGRECU_OK;
}

}  // namespace

namespace ogrecu {
int launch_skin_batched(ogrecu_ctx* ctx, const ogrecu_skin_job* djobs, const uint32_t* dprefix, uint32_t n, unsigned blocks)
{
    skin_batched_kernel<<<blocks, SKIN_THREADS, 0, ctx->stream>>>(djobs, dprefix, n);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}
int validate_skin_job_public(const ogrecu_skin_job& j) { return validate_skin_job(j); }
}  // namespace ogrecu

extern "C" {

int ogrecu_skin_dev(ogrecu_ctx* ctx, const ogrecu_skin_job* jobs, uint32_t n)
{
    OGRECU_REQUIRE(ctx && (n == 0 || jobs), "NULL argument");
    if (n == 0) return OGRECU_OK;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    // descriptor list + block prefix through pinned staging -> device scratch
    size_t jobs_bytes = (size_t)n * sizeof(ogrecu_skin_job);
    size_t prefix_bytes = ((size_t)n + 1) * sizeof(uint32_t);
    size_t total = jobs_bytes + prefix_bytes;
    OGRECU_CHECK(pinned_reserve(ctx, total));
    OGRECU_CHECK(scratch_reserve(ctx, ctx->jobs, total));
    // the pinned block may still be the source of an earlier async copy
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    uint8_t* h = (uint8_t*)ctx->pinned;
    uint32_t* prefix = (uint32_t*)(h + jobs_bytes);
    uint64_t blocks = 0;
    for (uint32_t k = 0; k < n; ++k) {
        OGRECU_CHECK(validate_skin_job(jobs[k]));
        prefix[k] = (uint32_t)blocks;
        blocks += div_up(jobs[k].num_vertices, SKIN_THREADS);
        OGRECU_REQUIRE(blocks < 0x7fffffffull, "too many vertices in one batch");
    }
    prefix[n] = (uint32_t)blocks;
    memcpy(h, jobs, jobs_bytes);
    if (blocks == 0) return OGRECU_OK;
    OGRECU_CUDA(cudaMemcpyAsync(ctx->jobs.ptr, h, total, cudaMemcpyHostToDevice, ctx->stream));
    const ogrecu_skin_job* djobs = (const ogrecu_skin_job*)ctx->jobs.ptr;
    const uint32_t* dprefix = (const uint32_t*)((uint8_t*)ctx->jobs.ptr + jobs_bytes);
    skin_batched_kernel<<<(unsigned)blocks, SKIN_THREADS, 0, ctx->stream>>>(djobs, dprefix, n);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

int ogrecu_skin_instanced_dev(ogrecu_ctx* ctx, const ogrecu_skin_job* job, uint32_t num_instances,
                              size_t palette_stride_floats, size_t src_instance_stride_bytes, size_t dst_instance_stride_bytes)
{
    OGRECU_REQUIRE((src_instance_stride_bytes & 3) == 0, "source instance stride must be a multiple of 4");
    OGRECU_REQUIRE(ctx && job, "NULL argument");
    OGRECU_CHECK(validate_skin_job(*job));
    OGRECU_REQUIRE((palette_stride_floats & 3) == 0, "palette stride must keep 16-byte alignment");
    OGRECU_REQUIRE((dst_instance_stride_bytes & 3) == 0, "destination instance stride must be a multiple of 4");
    OGRECU_REQUIRE(num_instances <= 0x7fffffffu, "too many instances");
    if (num_instances == 0 || job->num_vertices == 0) return OGRECU_OK;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    const ogrecu_skin_job& j = *job;
    // dense layouts take the tiled stream kernel (skin_tile.cu): pos+normal interleaved at stride 24
    // in source and destination (the reference's "shared packed" case, OgreOptimisedUtilSSE.cpp:1100-1130)
    // or positions only at stride 12
    const bool packed24 = j.src_norm && (const uint8_t*)j.src_norm == (const uint8_t*)j.src_pos + 12 &&
                          (uint8_t*)j.dst_norm == (uint8_t*)j.dst_pos + 12 && j.src_pos_stride == 24 && j.src_norm_stride == 24 &&
                          j.dst_pos_stride == 24 && j.dst_norm_stride == 24;
    const bool packed12 = !j.src_norm && j.src_pos_stride == 12 && j.dst_pos_stride == 12;
    if ((packed24 || packed12) && ctx->skin_mode != 2) {
        SkinTileParams P;
        memset(&P, 0, sizeof P);
        P.src = (const uint8_t*)j.src_pos; P.dst = (uint8_t*)j.dst_pos;
        P.blend_weight = j.blend_weight; P.blend_index = j.blend_index;
        P.palette = j.palette; P.bone_map = j.bone_map;
        P.palette_stride_floats = palette_stride_floats;
        P.src_inst_stride = src_instance_stride_bytes; P.dst_inst_stride = dst_instance_stride_bytes;
        P.blend_weight_stride = j.blend_weight_stride; P.blend_index_stride = j.blend_index_stride;
        P.num_matrices = j.num_matrices; P.num_vertices = j.num_vertices; P.num_instances = num_instances;
        return launch_skin_tile(ctx, P, j.num_weights, packed24, false);
    }
    OGRECU_REQUIRE(num_instances <= 65535, "at most 65535 instances per call for strided layouts");
    dim3 grid(div_up(job->num_vertices, SKIN_THREADS), num_instances);
    skin_instanced_kernel<<<grid, SKIN_THREADS, 0, ctx->stream>>>(*job, palette_stride_floats, src_instance_stride_bytes, dst_instance_stride_bytes);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

int ogrecu_anim_instanced_dev(ogrecu_ctx* ctx, const ogrecu_skin_job* job, const float* key2, const float* t_dev,
                              float* morph_out, size_t morph_instance_stride_bytes, uint32_t num_instances,
                              size_t palette_stride_floats, size_t dst_instance_stride_bytes)
{
    OGRECU_REQUIRE(ctx && job && key2 && t_dev, "NULL argument");
    OGRECU_CHECK(validate_skin_job(*job));
    const ogrecu_skin_job& j = *job;
    OGRECU_REQUIRE((palette_stride_floats & 3) == 0, "palette stride must keep 16-byte alignment");
    OGRECU_REQUIRE(((dst_instance_stride_bytes | morph_instance_stride_bytes | (uintptr_t)key2 | (uintptr_t)morph_out) & 3) == 0,
                   "streams must be 4-byte aligned");
    const bool packed24 = j.src_norm && (const uint8_t*)j.src_norm == (const uint8_t*)j.src_pos + 12 &&
                          (uint8_t*)j.dst_norm == (uint8_t*)j.dst_pos + 12 && j.src_pos_stride == 24 && j.src_norm_stride == 24 &&
                          j.dst_pos_stride == 24 && j.dst_norm_stride == 24;
    const bool packed12 = !j.src_norm && j.src_pos_stride == 12 && j.dst_pos_stride == 12;
    OGRECU_REQUIRE(packed24 || packed12, "fused morph+skin needs dense keyframes and destination (stride 24 pos+normal, or 12 pos only)");
    OGRECU_REQUIRE(num_instances <= 0x7fffffffu, "too many instances");
    if (num_instances == 0 || j.num_vertices == 0) return OGRECU_OK;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    SkinTileParams P;
    memset(&P, 0, sizeof P);
    P.src = (const uint8_t*)j.src_pos; P.key2 = (const uint8_t*)key2; P.tarr = t_dev;
    P.morph_out = (uint8_t*)morph_out; P.morph_inst_stride = morph_instance_stride_bytes;
    P.dst = (uint8_t*)j.dst_pos;
    P.blend_weight = j.blend_weight; P.blend_index = j.blend_index;
    P.palette = j.palette; P.bone_map = j.bone_map;
    P.palette_stride_floats = palette_stride_floats;
    P.dst_inst_stride = dst_instance_stride_bytes;
    P.blend_weight_stride = j.blend_weight_stride; P.blend_index_stride = j.blend_index_stride;
    P.num_matrices = j.num_matrices; P.num_vertices = j.num_vertices; P.num_instances = num_instances;
    return launch_skin_tile(ctx, P, j.num_weights, packed24, true);
}

// One frame of an instanced mesh = the fused morph+skin of each of its sub-meshes (Entity::updateAnimation walks the
// sub-entities, OgreEntity.cpp:823-864).  The launches are independent, so they alternate over two side streams, largest
// first: the tail of one launch (its last, partly filled wave of CTAs — a CTA lives for ~25 us here) overlaps the head of
// the next instead of idling SMs.  Forked from and joined back into the context's stream with events, so the call is
// ordered like a single launch for everything before and after it.  Measured on the bench step: 0.594 -> 0.558 ms.
int ogrecu_anim_instanced_multi_dev(ogrecu_ctx* ctx, const ogrecu_anim_job* jobs, uint32_t num_jobs, const float* t_dev, uint32_t num_instances,
                                    size_t palette_stride_floats)
{
    OGRECU_REQUIRE(ctx && (num_jobs == 0 || jobs), "NULL argument");
    if (num_jobs == 0) return OGRECU_OK;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    if (num_jobs == 1)
        return ogrecu_anim_instanced_dev(ctx, &jobs[0].skin, jobs[0].key2, t_dev, jobs[0].morph_out, jobs[0].morph_instance_stride_bytes, num_instances,
                                         palette_stride_floats, jobs[0].dst_instance_stride_bytes);
    if (!ctx->side_fork) {
        OGRECU_CUDA(cudaEventCreateWithFlags(&ctx->side_fork, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) {
            OGRECU_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream[k], cudaStreamNonBlocking));
            OGRECU_CUDA(cudaEventCreateWithFlags(&ctx->side_join[k], cudaEventDisableTiming));
        }
    }
    std::vector<uint32_t> order(num_jobs);
    for (uint32_t i = 0; i < num_jobs; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return (uint64_t)jobs[a].skin.num_vertices * jobs[a].skin.num_weights > (uint64_t)jobs[b].skin.num_vertices * jobs[b].skin.num_weights;
    });
    cudaStream_t main_stream = ctx->stream;
    OGRECU_CUDA(cudaEventRecord(ctx->side_fork, main_stream));
    for (int k = 0; k < 2; ++k) OGRECU_CUDA(cudaStreamWaitEvent(ctx->side_stream[k], ctx->side_fork, 0));
    int rc = OGRECU_OK;
    for (uint32_t k = 0; k < num_jobs && rc == OGRECU_OK; ++k) {
        const ogrecu_anim_job& J = jobs[order[k]];
        ctx->stream = ctx->side_stream[k & 1];                  // the launchers use the context's current stream
        rc = ogrecu_anim_instanced_dev(ctx, &J.skin, J.key2, t_dev, J.morph_out, J.morph_instance_stride_bytes, num_instances, palette_stride_floats,
                                       J.dst_instance_stride_bytes);
    }
    ctx->stream = main_stream;
    for (int k = 0; k < 2; ++k) {                               // join even after a failed launch: nothing may be left running unordered
        cudaEventRecord(ctx->side_join[k], ctx->side_stream[k]);
        cudaStreamWaitEvent(main_stream, ctx->side_join[k], 0);
    }
    return rc;
}

int ogrecu_morph_dev(ogrecu_ctx* ctx, float t, const float* src1, const float* src2, float* dst, size_t stride1,
                     size_t stride2, size_t dst_stride, size_t num_vertices, int morph_normals)
{
    OGRECU_REQUIRE(ctx, "ctx is NULL");
    if (num_vertices == 0) return OGRECU_OK;
    OGRECU_REQUIRE(src1 && src2 && dst, "NULL stream pointer");
    const size_t need = morph_normals ? 24 : 12;
    OGRECU_REQUIRE(stride1 >= need && stride2 >= need && dst_stride >= need, "stride smaller than the vertex");
    OGRECU_REQUIRE(((stride1 | stride2 | dst_stride | (uintptr_t)src1 | (uintptr_t)src2 | (uintptr_t)dst) & 3) == 0,
                   "streams must be 4-byte aligned");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    const bool dense = stride1 == need && stride2 == need && dst_stride == need && aligned16(src1) && aligned16(src2) && aligned16(dst);
    if (dense && !morph_normals) {
        size_t nfloat = num_vertices * 3, n4 = nfloat / 4;
        if (n4) {
            unsigned blocks = (unsigned)std::min<size_t>(div_up(n4, 256), (size_t)ctx->sm_count * 16);
            morph_dense_kernel<<<blocks, 256, 0, ctx->stream>>>(t, (const float4*)src1, (const float4*)src2, (float4*)dst, n4);
            OGRECU_CUDA(cudaGetLastError());
            count_launch(ctx);
        }
        size_t tail0 = n4 * 4;
        if (tail0 < nfloat) {
            morph_tail_kernel<<<1, 32, 0, ctx->stream>>>(t, src1 + tail0, src2 + tail0, dst + tail0, nfloat - tail0);
            OGRECU_CUDA(cudaGetLastError());
            count_launch(ctx);
        }
        return OGRECU_OK;
    }
    if (dense && morph_normals) {
        size_t npairs = num_vertices / 2;
        if (npairs) {
            morph_dense_normals_kernel<<<div_up(npairs, 256), 256, 0, ctx->stream>>>(t, (const float4*)src1, (const float4*)src2, (float4*)dst, npairs);
            OGRECU_CUDA(cudaGetLastError());
            count_launch(ctx);
        }
        if (num_vertices & 1) {
            size_t v0 = num_vertices - 1;
            morph_strided_kernel<true><<<1, 32, 0, ctx->stream>>>(t, src1 + v0 * 6, src2 + v0 * 6, dst + v0 * 6, 24, 24, 24, 1);
            OGRECU_CUDA(cudaGetLastError());
            count_launch(ctx);
        }
        return OGRECU_OK;
    }
    unsigned blocks = div_up(num_vertices, 256);
    if (morph_normals) morph_strided_kernel<true><<<blocks, 256, 0, ctx->stream>>>(t, src1, src2, dst, stride1, stride2, dst_stride, num_vertices);
    else morph_strided_kernel<false><<<blocks, 256, 0, ctx->stream>>>(t, src1, src2, dst, stride1, stride2, dst_stride, num_vertices);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

int ogrecu_morph_instanced_dev(ogrecu_ctx* ctx, const float* t_dev, const float* src1, const float* src2, float* dst,
                               size_t stride1, size_t stride2, size_t dst_stride, size_t n, int morph_normals,
                               uint32_t num_instances, size_t dst_instance_stride_bytes)
{
    OGRECU_REQUIRE(ctx, "ctx is NULL");
    if (n == 0 || num_instances == 0) return OGRECU_OK;
    OGRECU_REQUIRE(t_dev && src1 && src2 && dst, "NULL pointer");
    OGRECU_REQUIRE(num_instances <= 65535, "at most 65535 instances per call");
    const size_t need = morph_normals ? 24 : 12;
    OGRECU_REQUIRE(stride1 >= need && stride2 >= need && dst_stride >= need, "stride smaller than the vertex");
    OGRECU_REQUIRE(((stride1 | stride2 | dst_stride | dst_instance_stride_bytes | (uintptr_t)src1 | (uintptr_t)src2 | (uintptr_t)dst) & 3) == 0,
                   "streams must be 4-byte aligned");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    const bool dense = stride1 == need && stride2 == need && dst_stride == need && aligned16(src1) && aligned16(src2) &&
                       aligned16(dst) && (dst_instance_stride_bytes & 15) == 0;
    size_t threads_needed = n;
    if (dense) { size_t per = morph_normals ? 2 : 4; threads_needed = n / per + n % per; }
    dim3 grid(div_up(threads_needed, 256), num_instances);
    if (dense) {
        if (morph_normals) morph_instanced_kernel<true, true><<<grid, 256, 0, ctx->stream>>>(t_dev, src1, src2, dst, stride1, stride2, dst_stride, n, dst_instance_stride_bytes);
        else morph_instanced_kernel<false, true><<<grid, 256, 0, ctx->stream>>>(t_dev, src1, src2, dst, stride1, stride2, dst_stride, n, dst_instance_stride_bytes);
    } else {
        if (morph_normals) morph_instanced_kernel<true, false><<<grid, 256, 0, ctx->stream>>>(t_dev, src1, src2, dst, stride1, stride2, dst_stride, n, dst_instance_stride_bytes);
        else morph_instanced_kernel<false, false><<<grid, 256, 0, ctx->stream>>>(t_dev, src1, src2, dst, stride1, stride2, dst_stride, n, dst_instance_stride_bytes);
    }
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

int ogrecu_concat_affine_dev(ogrecu_ctx* ctx, const float* base_host16, const float* src, float* dst, size_t n)
{
    OGRECU_REQUIRE(ctx && base_host16, "NULL argument");
    if (n == 0) return OGRECU_OK;
    OGRECU_REQUIRE(src && dst && aligned16(src) && aligned16(dst), "matrices must be non-NULL and 16-byte aligned");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    Mat4 base;
    memcpy(base.m, base_host16, sizeof base.m);
    concat_affine_kernel<<<div_up(n, 64), 64, 0, ctx->stream>>>(base, (const float4*)src, (float4*)dst, n);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

int ogrecu_concat_affine_instanced_dev(ogrecu_ctx* ctx, const float* bases_dev, const float* src, float* dst, uint32_t n, uint32_t num_instances,
                                       size_t instance_stride_floats)
{
    OGRECU_REQUIRE(ctx, "ctx is NULL");
    if (n == 0 || num_instances == 0) return OGRECU_OK;
    OGRECU_REQUIRE(bases_dev && src && dst && aligned16(bases_dev) && aligned16(src) && aligned16(dst), "matrices must be non-NULL and 16-byte aligned");
    OGRECU_REQUIRE((instance_stride_floats & 3) == 0 && instance_stride_floats >= (size_t)n * 16, "instance stride must hold n matrices and keep 16-byte alignment");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    concat_affine_instanced_kernel<<<div_up((size_t)n * num_instances, 128), 128, 0, ctx->stream>>>((const float4*)bases_dev, (const float4*)src,
                                                                                                (float4*)dst, n, num_instances, instance_stride_floats / 4);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

int ogrecu_face_normals_dev(ogrecu_ctx* ctx, const float* positions, const void* triangles, float* face_normals, size_t n)
{
    OGRECU_REQUIRE(ctx, "ctx is NULL");
    if (n == 0) return OGRECU_OK;
    OGRECU_REQUIRE(positions && triangles && face_normals, "NULL pointer");
    OGRECU_REQUIRE(aligned16(face_normals), "face normals must be 16-byte aligned");   // asserted by the SSE path too (:1591)
    OGRECU_REQUIRE((((uintptr_t)triangles | (uintptr_t)positions) & 3) == 0, "inputs must be 4-byte aligned");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    face_normals_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(positions, (const uint32_t*)triangles, (float4*)face_normals, n);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

int ogrecu_light_facing_multi_dev(ogrecu_ctx* ctx, const float* lights_host, uint32_t n_lights, const float* face_normals,
                                  char* facings, size_t facing_stride, size_t n)
{
    OGRECU_REQUIRE(ctx && lights_host, "NULL argument");
    OGRECU_REQUIRE(n_lights >= 1 && n_lights <= 8, "1..8 lights per launch");
    if (n == 0) return OGRECU_OK;
    OGRECU_REQUIRE(face_normals && facings && aligned16(face_normals), "face normals must be non-NULL and 16-byte aligned");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    Lights L;
    memset(&L, 0, sizeof L);
    memcpy(L.l, lights_host, (size_t)n_lights * 16);
    light_facing_kernel<<<div_up(div_up(n, 4), 256), 256, 0, ctx->stream>>>(L, n_lights, (const float4*)face_normals, facings, facing_stride, n);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

int ogrecu_light_facing_dev(ogrecu_ctx* ctx, const float* light4_host, const float* face_normals, char* facings, size_t n)
{
    return ogrecu_light_facing_multi_dev(ctx, light4_host, 1, face_normals, facings, 0, n);
}

int ogrecu_pose_blend_instanced_dev(ogrecu_ctx* ctx, const float* weights_dev, const uint32_t* vertex_index_dev, const float* offsets_dev,
                                    uint32_t num_entries, float* dst_dev, uint32_t dst_stride_bytes, uint32_t num_instances,
                                    size_t dst_instance_stride_bytes)
{
    OGRECU_REQUIRE(ctx && weights_dev, "NULL argument");
    if (num_entries == 0 || num_instances == 0) return OGRECU_OK;
    OGRECU_REQUIRE(vertex_index_dev && offsets_dev && dst_dev, "NULL stream pointer");
    OGRECU_REQUIRE((((uintptr_t)dst_dev | dst_stride_bytes | dst_instance_stride_bytes) & 3) == 0, "destination must be 4-byte aligned");
    OGRECU_REQUIRE(num_instances <= 65535, "at most 65535 instances per call");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    pose_blend_kernel<<<dim3(div_up(num_entries, 256), num_instances), 256, 0, ctx->stream>>>(weights_dev, 0.f, vertex_index_dev, offsets_dev, num_entries,
                                                                                             (uint8_t*)dst_dev, dst_stride_bytes, dst_instance_stride_bytes);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

static int pose_normals_launch(ogrecu_ctx* ctx, float* dst, const float* src, uint32_t n, uint32_t dst_stride, uint32_t src_stride, uint32_t ninst,
                               size_t inst_stride, int zero_first)
{
    OGRECU_REQUIRE(ctx, "ctx is NULL");
    if (n == 0 || ninst == 0) return OGRECU_OK;
    OGRECU_REQUIRE(dst && (zero_first || src), "NULL stream pointer");
    OGRECU_REQUIRE((((uintptr_t)dst | (uintptr_t)src | dst_stride | src_stride | inst_stride) & 3) == 0, "streams must be 4-byte aligned");
    OGRECU_REQUIRE(ninst <= 65535, "at most 65535 instances per call");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    pose_normals_kernel<<<dim3(div_up(n, 256), ninst), 256, 0, ctx->stream>>>((uint8_t*)dst, (const uint8_t*)src, n, dst_stride, src_stride, inst_stride,
                                                                             zero_first);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

int ogrecu_pose_normals_reset_dev(ogrecu_ctx* ctx, float* dst_norm_dev, uint32_t num_vertices, uint32_t dst_stride_bytes, uint32_t num_instances,
                                  size_t dst_instance_stride_bytes)
{
    return pose_normals_launch(ctx, dst_norm_dev, nullptr, num_vertices, dst_stride_bytes, 0, num_instances, dst_instance_stride_bytes, 1);
}

int ogrecu_pose_normals_finalise_dev(ogrecu_ctx* ctx, float* dst_norm_dev, const float* base_norm_dev, uint32_t num_vertices,
                                     uint32_t dst_stride_bytes, uint32_t base_stride_bytes, uint32_t num_instances, size_t dst_instance_stride_bytes)
{
    return pose_normals_launch(ctx, dst_norm_dev, base_norm_dev, num_vertices, dst_stride_bytes, base_stride_bytes, num_instances,
                               dst_instance_stride_bytes, 0);
}

int ogrecu_pose_blend_dev(ogrecu_ctx* ctx, float weight, const uint32_t* vertex_index_dev, const float* offsets_dev, uint32_t num_entries,
                          float* dst_dev, uint32_t dst_stride_bytes)
{
    OGRECU_REQUIRE(ctx, "ctx is NULL");
    if (num_entries == 0 || weight == 0.0f) return OGRECU_OK;
    OGRECU_REQUIRE(vertex_index_dev && offsets_dev && dst_dev, "NULL stream pointer");
    OGRECU_REQUIRE((((uintptr_t)dst_dev | dst_stride_bytes) & 3) == 0, "destination must be 4-byte aligned");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    pose_blend_kernel<<<dim3(div_up(num_entries, 256), 1), 256, 0, ctx->stream>>>(nullptr, weight, vertex_index_dev, offsets_dev, num_entries,
                                                                                 (uint8_t*)dst_dev, dst_stride_bytes, 0);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

int ogrecu_extrude_multi_dev(ogrecu_ctx* ctx, const float* lights_host, uint32_t n_lights, float dist, const float* src, float* dst,
                             size_t dst_light_stride_floats, size_t n)
{
    OGRECU_REQUIRE(ctx && lights_host, "NULL argument");
    OGRECU_REQUIRE(n_lights >= 1 && n_lights <= 8, "1..8 lights per launch");
    if (n == 0) return OGRECU_OK;
    OGRECU_REQUIRE(src && dst && ((((uintptr_t)src | (uintptr_t)dst) & 3) == 0), "streams must be non-NULL and 4-byte aligned");
    OGRECU_REQUIRE(n_lights == 1 || dst_light_stride_floats >= 3 * n, "destinations of different lights overlap");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    Lights L;
    memset(&L, 0, sizeof L);
    memcpy(L.l, lights_host, (size_t)n_lights * 16);
    const int al = aligned16(src) && aligned16(dst) && (n_lights == 1 || (dst_light_stride_floats & 3) == 0);
    extrude_multi_kernel<<<dim3(div_up(div_up(n, 4), 256), n_lights), 256, 0, ctx->stream>>>(L, n_lights, dist, src, dst, dst_light_stride_floats, n, al);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

int ogrecu_extrude_dev(ogrecu_ctx* ctx, const float* l, float dist, const float* src, float* dst, size_t n)
{
    OGRECU_REQUIRE(ctx && l, "NULL argument");
    if (n == 0) return OGRECU_OK;
    OGRECU_REQUIRE(src && dst && ((((uintptr_t)src | (uintptr_t)dst) & 3) == 0), "streams must be non-NULL and 4-byte aligned");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    const bool al = aligned16(src) && aligned16(dst);
    const bool point = !(l[3] == 0.0f);
    unsigned blocks = div_up(div_up(n, 4), 256);
    if (point) {
        if (al) extrude_kernel<true, true><<<blocks, 256, 0, ctx->stream>>>(l[0], l[1], l[2], dist, src, dst, n);
        else extrude_kernel<true, false><<<blocks, 256, 0, ctx->stream>>>(l[0], l[1], l[2], dist, src, dst, n);
    } else {
        if (al) extrude_kernel<false, true><<<blocks, 256, 0, ctx->stream>>>(l[0], l[1], l[2], dist, src, dst, n);
        else extrude_kernel<false, false><<<blocks, 256, 0, ctx->stream>>>(l[0], l[1], l[2], dist, src, dst, n);
    }
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

}  // extern "C"
