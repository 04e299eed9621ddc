// skin_math.cuh — the arithmetic of softwareVertexSkinning / softwareVertexMorph, shared by the
// generic (strided) kernels in optutil.cu and the tiled stream kernels in skin_tile.cu.
//
// Operation order replays OgreMain/src/OgreOptimisedUtilSSE.cpp (collapse :278-368, transform
// OgreSIMDHelper.h:204-227) with explicit round-to-nearest intrinsics so nothing is contracted
// into FMA; normalisation is the exact General form (OgreOptimisedUtilGeneral.cpp:261-276)
// where SSE takes the ~12-bit rsqrt estimate.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ogrecu {

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

__device__ __forceinline__ float4 mul4(float4 a, float w) { return make_float4(fmul(a.x, w), fmul(a.y, w), fmul(a.z, w), fmul(a.w, w)); }
__device__ __forceinline__ float4 madd4(float4 a, float w, float4 c)  // a*w + c, two roundings
{
    return make_float4(fadd(fmul(a.x, w), c.x), fadd(fmul(a.y, w), c.y), fadd(fmul(a.z, w), c.z), fadd(fmul(a.w, w), c.w));
}
__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float t)  // (b-a)*t + a
{
    return make_float4(fadd(fmul(fsub(b.x, a.x), t), a.x), fadd(fmul(fsub(b.y, a.y), t), a.y),
                       fadd(fmul(fsub(b.z, a.z), t), a.z), fadd(fmul(fsub(b.w, a.w), t), a.w));
}

// Collapse the weighted bone matrix (rows 0..2) of one vertex from a palette of 3 float4 rows
// per matrix.  NW==1 ignores the weight, NW==2 is lerp(M0, M1, w1) (OgreOptimisedUtilSSE.cpp:278-332),
// NW>=3 is w0*M0 then accumulate (:336-368).
template <int NW>
__device__ __forceinline__ void collapse_matrix(const float4* __restrict__ pal, const uint32_t idx4, const float* w,
                                                float4& r0, float4& r1, float4& r2)
{
    const float4* m0 = pal + 3 * (int)(idx4 & 0xffu);
    if (NW == 1) {
        r0 = m0[0]; r1 = m0[1]; r2 = m0[2];
    } else if (NW == 2) {
        const float4* m1 = pal + 3 * (int)((idx4 >> 8) & 0xffu);
        r0 = lerp4(m0[0], m1[0], w[1]); r1 = lerp4(m0[1], m1[1], w[1]); r2 = lerp4(m0[2], m1[2], w[1]);
    } else {
        r0 = mul4(m0[0], w[0]); r1 = mul4(m0[1], w[0]); r2 = mul4(m0[2], w[0]);
#pragma unroll
        for (int b = 1; b < NW; ++b) {
            const float4* mb = pal + 3 * (int)((idx4 >> (8 * b)) & 0xffu);
            r0 = madd4(mb[0], w[b], r0); r1 = madd4(mb[1], w[b], r1); r2 = madd4(mb[2], w[b], r2);
        }
    }
}

// position: (r.x*x + r.y*y) + (r.z*z + r.w)   (__MM_DOT4x3_PS, OgreSIMDHelper.h:204-217)
__device__ __forceinline__ void xform_pos(const float4& r0, const float4& r1, const float4& r2, float x, float y, float z,
                                          float& ox, float& oy, float& oz)
{
    ox = fadd(fadd(fmul(r0.x, x), fmul(r0.y, y)), fadd(fmul(r0.z, z), r0.w));
    oy = fadd(fadd(fmul(r1.x, x), fmul(r1.y, y)), fadd(fmul(r1.z, z), r1.w));
    oz = fadd(fadd(fmul(r2.x, x), fmul(r2.y, y)), fadd(fmul(r2.z, z), r2.w));
}

// normal: (r.x*x + r.y*y) + r.z*z (OgreSIMDHelper.h:220-227), then exact 1/sqrt where SSE takes
// _mm_rsqrt_ps (OgreOptimisedUtilSSE.cpp:590,695)
__device__ __forceinline__ void xform_norm(const float4& r0, const float4& r1, const float4& r2, float nx, float ny, float nz,
                                           float& ox, float& oy, float& oz)
{
    float a = fadd(fadd(fmul(r0.x, nx), fmul(r0.y, ny)), fmul(r0.z, nz));
    float b = fadd(fadd(fmul(r1.x, nx), fmul(r1.y, ny)), fmul(r1.z, nz));
    float c = fadd(fadd(fmul(r2.x, nx), fmul(r2.y, ny)), fmul(r2.z, nz));
    float len2 = fadd(fadd(fmul(a, a), fmul(b, b)), fmul(c, c));
    float rs = __frcp_rn(fsqrt(len2));   // correctly rounded 1/x == 1.0f/x
    ox = fmul(a, rs); oy = fmul(b, rs); oz = fmul(c, rs);
}

// dst = a + t*(b-a)  (OgreOptimisedUtilGeneral.cpp:249-256; bit-identical to SSE's (b-a)*t + a)
__device__ __forceinline__ float morph1(float a, float b, float t) { return fadd(a, fmul(t, fsub(b, a))); }

// Vector3::normalise (OgreVector.h:494-508): len = sqrt(((0+x*x)+y*y)+z*z); if len>0 scale by 1/len
__device__ __forceinline__ void normalise3(float& n0, float& n1, float& n2)
{
    float len = fsqrt(fadd(fadd(fadd(0.0f, fmul(n0, n0)), fmul(n1, n1)), fmul(n2, n2)));
    if (len > 0.0f) { float inv = __frcp_rn(len); n0 = fmul(n0, inv); n1 = fmul(n1, inv); n2 = fmul(n2, inv); }
}

}  // namespace ogrecu
