/*
 * ogre_oracle.c — CPU restatement of the reference algorithms on the hot path.
 * TEST INFRASTRUCTURE ONLY (see ogre_oracle.h for the rules and the pinning status).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
 * Every float operator below is one IEEE binary32 rounding, written in the reference's
 * association; do not "simplify".
 */
#include "ogre_oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* std::min / std::max exactly as libstdc++ defines them (NaN handling matters) */
static inline float std_minf(float a, float b) { return (b < a) ? b : a; }
static inline float std_maxf(float a, float b) { return (a < b) ? b : a; }

/* x86 cvttss2si: NaN and out-of-range give the "integer indefinite" value INT_MIN */
static inline int x86_f2i(float f)
{
    if (!(f < 2147483648.0f) || f < -2147483648.0f) return INT_MIN;
    return (int)f;
}

static inline const float* fptr(const void* p, size_t byte_off) { return (const float*)((const uint8_t*)p + byte_off); }

/* ======================================================================= OptimisedUtil */

void oracle_skin_general(const float* pSrcPos, float* pDestPos, const float* pSrcNorm, float* pDestNorm,
                         const float* pBlendWeight, const uint8_t* pBlendIndex,
                         const float* const* blendMatrices,
                         size_t srcPosStride, size_t destPosStride, size_t srcNormStride,
                         size_t destNormStride, size_t blendWeightStride, size_t blendIndexStride,
                         size_t numWeightsPerVertex, size_t numVertices)
{
    /* OgreOptimisedUtilGeneral.cpp:98-219: per bone transform, scale by weight, accumulate */
    for (size_t v = 0; v < numVertices; ++v) {
        float sx = pSrcPos[0], sy = pSrcPos[1], sz = pSrcPos[2];
        float nx = 0, ny = 0, nz = 0;
        if (pSrcNorm) { nx = pSrcNorm[0]; ny = pSrcNorm[1]; nz = pSrcNorm[2]; }
        float ap[3] = {0, 0, 0}, an[3] = {0, 0, 0};
        for (size_t b = 0; b < numWeightsPerVertex; ++b) {
            float weight = pBlendWeight[b];
            if (weight) {
                const float* m = blendMatrices[pBlendIndex[b]];
                for (int r = 0; r < 3; ++r) {
                    float t = ((m[4 * r + 0] * sx + m[4 * r + 1] * sy) + m[4 * r + 2] * sz) + m[4 * r + 3];
                    ap[r] += t * weight;
                }
                if (pSrcNorm) {
                    for (int r = 0; r < 3; ++r) {
                        float t = (m[4 * r + 0] * nx + m[4 * r + 1] * ny) + m[4 * r + 2] * nz;
                        an[r] += t * weight;
                    }
                }
            }
        }
        pDestPos[0] = ap[0]; pDestPos[1] = ap[1]; pDestPos[2] = ap[2];
        if (pSrcNorm) {
            /* Vector3::normalise, OgreVector.h:494-508 */
            float len2 = 0; len2 += an[0] * an[0]; len2 += an[1] * an[1]; len2 += an[2] * an[2];
            float len = sqrtf(len2);
            if (len > 0.0f) { float inv = 1.0f / len; an[0] *= inv; an[1] *= inv; an[2] *= inv; }
            pDestNorm[0] = an[0]; pDestNorm[1] = an[1]; pDestNorm[2] = an[2];
            pSrcNorm = fptr(pSrcNorm, srcNormStride);
            pDestNorm = (float*)fptr(pDestNorm, destNormStride);
        }
        pSrcPos = fptr(pSrcPos, srcPosStride);
        pDestPos = (float*)fptr(pDestPos, destPosStride);
        pBlendWeight = fptr(pBlendWeight, blendWeightStride);
        pBlendIndex += blendIndexStride;
    }
}

/* collapse one vertex' matrix, OgreOptimisedUtilSSE.cpp:278-368 (rows 0..2, 4 lanes each) */
static void collapse_matrix(float M[12], const float* w, const uint8_t* idx, const float* const* pal, size_t nw)
{
    const float* m0 = pal[idx[0]];
    switch (nw) {
    default:
    case 1: /* __COLLAPSE_MATRIX_W1 :315-319: weight ignored */
        for (int i = 0; i < 12; ++i) M[i] = m0[i];
        break;
    case 2: { /* __COLLAPSE_MATRIX_W2 :325-332: lerp(M0, M1, w[1]) = (M1-M0)*w1 + M0 */
        const float* m1 = pal[idx[1]];
        float w1 = w[1];
        for (int i = 0; i < 12; ++i) M[i] = (m1[i] - m0[i]) * w1 + m0[i];
        break;
    }
    case 3:
    case 4: { /* __COLLAPSE_MATRIX_W3/W4 :336-368: M0*w0, then += Mi*wi in order */
        for (int i = 0; i < 12; ++i) M[i] = m0[i] * w[0];
        for (size_t b = 1; b < nw; ++b) {
            const float* mb = pal[idx[b]];
            for (int i = 0; i < 12; ++i) M[i] = mb[i] * w[b] + M[i];
        }
        break;
    }
    }
}

void oracle_skin_sse(const float* pSrcPos, float* pDestPos, const float* pSrcNorm, float* pDestNorm,
                     const float* pBlendWeight, const uint8_t* pBlendIndex,
                     const float* const* blendMatrices,
                     size_t srcPosStride, size_t destPosStride, size_t srcNormStride,
                     size_t destNormStride, size_t blendWeightStride, size_t blendIndexStride,
                     size_t numWeightsPerVertex, size_t numVertices)
{
    for (size_t v = 0; v < numVertices; ++v) {
        float M[12];
        collapse_matrix(M, pBlendWeight, pBlendIndex, blendMatrices, numWeightsPerVertex);
        float x = pSrcPos[0], y = pSrcPos[1], z = pSrcPos[2];
        /* __MM_DOT4x3_PS -> __MM_ACCUM4_PS: (r0*x + r1*y) + (r2*z + r3), OgreSIMDHelper.h:204-217 */
        for (int r = 0; r < 3; ++r)
            pDestPos[r] = (M[4 * r + 0] * x + M[4 * r + 1] * y) + (M[4 * r + 2] * z + M[4 * r + 3]);
        if (pSrcNorm) {
            float nx = pSrcNorm[0], ny = pSrcNorm[1], nz = pSrcNorm[2];
            float n[3];
            /* __MM_DOT3x3_PS: (r0*x + r1*y) + r2*z, OgreSIMDHelper.h:220-227 */
            for (int r = 0; r < 3; ++r)
                n[r] = (M[4 * r + 0] * nx + M[4 * r + 1] * ny) + M[4 * r + 2] * nz;
            /* :692-696 length^2 = __MM_DOT3x3_PS(d3,d4,d5,d3,d4,d5); rsqrt -> exact here */
            float len2 = (n[0] * n[0] + n[1] * n[1]) + n[2] * n[2];
            float rs = 1.0f / sqrtf(len2);
            pDestNorm[0] = n[0] * rs; pDestNorm[1] = n[1] * rs; pDestNorm[2] = n[2] * rs;
            pSrcNorm = fptr(pSrcNorm, srcNormStride);
            pDestNorm = (float*)fptr(pDestNorm, destNormStride);
        }
        pSrcPos = fptr(pSrcPos, srcPosStride);
        pDestPos = (float*)fptr(pDestPos, destPosStride);
        pBlendWeight = fptr(pBlendWeight, blendWeightStride);
        pBlendIndex += blendIndexStride;
    }
}

void oracle_morph(float t, const float* pSrc1, const float* pSrc2, float* pDst, size_t pos1VSize,
                  size_t pos2VSize, size_t dstVSize, size_t numVertices, int morphNormals)
{
    /* OgreOptimisedUtilGeneral.cpp:236-283 */
    for (size_t i = 0; i < numVertices; ++i) {
        for (int k = 0; k < 3; ++k) pDst[k] = pSrc1[k] + t * (pSrc2[k] - pSrc1[k]);
        if (morphNormals) {
            float n[3];
            for (int k = 0; k < 3; ++k) n[k] = pSrc1[3 + k] + t * (pSrc2[3 + k] - pSrc1[3 + k]);
            float len2 = 0; len2 += n[0] * n[0]; len2 += n[1] * n[1]; len2 += n[2] * n[2];
            float len = sqrtf(len2);
            if (len > 0.0f) { float inv = 1.0f / len; n[0] *= inv; n[1] *= inv; n[2] *= inv; }
            pDst[3] = n[0]; pDst[4] = n[1]; pDst[5] = n[2];
        }
        pSrc1 = fptr(pSrc1, pos1VSize); pSrc2 = fptr(pSrc2, pos2VSize); pDst = (float*)fptr(pDst, dstVSize);
    }
}

void oracle_concat_affine_sse(const float* m, const float* src, float* dst, size_t n)
{
    /* OgreOptimisedUtilSSE.cpp:1529-1582; m3 is the literal fourth row of the base matrix */
    for (size_t i = 0; i < n; ++i, src += 16, dst += 16) {
        float out[16];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c) {
                float t0 = m[4 * r + 0] * src[0 + c];
                float t1 = m[4 * r + 1] * src[4 + c];
                float t2 = m[4 * r + 2] * src[8 + c];
                float t3 = m[4 * r + c] * m[12 + c];
                out[4 * r + c] = (t0 + t1) + (t2 + t3);
            }
        out[12] = 0; out[13] = 0; out[14] = 0; out[15] = 1;
        memcpy(dst, out, sizeof out);
    }
}

void oracle_concat_affine_general(const float* m, const float* src, float* dst, size_t n)
{
    /* OgreMatrix4.h:522-539 */
    for (size_t i = 0; i < n; ++i, src += 16, dst += 16) {
        float out[16];
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c)
                out[4 * r + c] = (m[4 * r + 0] * src[0 + c] + m[4 * r + 1] * src[4 + c]) + m[4 * r + 2] * src[8 + c];
            out[4 * r + 3] = ((m[4 * r + 0] * src[3] + m[4 * r + 1] * src[7]) + m[4 * r + 2] * src[11]) + m[4 * r + 3];
        }
        out[12] = 0; out[13] = 0; out[14] = 0; out[15] = 1;
        memcpy(dst, out, sizeof out);
    }
}

void oracle_face_normals(const float* positions, const void* triangles32, float* fn, size_t n)
{
    /* OgreOptimisedUtilGeneral.cpp:285-307, OgreVector.h:870-883; vertIndex[3] at byte 8 */
    const uint8_t* t = (const uint8_t*)triangles32;
    for (size_t i = 0; i < n; ++i, t += 32, fn += 4) {
        uint32_t vi[3];
        memcpy(vi, t + 8, 12);
        const float* v1 = positions + (size_t)vi[0] * 3;
        const float* v2 = positions + (size_t)vi[1] * 3;
        const float* v3 = positions + (size_t)vi[2] * 3;
        float ax = v2[0] - v1[0], ay = v2[1] - v1[1], az = v2[2] - v1[2];
        float bx = v3[0] - v1[0], by = v3[1] - v1[1], bz = v3[2] - v1[2];
        float nx = ay * bz - az * by;
        float ny = az * bx - ax * bz;
        float nz = ax * by - ay * bx;
        float d = 0; d += nx * v1[0]; d += ny * v1[1]; d += nz * v1[2];
        fn[0] = nx; fn[1] = ny; fn[2] = nz; fn[3] = -d;
    }
}

void oracle_light_facing_sse(const float* l, const float* fn, char* facings, size_t n)
{
    for (size_t i = 0; i < n; ++i, fn += 4) {
        float x = fn[0] * l[0], y = fn[1] * l[1], z = fn[2] * l[2], w = fn[3] * l[3];
        float dp = (x + z) + (y + w);           /* OgreOptimisedUtilSSE.cpp:1776-1786 */
        facings[i] = !(dp <= 0.0f);             /* _mm_cmpnle_ps :1793 */
    }
}

void oracle_light_facing_general(const float* l, const float* fn, char* facings, size_t n)
{
    for (size_t i = 0; i < n; ++i, fn += 4) {
        float dp = 0; dp += l[0] * fn[0]; dp += l[1] * fn[1]; dp += l[2] * fn[2]; dp += l[3] * fn[3];
        facings[i] = dp > 0;
    }
}

void oracle_extrude(const float* l, float dist, const float* src, float* dst, size_t n)
{
    /* OgreOptimisedUtilGeneral.cpp:321-365 */
    if (l[3] == 0.0f) {
        float d[3] = {-l[0], -l[1], -l[2]};
        float len2 = 0; len2 += d[0] * d[0]; len2 += d[1] * d[1]; len2 += d[2] * d[2];
        float len = sqrtf(len2);
        if (len > 0.0f) { float inv = 1.0f / len; d[0] *= inv; d[1] *= inv; d[2] *= inv; }
        d[0] *= dist; d[1] *= dist; d[2] *= dist;
        for (size_t i = 0; i < n; ++i, src += 3, dst += 3) {
            float a = src[0] + d[0], b = src[1] + d[1], c = src[2] + d[2];
            dst[0] = a; dst[1] = b; dst[2] = c;
        }
    } else {
        for (size_t i = 0; i < n; ++i, src += 3, dst += 3) {
            float d[3] = {src[0] - l[0], src[1] - l[1], src[2] - l[2]};
            float len2 = 0; len2 += d[0] * d[0]; len2 += d[1] * d[1]; len2 += d[2] * d[2];
            float len = sqrtf(len2);
            if (len > 0.0f) { float inv = 1.0f / len; d[0] *= inv; d[1] *= inv; d[2] *= inv; }
            d[0] *= dist; d[1] *= dist; d[2] *= dist;
            float a = src[0] + d[0], b = src[1] + d[1], c = src[2] + d[2];
            dst[0] = a; dst[1] = b; dst[2] = c;
        }
    }
}

/* ======================================================================= Tiny rasteriser */

void oracle_tiny_clear(uint8_t* colour, float* depth, uint32_t w, uint32_t h, unsigned buffers,
                       const float c[4], float depth_value)
{
    size_t npx = (size_t)w * h;
    if (buffers & 1u) {
        if (c[0] == 0.0f && c[1] == 0.0f && c[2] == 0.0f && c[3] == 0.0f) {
            memset(colour, 0, npx * 3);          /* Image::setTo ZERO shortcut, OgreImage.cpp:124-128 */
        } else {
            uint8_t px[3];
            for (int k = 0; k < 3; ++k) {        /* Bitwise::floatToFixed(v, 8), OgreBitwise.h:237-242 */
                float v = c[k];
                unsigned q = (v <= 0.0f) ? 0u : (v >= 1.0f) ? 255u : (unsigned)(v * 256.0f);
                px[k] = (uint8_t)q;
            }
            for (size_t p = 0; p < npx; ++p) { colour[3 * p] = px[0]; colour[3 * p + 1] = px[1]; colour[3 * p + 2] = px[2]; }
        }
    }
    if (buffers & 2u) {
        /* setTo(ColourValue(depth)) = (depth,1,1,1): never ZERO, PF_FLOAT32_R stores r */
        for (size_t p = 0; p < npx; ++p) depth[p] = depth_value;
    }
}

static inline void mat4_mul_vec4(const float* m, const float v[4], float out[4])
{
    /* OgreMatrix4.h:571-578 */
    for (int r = 0; r < 4; ++r)
        out[r] = ((m[4 * r + 0] * v[0] + m[4 * r + 1] * v[1]) + m[4 * r + 2] * v[2]) + m[4 * r + 3] * v[3];
}

typedef struct tri_setup {
    float c[9];       /* adjugate or inverse of ABC */
    float iw[3], z[3];
    float uv[3][2];
    float n[3][3];
    float vc[3][3];   /* vertex colours (beyond-Tiny state only) */
    int x0, x1, y0, y1;
} tri_setup;

/* Ogre::CompareFunction (OgreMain/include/OgreCommon.h:61-71) on depths; CMPF_LESS_EQUAL keeps Tiny's own form
 * `!(new > old)` (tinyrenderer.h:79), under which a NaN passes */
static int cmp_depth(uint32_t func, float a, float b)
{
    switch (func) {
    case 0: return 0;
    case 1: return 1;
    case 2: return a < b;
    case 3: return !(a > b);
    case 4: return a == b;
    case 5: return a != b;
    case 6: return a >= b;
    default: return a > b;
    }
}

static int cmp_alpha(uint32_t func, uint32_t a, uint32_t b)
{
    switch (func) {
    case 0: return 0;
    case 1: return 1;
    case 2: return a < b;
    case 3: return a <= b;
    case 4: return a == b;
    case 5: return a != b;
    case 6: return a >= b;
    default: return a > b;
    }
}

static uint64_t raster_triangle(const oracle_draw* d, const oracle_draw_state* st, const tri_setup* t, uint8_t* colour, float* depth, uint32_t W)
{
    const uint32_t en = st ? st->enable : 0u;
    uint64_t written = 0;
    const float inv255 = 1.0f / 255.0f;
    for (int x = t->x0; x <= t->x1; ++x) {
        for (int y = t->y0; y <= t->y1; ++y) {
            float Px = (float)x, Py = (float)y;
            /* barycentric(): inverse * (x,y,1), OgreMatrix3.h:321-327 */
            float bs0 = (t->c[0] * Px + t->c[1] * Py) + t->c[2] * 1.0f;
            float bs1 = (t->c[3] * Px + t->c[4] * Py) + t->c[5] * 1.0f;
            float bs2 = (t->c[6] * Px + t->c[7] * Py) + t->c[8] * 1.0f;
            float bc0 = bs0 * t->iw[0], bc1 = bs1 * t->iw[1], bc2 = bs2 * t->iw[2];
            float s = (bc0 + bc1) + bc2;
            float f = 1.0f / s;                                  /* OgreVector.h:659-664 */
            bc0 = f * bc0; bc1 = f * bc1; bc2 = f * bc2;
            float fd = 0.0f; fd += t->z[0] * bc0; fd += t->z[1] * bc1; fd += t->z[2] * bc2;
            if (bs0 < 0 || bs1 < 0 || bs2 < 0) continue;         /* tinyrenderer.h:74 */
            if (fd < 0.0) continue;                              /* :76 */
            size_t p = (size_t)y * W + (size_t)x;
            if (d->depth_test) {
                if (en & ORACLE_EXT_DEPTH_FUNC) { if (!cmp_depth(st->depth_func, fd, depth[p])) continue; }
                else if (fd > depth[p]) continue;                /* :79 */
            }
            uint32_t alpha = 255;                                /* no texture: white, alpha 1 */

            float col[4] = {1.0f, 1.0f, 1.0f, 1.0f};             /* default ColourValue: white */
            if (d->tex_rgba) {                                   /* DefaultShader::fragment .cpp:312-322 */
                float u = (t->uv[0][0] * bc0 + t->uv[1][0] * bc1) + t->uv[2][0] * bc2;
                float v = (t->uv[0][1] * bc0 + t->uv[1][1] * bc1) + t->uv[2][1] * bc2;
                int ix = x86_f2i(u * (float)(d->tex_w - 1u));    /* sample2D, OgreTinyRenderSystem.h:34-42 */
                int iy = x86_f2i(v * (float)(d->tex_h - 1u));
                int tw = (int)d->tex_w, th = (int)d->tex_h;
                ix = (tw + (ix % tw)) % tw;
                iy = (th + (iy % th)) % th;
                const uint8_t* tx = d->tex_rgba + ((size_t)iy * d->tex_w + (size_t)ix) * 4;
                if (tx[3] < 1) continue;                         /* alpha discard .cpp:318 */
                for (int k = 0; k < 4; ++k) col[k] = (float)tx[k] * inv255;
                alpha = tx[3];
            }
            if ((en & ORACLE_EXT_ALPHA_REJECT) && !cmp_alpha(st->alpha_func, alpha, st->alpha_ref)) continue;
            if (en & ORACLE_EXT_VERTEX_COLOUR)                   /* interpolated like uv, modulates rgb */
                for (int k = 0; k < 3; ++k) col[k] *= (t->vc[0][k] * bc0 + t->vc[1][k] * bc1) + t->vc[2][k] * bc2;
            if (d->lighting) {                                   /* .cpp:324-330 */
                float n[3];
                for (int k = 0; k < 3; ++k) n[k] = (t->n[0][k] * bc0 + t->n[1][k] * bc1) + t->n[2][k] * bc2;
                float dp = 0.0f; dp += n[0] * d->light_dir[0]; dp += n[1] * d->light_dir[1]; dp += n[2] * d->light_dir[2];
                float diffuse = std_maxf(0.f, dp);
                for (int k = 0; k < 4; ++k) { col[k] *= diffuse; col[k] += d->ambient[k]; }
            }
            uint8_t* dst = colour + 3 * p;
            if (d->blend_add)                                    /* tinyrenderer.h:86-87 */
                for (int k = 0; k < 3; ++k) col[k] += (float)dst[k] * inv255;
            for (int k = 0; k < 3; ++k) {
                float c = std_maxf(std_minf(col[k], 1.0f), 0.0f); /* Math::saturate -> Clamp, OgreMath.h:395,704-708 */
                c *= 255;
                dst[k] = (uint8_t)x86_f2i(c);                    /* truncating float->uchar, OgreVector.h:294-298 */
            }
            if (d->depth_write) depth[p] = fd;
            ++written;
        }
    }
    return written;
}

uint64_t oracle_tiny_draw(const oracle_draw* d, uint8_t* colour, float* depth, uint32_t W, uint32_t H)
{
    return oracle_tiny_draw_state(d, NULL, colour, depth, W, H);
}

uint64_t oracle_tiny_draw_state(const oracle_draw* d, const oracle_draw_state* st, uint8_t* colour, float* depth, uint32_t W, uint32_t H)
{
    uint64_t written = 0;
    const uint32_t en = st ? st->enable : 0u;
    /* Tiny: lists are culled by its fixed rule, strips never (.cpp:410); CULL_CLOCKWISE is that rule */
    const uint32_t cull = (en & ORACLE_EXT_CULL_MODE) ? st->cull_mode : (d->is_strip ? 1u : 2u);
    size_t tri = 0;
    /* OgreTinyRenderSystem.cpp:396-411 */
    for (size_t i = 0; i < d->draw_count; i += 3) {
        if (i && d->is_strip) i -= 2;
        if (i + 3 > d->draw_count) break;   /* the reference reads past the end here (UB); we stop */
        const size_t tri_index = tri++;
        tri_setup t;
        float X[3], Y[3];
        int finite = 1;
        for (int j = 0; j < 3; ++j) {
            size_t k = i + j;
            int64_t idx;
            if (d->index_type == 2) idx = ((const uint16_t*)d->indices)[k];   /* read unsigned; the reference reads int16 (UB >= 32768) */
            else if (d->index_type == 4) idx = ((const uint32_t*)d->indices)[k];
            else idx = (int64_t)k;
            const float* p = (const float*)(d->positions + (size_t)d->pos_stride * idx);
            float v[4] = {p[0], p[1], p[2], 1.0f}, clip[4], q[4];
            mat4_mul_vec4(d->mvp, v, clip);                       /* DefaultShader::vertex .cpp:302 */
            if (d->uvs) {
                const float* uv = (const float*)(d->uvs + (size_t)d->uv_stride * idx);
                float tv[4] = {uv[0], uv[1], 0.0f, 1.0f}, to[4];
                mat4_mul_vec4(d->tex_matrix, tv, to);
                t.uv[j][0] = to[0]; t.uv[j][1] = to[1];
            } else { t.uv[j][0] = t.uv[j][1] = 0; }
            if (d->normals) {
                const float* n = (const float*)(d->normals + (size_t)d->normal_stride * idx);
                for (int r = 0; r < 3; ++r)                       /* Matrix3 * Vector3, OgreMatrix3.h:321-327 */
                    t.n[j][r] = (d->mvit[4 * r + 0] * n[0] + d->mvit[4 * r + 1] * n[1]) + d->mvit[4 * r + 2] * n[2];
            } else { t.n[j][0] = t.n[j][1] = t.n[j][2] = 0; }
            if (en & ORACLE_EXT_VERTEX_COLOUR) {                  /* VET_UBYTE4_NORM bytes r,g,b,a -> ColourValue(const uchar*) */
                const uint8_t* c = st->colours + (size_t)st->colour_stride * idx;
                for (int r = 0; r < 3; ++r) t.vc[j][r] = (float)c[r] * (1.0f / 255.0f);
            }
            mat4_mul_vec4(d->viewport, clip, q);                  /* tinyrenderer.h:45 */
            float w = q[3];
            float fi = 1.0f / w;                                  /* pts[i] /= w, OgreVector.h:563-570 */
            X[j] = q[0] * fi; Y[j] = q[1] * fi; t.z[j] = q[2] * fi;
            t.iw[j] = 1 / w;                                      /* tinyrenderer.h:50 */
            if (!isfinite(X[j]) || !isfinite(Y[j]) || !isfinite(t.z[j]) || !isfinite(t.iw[j])) finite = 0;
        }
        if (!finite) continue;   /* reference behaviour undefined here (see header) */

        if (cull != 1u) {                                         /* tinyrenderer.h:37-39,55 */
            float ax = X[2] - X[0], ay = Y[2] - Y[0], bx = X[2] - X[1], by = Y[2] - Y[1];
            float cr = ax * by - ay * bx;
            /* beyond Tiny: a strip's odd triangles have the opposite winding */
            if ((en & ORACLE_EXT_CULL_MODE) && d->is_strip && (tri_index & 1u)) cr = -cr;
            if (cull == 2u ? cr > 0 : cr < 0) continue;
        }
        float mnx = FLT_MAX, mny = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX;
        float cx = (float)(W - 1u), cy = (float)(H - 1u);
        for (int j = 0; j < 3; ++j) {                             /* :58-65 */
            mnx = std_maxf(0.f, std_minf(mnx, X[j]));
            mxx = std_minf(cx, std_maxf(mxx, X[j]));
            mny = std_maxf(0.f, std_minf(mny, Y[j]));
            mxy = std_minf(cy, std_maxf(mxy, Y[j]));
        }
        /* (int) casts :68-69; beyond +-2^31 the reference loop bounds are undefined -> skip */
        if (!(mnx < 2147483648.0f) || !(mny < 2147483648.0f) || !(mxx > -2147483648.0f) || !(mxy > -2147483648.0f))
            continue;
        t.x0 = (int)mnx; t.x1 = (int)mxx; t.y0 = (int)mny; t.y1 = (int)mxy;
        if (en & ORACLE_EXT_SCISSOR) {                            /* pixels outside [left,right) x [top,bottom) are not touched */
            if ((int64_t)st->scissor[0] > t.x0) t.x0 = (int)st->scissor[0];
            if ((int64_t)st->scissor[1] > t.y0) t.y0 = (int)st->scissor[1];
            if ((int64_t)st->scissor[2] - 1 < t.x1) t.x1 = (int)st->scissor[2] - 1;
            if ((int64_t)st->scissor[3] - 1 < t.y1) t.y1 = (int)st->scissor[3] - 1;
        }
        if (t.x0 > t.x1 || t.y0 > t.y1) continue;

        /* Matrix3::Inverse of ABC, OgreMatrix3.cpp:155-202 (products with the 1-row are exact) */
        t.c[0] = Y[1] * 1.0f - Y[2] * 1.0f; t.c[1] = X[2] * 1.0f - X[1] * 1.0f; t.c[2] = X[1] * Y[2] - X[2] * Y[1];
        t.c[3] = Y[2] * 1.0f - Y[0] * 1.0f; t.c[4] = X[0] * 1.0f - X[2] * 1.0f; t.c[5] = X[2] * Y[0] - X[0] * Y[2];
        t.c[6] = Y[0] * 1.0f - Y[1] * 1.0f; t.c[7] = X[1] * 1.0f - X[0] * 1.0f; t.c[8] = X[0] * Y[1] - X[1] * Y[0];
        float det = (X[0] * t.c[0] + X[1] * t.c[3]) + X[2] * t.c[6];
        if (!(fabsf(det) <= 1e-06f)) {
            float g = 1.0f / det;
            for (int k = 0; k < 9; ++k) t.c[k] *= g;
        }
        written += raster_triangle(d, st, &t, colour, depth, W);
    }
    return written;
}


/* ============================================================ Skeleton bone palettes */

typedef struct { float w, x, y, z; } quat;
typedef struct { float x, y, z; } vec3;

/* Quaternion::operator*(const Quaternion&), OgreQuaternion.cpp:260-272 */
static quat q_mul(quat a, quat b)
{
    quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}

/* Quaternion::normalise, OgreQuaternion.h:213-218: *this = 1.0f / len * *this */
static quat q_normalise(quat q)
{
    float len = sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    float s = 1.0f / len;
    quat r = {s * q.w, s * q.x, s * q.y, s * q.z};
    return r;
}

/* Quaternion::Slerp, OgreQuaternion.cpp:373-413 (RIM_SPHERICAL): msEpsilon = 1e-03f (:42); Math::Sqrt / ATan2 / Sin are the
 * float overloads of libm (sqrtf, atan2f, sinf, OgreMath.h), so the result is bit-identical to the engine's only where both
 * run on the same libm — which is how it is pinned (tests/test_oracle_vs_ref.py, same process) */
static quat q_slerp(float t, quat p, quat q, int shortest)
{
    float c = p.w * q.w + p.x * q.x + p.y * q.y + p.z * q.z;         /* Quaternion::Dot */
    quat r;
    if (c < 0.0f && shortest) { c = -c; r.w = -q.w; r.x = -q.x; r.y = -q.y; r.z = -q.z; }
    else r = q;
    if (fabsf(c) < 1 - 1e-03f) {
        float sn = sqrtf(1 - c * c);
        float angle = atan2f(sn, c);
        float inv = 1.0f / sn;
        float c0 = sinf((1.0f - t) * angle) * inv;
        float c1 = sinf(t * angle) * inv;
        quat o = {c0 * p.w + c1 * r.w, c0 * p.x + c1 * r.x, c0 * p.y + c1 * r.y, c0 * p.z + c1 * r.z};
        return o;
    } else {
        float u = 1.0f - t;
        quat o = {u * p.w + t * r.w, u * p.x + t * r.x, u * p.y + t * r.y, u * p.z + t * r.z};
        return q_normalise(o);
    }
}

/* ---- IM_SPLINE: SimpleSpline (positions, scales) and RotationalSpline over all key frames of a track ---- */

/* Quaternion::Log, OgreQuaternion.cpp:325-357 */
static quat q_log(quat q)
{
    quat r = {0.0f, q.x, q.y, q.z};
    if (fabsf(q.w) < 1.0) {
        float nv = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
        float angle = atan2f(nv, q.w);
        float sn = sinf(angle);
        if (fabsf(sn) >= 1e-03f) {
            float c = angle / sn;
            r.x = c * q.x; r.y = c * q.y; r.z = c * q.z;
        }
    }
    return r;
}

/* Quaternion::Exp, OgreQuaternion.cpp:295-323 */
static quat q_exp(quat q)
{
    float angle = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
    float sn = sinf(angle);
    float ew = expf(q.w);
    quat r;
    r.w = ew * cosf(angle);
    if (fabsf(angle) >= 1e-03f) {
        float c = ew * (sn / angle);
        r.x = c * q.x; r.y = c * q.y; r.z = c * q.z;
    } else {
        r.x = ew * q.x; r.y = ew * q.y; r.z = ew * q.z;
    }
    return r;
}

/* Quaternion::Inverse, OgreQuaternion.cpp:274-287 */
static quat q_inverse(quat q)
{
    float n = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    quat r = {0.0f, 0.0f, 0.0f, 0.0f};
    if (n > 0.0) { float inv = 1.0f / n; r.w = q.w * inv; r.x = -q.x * inv; r.y = -q.y * inv; r.z = -q.z * inv; }
    return r;
}

static quat q_slerp(float t, quat p, quat q, int shortest);

/* Quaternion::Squad, OgreQuaternion.cpp:449-457 */
static quat q_squad(float t, quat p, quat a, quat b, quat q, int shortest)
{
    float st = 2.0f * t * (1.0f - t);
    quat sp = q_slerp(t, p, q, shortest);
    quat sq = q_slerp(t, a, b, 0);
    return q_slerp(st, sp, sq, 0);
}

/* RotationalSpline::recalcTangents, OgreRotationSpline.cpp (keys: 11 floats per key, rotation wxyz at [4..7]) */
static void rot_spline_tangents(const float* K, uint32_t n, quat* tan)
{
    if (n < 2) return;
#define KQ(i) ((quat){K[11 * (size_t)(i) + 4], K[11 * (size_t)(i) + 5], K[11 * (size_t)(i) + 6], K[11 * (size_t)(i) + 7]})
    const quat first = KQ(0), last = KQ(n - 1);
    const int closed = first.x == last.x && first.y == last.y && first.z == last.z && first.w == last.w;
    for (uint32_t i = 0; i < n; ++i) {
        const quat p = KQ(i);
        const quat invp = q_inverse(p);
        quat part1, part2;
        if (i == 0) {
            part1 = q_log(q_mul(invp, KQ(i + 1)));
            part2 = closed ? q_log(q_mul(invp, KQ(n - 2))) : q_log(q_mul(invp, p));
        } else if (i == n - 1) {
            part1 = closed ? q_log(q_mul(invp, KQ(1))) : q_log(q_mul(invp, p));
            part2 = q_log(q_mul(invp, KQ(i - 1)));
        } else {
            part1 = q_log(q_mul(invp, KQ(i + 1)));
            part2 = q_log(q_mul(invp, KQ(i - 1)));
        }
        const float m = -0.25f;
        quat pre = {m * (part1.w + part2.w), m * (part1.x + part2.x), m * (part1.y + part2.y), m * (part1.z + part2.z)};
        tan[i] = q_mul(p, q_exp(pre));
    }
#undef KQ
}

/* SimpleSpline::recalcTangents, OgreSimpleSpline.cpp:156-207, for the 3 floats at K[11 i + off] */
static void simple_spline_tangents(const float* K, uint32_t n, int off, vec3* tan)
{
    if (n < 2) return;
#define KP(i, c) K[11 * (size_t)(i) + off + (c)]
    const int closed = KP(0, 0) == KP(n - 1, 0) && KP(0, 1) == KP(n - 1, 1) && KP(0, 2) == KP(n - 1, 2);
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t a, b;                                  /* tangent = 0.5 * (P[a] - P[b]) */
        if (i == 0) { a = 1; b = closed ? n - 2 : 0; }
        else if (i == n - 1) { if (closed) { tan[i] = tan[0]; continue; } a = i; b = i - 1; }
        else { a = i + 1; b = i - 1; }
        tan[i].x = 0.5f * (KP(a, 0) - KP(b, 0)); tan[i].y = 0.5f * (KP(a, 1) - KP(b, 1)); tan[i].z = 0.5f * (KP(a, 2) - KP(b, 2));
    }
#undef KP
}

/* SimpleSpline::interpolate(fromIndex, t), OgreSimpleSpline.cpp:87-154: powers * hCoeffs * pt with Vector4 * Matrix4
 * (OgreMatrix4.h:588-596), every product and sum of the reference kept, zeros included */
static vec3 simple_spline_interpolate(const float* K, uint32_t n, int off, const vec3* tan, uint32_t from, float t)
{
    const float* p1 = K + 11 * (size_t)from + off;
    vec3 r = {p1[0], p1[1], p1[2]};
    if (from + 1 == n || t == 0.0f) return r;
    const float* p2 = K + 11 * (size_t)(from + 1) + off;
    if (t == 1.0f) { r.x = p2[0]; r.y = p2[1]; r.z = p2[2]; return r; }
    const float t2 = t * t, t3 = t2 * t;
    const float v0 = t3 * 2.0f + t2 * -3.0f + t * 0.0f + 1.0f * 1.0f;
    const float v1 = t3 * -2.0f + t2 * 3.0f + t * 0.0f + 1.0f * 0.0f;
    const float v2 = t3 * 1.0f + t2 * -2.0f + t * 1.0f + 1.0f * 0.0f;
    const float v3 = t3 * 1.0f + t2 * -1.0f + t * 0.0f + 1.0f * 0.0f;
    const vec3 a = tan[from], b = tan[from + 1];
    r.x = v0 * p1[0] + v1 * p2[0] + v2 * a.x + v3 * b.x;
    r.y = v0 * p1[1] + v1 * p2[1] + v2 * a.y + v3 * b.y;
    r.z = v0 * p1[2] + v1 * p2[2] + v2 * a.z + v3 * b.z;
    return r;
}

/* RotationalSpline::interpolate(fromIndex, t, useShortestPath) */
static quat rot_spline_interpolate(const float* K, uint32_t n, const quat* tan, uint32_t from, float t, int shortest)
{
    const float* k = K + 11 * (size_t)from;
    quat p = {k[4], k[5], k[6], k[7]};
    if (from + 1 == n || t == 0.0f) return p;
    quat q = {k[15], k[16], k[17], k[18]};
    if (t == 1.0f) return q;
    return q_squad(t, p, tan[from], tan[from + 1], q, shortest);
}

/* Quaternion::nlerp, OgreQuaternion.cpp:534-549 */
static quat q_nlerp(float t, quat p, quat q, int shortest)
{
    float c = p.w * q.w + p.x * q.x + p.y * q.y + p.z * q.z;
    quat r;
    if (c < 0.0f && shortest) {
        r.w = p.w + t * ((-q.w) - p.w); r.x = p.x + t * ((-q.x) - p.x); r.y = p.y + t * ((-q.y) - p.y); r.z = p.z + t * ((-q.z) - p.z);
    } else {
        r.w = p.w + t * (q.w - p.w); r.x = p.x + t * (q.x - p.x); r.y = p.y + t * (q.y - p.y); r.z = p.z + t * (q.z - p.z);
    }
    return q_normalise(r);
}

/* Quaternion::operator*(const Vector3&), OgreQuaternion.cpp:359-371 */
static vec3 q_rot(quat q, vec3 v)
{
    vec3 uv = {q.y * v.z - q.z * v.y, q.z * v.x - q.x * v.z, q.x * v.y - q.y * v.x};
    vec3 uuv = {q.y * uv.z - q.z * uv.y, q.z * uv.x - q.x * uv.z, q.x * uv.y - q.y * uv.x};
    float tw = 2.0f * q.w;
    uv.x *= tw; uv.y *= tw; uv.z *= tw;
    uuv.x *= 2.0f; uuv.y *= 2.0f; uuv.z *= 2.0f;
    vec3 r = {(v.x + uv.x) + uuv.x, (v.y + uv.y) + uuv.y, (v.z + uv.z) + uuv.z};
    return r;
}

typedef struct { vec3 p; quat q; vec3 s; } xform;

static xform load10(const float* f)
{
    xform t = {{f[0], f[1], f[2]}, {f[3], f[4], f[5], f[6]}, {f[7], f[8], f[9]}};
    return t;
}

/* Node::updateFromParentImpl, OgreNode.cpp:202-256, for every bone (parents first, whatever the handle order) */
static void derive_all(const oracle_skeleton* sk, const xform* local, xform* derived)
{
    uint32_t n = sk->num_bones, done = 0;
    uint8_t* ok = (uint8_t*)calloc(n ? n : 1, 1);
    while (done < n) {
        uint32_t before = done;
        for (uint32_t b = 0; b < n; ++b) {
            if (ok[b]) continue;
            int32_t p = sk->parent[b];
            if (p >= 0 && !ok[p]) continue;
            if (p < 0) {
                derived[b] = local[b];
            } else {
                const xform* P = &derived[p];
                int io = sk->inherit ? sk->inherit[2 * b] : 1, is = sk->inherit ? sk->inherit[2 * b + 1] : 1;
                derived[b].q = io ? q_mul(P->q, local[b].q) : local[b].q;
                if (is) { derived[b].s.x = P->s.x * local[b].s.x; derived[b].s.y = P->s.y * local[b].s.y; derived[b].s.z = P->s.z * local[b].s.z; }
                else derived[b].s = local[b].s;
                vec3 sp = {P->s.x * local[b].p.x, P->s.y * local[b].p.y, P->s.z * local[b].p.z};
                vec3 r = q_rot(P->q, sp);
                derived[b].p.x = r.x + P->p.x; derived[b].p.y = r.y + P->p.y; derived[b].p.z = r.z + P->p.z;
            }
            ok[b] = 1; ++done;
        }
        if (done == before) break;      /* cycle in the parent table: leave the rest untouched */
    }
    free(ok);
}

void oracle_skeleton_bind_inverse(const oracle_skeleton* sk, float* out10)
{
    uint32_t n = sk->num_bones;
    xform* local = (xform*)calloc(n ? n : 1, sizeof(xform));
    xform* derived = (xform*)calloc(n ? n : 1, sizeof(xform));
    for (uint32_t b = 0; b < n; ++b) local[b] = load10(sk->bind + 10 * b);
    derive_all(sk, local, derived);
    for (uint32_t b = 0; b < n; ++b) {
        float* o = out10 + 10 * b;
        const xform* d = &derived[b];
        o[0] = -d->p.x; o[1] = -d->p.y; o[2] = -d->p.z;
        /* Quaternion::Inverse, OgreQuaternion.cpp:274-287 */
        float norm = d->q.w * d->q.w + d->q.x * d->q.x + d->q.y * d->q.y + d->q.z * d->q.z;
        if (norm > 0.0f) { float inv = 1.0f / norm; o[3] = d->q.w * inv; o[4] = -d->q.x * inv; o[5] = -d->q.y * inv; o[6] = -d->q.z * inv; }
        else { o[3] = o[4] = o[5] = o[6] = 0.0f; }
        o[7] = 1.0f / d->s.x; o[8] = 1.0f / d->s.y; o[9] = 1.0f / d->s.z;
    }
    free(local); free(derived);
}

static int cmp_float(const void* a, const void* b)
{
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

/* mask: AnimationState blend masks, [nstates][num_bones] (state s scales its weight per bone: Animation::apply with a blend mask,
 * OgreAnimation.cpp:344-357), NULL = none.  manual_slot: [num_bones], >= 0 for a manually controlled bone = row of manual10 that
 * holds the local transform the bone has when setAnimationState runs (Skeleton::reset leaves such bones alone,
 * OgreSkeleton.cpp:233-240; tracks still apply on top of it), NULL = no manual bones. */
int oracle_skeleton_palette_ex(const oracle_skeleton* sk, const float* bind_inverse10, const oracle_animation* anims,
                               uint32_t nstates, const int32_t* anim, const float* time, const float* weight, const float* mask,
                               const int32_t* manual_slot, const float* manual10, float* out16)
{
    uint32_t n = sk->num_bones;
    xform* local = (xform*)malloc(sizeof(xform) * (n ? n : 1));
    xform* derived = (xform*)malloc(sizeof(xform) * (n ? n : 1));
    int rc = 0;
    /* Skeleton::reset -> Node::resetToInitialState (manually controlled bones keep what they hold) */
    for (uint32_t b = 0; b < n; ++b)
        local[b] = (manual_slot && manual_slot[b] >= 0) ? load10(manual10 + 10 * (size_t)manual_slot[b]) : load10(sk->bind + 10 * b);

    /* ANIMBLEND_AVERAGE: rebalance when the enabled weights sum to more than 1 (OgreSkeleton.cpp:174-195) */
    float weight_factor = 1.0f;
    if (sk->blend_mode == 0) {
        float total = 0.0f;
        for (uint32_t s = 0; s < nstates; ++s) total += weight[s];
        if (total > 1.0f) weight_factor = 1.0f / total;
    }
    for (uint32_t s = 0; s < nstates && !rc; ++s) {
        const oracle_animation* A = &anims[anim[s]];
        if (A->interpolation > 1 || A->rotation_interpolation > 1) { rc = 1; break; }
        const int spline = A->interpolation == 1;
        const int spherical = A->rotation_interpolation == 1;
        const float w_state = weight[s] * weight_factor;
        const float scl = 1.0f;                                  /* linked-skeleton scale: none */
        uint32_t nkeys = A->key_offset[A->num_tracks];
        /* Animation::buildKeyFrameTimeList: the sorted set of all key times */
        float* times = (float*)malloc(sizeof(float) * (nkeys ? nkeys : 1));
        for (uint32_t k = 0; k < nkeys; ++k) times[k] = A->keys[11 * k];
        qsort(times, nkeys, sizeof(float), cmp_float);
        uint32_t nt = 0;
        for (uint32_t k = 0; k < nkeys; ++k) if (nt == 0 || times[k] != times[nt - 1]) times[nt++] = times[k];
        /* Animation::_getTimeIndex */
        float tp = time[s];
        if (tp > A->length && A->length > 0.0f) tp = fmodf(tp, A->length);
        uint32_t gi = 0;
        if (nt) { while (gi + 1 < nt && times[gi] < tp) ++gi; }   /* lower_bound over [begin, end-1) */
        for (uint32_t t = 0; t < A->num_tracks; ++t) {
            uint32_t k0 = A->key_offset[t], nk = A->key_offset[t + 1] - k0;
            /* (*blendMask)[b->getHandle()] * weight, OgreAnimation.cpp:355 */
            const float w = mask ? mask[(size_t)s * n + (uint32_t)A->track_bone[t]] * w_state : w_state;
            if (nk == 0 || !w) continue;
            const float* K = A->keys + 11 * (size_t)k0;
            /* AnimationTrack::_buildKeyFrameIndexMap, entry gi only: the first local key whose time is >= times[gi]
             * (clamped to the last key) */
            uint32_t i = 0;
            for (uint32_t j = 0; j < gi; ++j)
                while (i + 1 < nk && K[11 * i] <= times[j]) ++i;
            /* getKeyFramesAtTime */
            const float* k2 = K + 11 * i;
            float t2 = k2[0];
            if (i != 0 && tp < K[11 * i]) --i;
            const float* k1 = K + 11 * i;
            float t1 = k1[0];
            float f = (t1 == t2) ? 0.0f : (tp - t1) / (t2 - t1);
            /* getInterpolatedKeyFrame (IM_LINEAR with RIM_LINEAR or RIM_SPHERICAL; IM_SPLINE) */
            vec3 tr, sc; quat ro;
            if (f == 0.0f) {
                tr.x = k1[1]; tr.y = k1[2]; tr.z = k1[3]; ro.w = k1[4]; ro.x = k1[5]; ro.y = k1[6]; ro.z = k1[7]; sc.x = k1[8]; sc.y = k1[9]; sc.z = k1[10];
            } else if (spline) {
                /* buildInterpolationSplines (OgreAnimationTrack.cpp:515-549) + the three interpolate calls (:438-450) */
                quat* rt = (quat*)calloc(nk, sizeof(quat));
                vec3* pt = (vec3*)calloc(nk, sizeof(vec3));
                vec3* st = (vec3*)calloc(nk, sizeof(vec3));
                rot_spline_tangents(K, nk, rt);
                simple_spline_tangents(K, nk, 1, pt);
                simple_spline_tangents(K, nk, 8, st);
                ro = rot_spline_interpolate(K, nk, rt, i, f, A->track_shortest[t]);
                tr = simple_spline_interpolate(K, nk, 1, pt, i, f);
                sc = simple_spline_interpolate(K, nk, 8, st, i, f);
                free(rt); free(pt); free(st);
            } else {
                quat a = {k1[4], k1[5], k1[6], k1[7]}, b = {k2[4], k2[5], k2[6], k2[7]};
                ro = spherical ? q_slerp(f, a, b, A->track_shortest[t]) : q_nlerp(f, a, b, A->track_shortest[t]);
                tr.x = k1[1] + f * (k2[1] - k1[1]); tr.y = k1[2] + f * (k2[2] - k1[2]); tr.z = k1[3] + f * (k2[3] - k1[3]);
                sc.x = k1[8] + f * (k2[8] - k1[8]); sc.y = k1[9] + f * (k2[9] - k1[9]); sc.z = k1[10] + f * (k2[10] - k1[10]);
            }
            /* applyToNode */
            xform* B = &local[A->track_bone[t]];
            B->p.x += tr.x * w * scl; B->p.y += tr.y * w * scl; B->p.z += tr.z * w * scl;
            quat ident = {1.0f, 0.0f, 0.0f, 0.0f};
            quat rot = spherical ? q_slerp(w, ident, ro, A->track_shortest[t]) : q_nlerp(w, ident, ro, A->track_shortest[t]);
            B->q = q_normalise(q_mul(B->q, rot));                 /* Node::rotate, TS_LOCAL */
            if (sc.x != 1.0f || sc.y != 1.0f || sc.z != 1.0f) {
                if (scl != 1.0f) { sc.x = 1.0f + scl * (sc.x - 1.0f); sc.y = 1.0f + scl * (sc.y - 1.0f); sc.z = 1.0f + scl * (sc.z - 1.0f); }
                else if (w != 1.0f) { sc.x = 1.0f + w * (sc.x - 1.0f); sc.y = 1.0f + w * (sc.y - 1.0f); sc.z = 1.0f + w * (sc.z - 1.0f); }
            }
            B->s.x *= sc.x; B->s.y *= sc.y; B->s.z *= sc.z;
        }
        free(times);
    }
    if (!rc) {
        derive_all(sk, local, derived);
        for (uint32_t b = 0; b < n; ++b) {
            /* Bone::_getOffsetTransform */
            xform inv = load10(bind_inverse10 + 10 * b);
            const xform* d = &derived[b];
            vec3 ls = {d->s.x * inv.s.x, d->s.y * inv.s.y, d->s.z * inv.s.z};
            quat lr = q_mul(d->q, inv.q);
            vec3 sp = {ls.x * inv.p.x, ls.y * inv.p.y, ls.z * inv.p.z};
            vec3 rp = q_rot(lr, sp);
            vec3 lt = {d->p.x + rp.x, d->p.y + rp.y, d->p.z + rp.z};
            /* Quaternion::ToRotationMatrix (OgreQuaternion.cpp:84-108) + makeTransform */
            float tx = lr.x + lr.x, ty = lr.y + lr.y, tz = lr.z + lr.z;
            float twx = tx * lr.w, twy = ty * lr.w, twz = tz * lr.w;
            float txx = tx * lr.x, txy = ty * lr.x, txz = tz * lr.x;
            float tyy = ty * lr.y, tyz = tz * lr.y, tzz = tz * lr.z;
            float* m = out16 + 16 * b;
            m[0] = ls.x * (1.0f - (tyy + tzz)); m[1] = ls.y * (txy - twz);          m[2] = ls.z * (txz + twy);           m[3] = lt.x;
            m[4] = ls.x * (txy + twz);          m[5] = ls.y * (1.0f - (txx + tzz)); m[6] = ls.z * (tyz - twx);           m[7] = lt.y;
            m[8] = ls.x * (txz - twy);          m[9] = ls.y * (tyz + twx);          m[10] = ls.z * (1.0f - (txx + tyy)); m[11] = lt.z;
            m[12] = 0.0f; m[13] = 0.0f; m[14] = 0.0f; m[15] = 1.0f;
        }
    }
    free(local); free(derived);
    return rc;
}

int oracle_skeleton_palette(const oracle_skeleton* sk, const float* bind_inverse10, const oracle_animation* anims,
                            uint32_t nstates, const int32_t* anim, const float* time, const float* weight, float* out16)
{
    return oracle_skeleton_palette_ex(sk, bind_inverse10, anims, nstates, anim, time, weight, NULL, NULL, NULL, out16);
}


/* ================================================== Silhouette / shadow-volume indices */

int oracle_shadow_use_mcguire(uint32_t num_groups, const float* b, const float* light4)
{
    if (num_groups > 1) return 0;
    if (light4[3] == 0.0f) return 1;                        /* LT_DIRECTIONAL */
    const float lx = light4[0], ly = light4[1], lz = light4[2];
    const float mnx = b[0], mny = b[1], mnz = b[2], mxx = b[3], mxy = b[4], mxz = b[5];
    if (mnx <= lx && lx <= mxx && mny <= ly && ly <= mxy && mnz <= lz && lz <= mxz) return 0;      /* contains */
    {
        float hx = (mxx - mnx) * 0.5f, hy = (mxy - mny) * 0.5f, hz = (mxz - mnz) * 0.5f;
        float cx = (mxx + mnx) * 0.5f - lx, cy = (mxy + mny) * 0.5f - ly, cz = (mxz + mnz) * 0.5f - lz;
        float hl = sqrtf(((0.0f + hx * hx) + hy * hy) + hz * hz), cl = sqrtf(((0.0f + cx * cx) + cy * cy) + cz * cz);
        if (!(hl / cl > 0.5f)) return 1;
    }
    enum { L = 1, R = 2, B = 4, T = 8, F = 16, N = 32 };
    unsigned mask = (lx < mnx ? L : 0) | (lx > mxx ? R : 0) | (ly < mny ? B : 0) | (ly > mxy ? T : 0) | (lz < mnz ? F : 0) | (lz > mxz ? N : 0);
    /* cornerMap order of the reference, corners as AxisAlignedBox::getCorner gives them */
    const unsigned cm[8] = {F | L | B, F | R | B, F | L | T, F | R | T, N | L | B, N | R | B, N | L | T, N | R | T};
    const float cx8[8] = {mnx, mxx, mnx, mxx, mnx, mxx, mnx, mxx};
    const float cy8[8] = {mny, mny, mxy, mxy, mny, mny, mxy, mxy};
    const float cz8[8] = {mnz, mnz, mnz, mnz, mxz, mxz, mxz, mxz};
    float ex[8], ey[8], ez[8];
    unsigned n = 0;
    for (int k = 0; k < 8; ++k)
        if ((mask & cm[k]) != 0 && (mask & cm[k]) != cm[k]) { ex[n] = cx8[k]; ey[n] = cy8[k]; ez[n] = cz8[k]; ++n; }
    float cosAngle = 1.0f;
    for (unsigned i0 = 0; i0 + 1 < n; ++i0)
        for (unsigned i1 = i0 + 1; i1 < n; ++i1) {
            float a[3] = {ex[i0] - lx, ey[i0] - ly, ez[i0] - lz}, c[3] = {ex[i1] - lx, ey[i1] - ly, ez[i1] - lz};
            float la = sqrtf(((0.0f + a[0] * a[0]) + a[1] * a[1]) + a[2] * a[2]);
            if (la > 0.0f) { float inv = 1.0f / la; a[0] *= inv; a[1] *= inv; a[2] *= inv; }
            float lc = sqrtf(((0.0f + c[0] * c[0]) + c[1] * c[1]) + c[2] * c[2]);
            if (lc > 0.0f) { float inv = 1.0f / lc; c[0] *= inv; c[1] *= inv; c[2] *= inv; }
            float cosAB = ((0.0f + a[0] * c[0]) + a[1] * c[1]) + a[2] * c[2];
            if (cosAngle > cosAB) cosAngle = cosAB;
        }
    return cosAngle < 0.5f ? 0 : 1;
}

static int silhouette_edge(const uint32_t* e, const int8_t* facings)
{
    int lf = facings[e[0]];
    return (e[4] && lf) || (!e[4] && (lf != facings[e[1]]));
}

long oracle_shadow_volume(uint32_t num_groups, const uint32_t* groups, const uint32_t* edges, const uint32_t* tris, const int8_t* facings,
                          int directional, int use_mcguire, uint32_t flags, int separate_light_cap, uint16_t* indices,
                          uint32_t buffer_indices, uint32_t* used, uint32_t* ranges)
{
    const int two_tris = !(directional && (flags & ORACLE_SRF_EXTRUDE_TO_INFINITY));
    /* pre-count */
    uint32_t pre = 0;
    for (uint32_t g = 0; g < num_groups; ++g) {
        const uint32_t* G = groups + 5 * g;
        int first = 1;
        for (uint32_t k = 0; k < G[1]; ++k) {
            const uint32_t* e = edges + 5 * (size_t)(G[0] + k);
            if (silhouette_edge(e, facings)) {
                pre += 3;
                if (two_tris) pre += 3;
                if (use_mcguire && (flags & ORACLE_SRF_INCLUDE_DARK_CAP)) { if (first) first = 0; else pre += 3; }
            }
        }
        uint32_t lit = 0;
        for (uint32_t t = 0; t < G[3]; ++t) lit += facings[G[2] + t] ? 1u : 0u;
        if (use_mcguire) { if (flags & ORACLE_SRF_INCLUDE_LIGHT_CAP) pre += 3 * lit; }
        else pre += lit * (((flags & ORACLE_SRF_INCLUDE_DARK_CAP) ? 3u : 0u) + ((flags & ORACLE_SRF_INCLUDE_LIGHT_CAP) ? 3u : 0u));
    }
    if (pre > buffer_indices) return -1;
    if (*used + pre > buffer_indices) *used = 0;
    uint16_t* p = indices + *used;
    uint32_t num = *used;
    for (uint32_t g = 0; g < num_groups; ++g) {
        const uint32_t* G = groups + 5 * g;
        uint32_t* R = ranges + 4 * g;
        R[0] = num; R[2] = 0; R[3] = 0;
        const uint32_t ovc = G[4];
        int first = 1;
        uint16_t dark_start = 0;
        for (uint32_t k = 0; k < G[1]; ++k) {
            const uint32_t* e = edges + 5 * (size_t)(G[0] + k);
            if (!silhouette_edge(e, facings)) continue;
            uint32_t v0 = e[2], v1 = e[3];
            if (!facings[e[0]]) { uint32_t t = v0; v0 = v1; v1 = t; }
            *p++ = (uint16_t)v1; *p++ = (uint16_t)v0; *p++ = (uint16_t)(v0 + ovc); num += 3;
            if (two_tris) { *p++ = (uint16_t)(v0 + ovc); *p++ = (uint16_t)(v1 + ovc); *p++ = (uint16_t)v1; num += 3; }
            if (use_mcguire && (flags & ORACLE_SRF_INCLUDE_DARK_CAP)) {
                if (first) { dark_start = (uint16_t)(v0 + ovc); first = 0; }
                else { *p++ = dark_start; *p++ = (uint16_t)(v1 + ovc); *p++ = (uint16_t)(v0 + ovc); num += 3; }
            }
        }
        if (!use_mcguire && (flags & ORACLE_SRF_INCLUDE_DARK_CAP)) {
            for (uint32_t t = 0; t < G[3]; ++t) {
                if (!facings[G[2] + t]) continue;
                const uint32_t* T = tris + 3 * (size_t)(G[2] + t);
                *p++ = (uint16_t)(T[1] + ovc); *p++ = (uint16_t)(T[0] + ovc); *p++ = (uint16_t)(T[2] + ovc); num += 3;
            }
        }
        if (flags & ORACLE_SRF_INCLUDE_LIGHT_CAP) {
            uint32_t* cnt = &R[1];
            if (separate_light_cap) { R[1] = num - R[0]; R[2] = num; cnt = &R[3]; }
            uint32_t start = separate_light_cap ? R[2] : R[0];
            for (uint32_t t = 0; t < G[3]; ++t) {
                if (!facings[G[2] + t]) continue;
                const uint32_t* T = tris + 3 * (size_t)(G[2] + t);
                *p++ = (uint16_t)T[0]; *p++ = (uint16_t)T[1]; *p++ = (uint16_t)T[2]; num += 3;
            }
            *cnt = num - start;
        } else {
            R[1] = num - R[0];
        }
    }
    *used = num;
    return (long)pre;
}


/* ======================================================================= Pose blending */

void oracle_pose_blend(float weight, const uint32_t* idx, const float* off3, uint32_t n, float* dst, uint32_t stride_floats)
{
    if (weight == 0.0f) return;
    for (uint32_t i = 0; i < n; ++i) {
        float* p = dst + (size_t)idx[i] * stride_floats;
        p[0] = p[0] + (off3[3 * i] * weight);
        p[1] = p[1] + (off3[3 * i + 1] * weight);
        p[2] = p[2] + (off3[3 * i + 2] * weight);
    }
}

void oracle_finalise_pose_normals(float* dst_norm, const float* src_norm, uint32_t n, uint32_t dst_stride_floats, uint32_t src_stride_floats)
{
    for (uint32_t v = 0; v < n; ++v) {
        float* d = dst_norm + (size_t)v * dst_stride_floats;
        const float* s = src_norm + (size_t)v * src_stride_floats;
        float x = d[0], y = d[1], z = d[2];
        float len = sqrtf(((0.0f + x * x) + y * y) + z * z);          /* Vector3::length */
        if (len + 1e-4f < 1.0f) {
            float bw = 1.0f - len;
            x += s[0] * bw; y += s[1] * bw; z += s[2] * bw;
        }
        float l2 = sqrtf(((0.0f + x * x) + y * y) + z * z);           /* Vector3::normalise */
        if (l2 > 0.0f) { float inv = 1.0f / l2; x *= inv; y *= inv; z *= inv; }
        d[0] = x; d[1] = y; d[2] = z;
    }
}

/* ====================================================================================================
 * Scaled read-back: Image::scale(FILTER_BILINEAR) on byte formats
 * ==================================================================================================== */

/* OgreImageResampler.h:297-311,316-322: position -> (first sample, 12-bit weight of the second sample) */
static void scale_axis(uint64_t pos_48, uint32_t extent, uint32_t* s1, uint32_t* s2, uint32_t* frac)
{
    uint32_t temp = (uint32_t)(pos_48 >> 36);
    temp = (temp > 0x800u) ? temp - 0x800u : 0u;
    *frac = temp & 0xFFFu;
    *s1 = temp >> 12;
    *s2 = (*s1 + 1u < extent - 1u) ? *s1 + 1u : extent - 1u;        /* std::min(s1 + 1, extent - 1) */
}

void oracle_scale_bilinear_bytes(const uint8_t* src, uint32_t src_w, uint32_t src_h, size_t src_pitch_px, uint32_t channels,
                                 uint8_t* dst, uint32_t dst_w, uint32_t dst_h, size_t dst_pitch_px)
{
    if (!dst_w || !dst_h || !src_w || !src_h) return;
    const uint64_t stepx = ((uint64_t)src_w << 48) / dst_w;           /* :299-300 */
    const uint64_t stepy = ((uint64_t)src_h << 48) / dst_h;
    uint64_t sy_48 = (stepy >> 1) - 1;                                /* :302 */
    for (uint32_t y = 0; y < dst_h; ++y, sy_48 += stepy) {
        uint32_t sy1, sy2, syf;
        scale_axis(sy_48, src_h, &sy1, &sy2, &syf);
        const size_t syoff1 = (size_t)sy1 * src_pitch_px, syoff2 = (size_t)sy2 * src_pitch_px;
        uint8_t* pdst = dst + (size_t)y * dst_pitch_px * channels;
        uint64_t sx_48 = (stepx >> 1) - 1;                            /* :316 */
        for (uint32_t x = 0; x < dst_w; ++x, sx_48 += stepx) {
            uint32_t sx1, sx2, sxf;
            scale_axis(sx_48, src_w, &sx1, &sx2, &sxf);
            const uint32_t sxfsyf = sxf * syf;                        /* :324 */
            for (uint32_t k = 0; k < channels; ++k) {                 /* :325-334, unsigned 32-bit arithmetic */
                uint32_t accum = src[(sx1 + syoff1) * channels + k] * (0x1000000u - (sxf << 12) - (syf << 12) + sxfsyf) +
                                 src[(sx2 + syoff1) * channels + k] * ((sxf << 12) - sxfsyf) +
                                 src[(sx1 + syoff2) * channels + k] * ((syf << 12) - sxfsyf) +
                                 src[(sx2 + syoff2) * channels + k] * sxfsyf;
                *pdst++ = (uint8_t)((accum + 0x800000u) >> 24);
            }
        }
    }
}
