/*
 * ogre_oracle.h — CPU restatement of the reference algorithms on the hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use this; the product (libogrecu.so, the
 * plugin) never links, loads or calls it.
 *
 * Parity status: PINNED against the reference itself.  The reference's tests hold no
 * golden vector for this path, so the restatement is checked against the
 * unmodified reference compiled by oracle/build_ref.py (oracle/_ref: real
 * OptimisedUtilSSE/General, real TinyRenderSystem::_render) in tests/test_oracle_vs_ref.py,
 * and against the fixtures that script's sibling tests/golden/make_golden.py dumped from it.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * OGRE 14.6 tree).  All arithmetic is IEEE binary32, one rounding per written operator,
 * no FMA (build: -O2 -ffp-contract=off, x86-64 SSE2 scalar math).
 */
#ifndef OGRE_ORACLE_H
#define OGRE_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- OptimisedUtil (oracle B) ---- */

/* OptimisedUtilGeneral::softwareVertexSkinning, OgreMain/src/OgreOptimisedUtilGeneral.cpp:98-219 */
void oracle_skin_general(const float* src_pos, float* dst_pos, const float* src_norm, float* dst_norm,
                         const float* blend_weight, const uint8_t* blend_index,
                         const float* const* blend_matrices,
                         size_t src_pos_stride, size_t dst_pos_stride, size_t src_norm_stride,
                         size_t dst_norm_stride, size_t blend_weight_stride, size_t blend_index_stride,
                         size_t num_weights, size_t num_vertices);

/* SSE formulation (collapse the weighted matrix, then transform),
 * OgreMain/src/OgreOptimisedUtilSSE.cpp:278-421 (collapse macros), :516-602 (general SSE),
 * :606-729 (shared packed); dot-product association OgreMain/src/OgreSIMDHelper.h:204-227.
 * Normals are normalised with an exact 1.0f/sqrtf() where SSE uses the ~12-bit
 * _mm_rsqrt_ps estimate (:590,:695), length^2 = (x*x + y*y) + z*z (:692). */
void oracle_skin_sse(const float* src_pos, float* dst_pos, const float* src_norm, float* dst_norm,
                     const float* blend_weight, const uint8_t* blend_index,
                     const float* const* blend_matrices,
                     size_t src_pos_stride, size_t dst_pos_stride, size_t src_norm_stride,
                     size_t dst_norm_stride, size_t blend_weight_stride, size_t blend_index_stride,
                     size_t num_weights, size_t num_vertices);

/* OptimisedUtilGeneral::softwareVertexMorph, OgreOptimisedUtilGeneral.cpp:236-283 */
void oracle_morph(float t, const float* src1, const float* src2, float* dst, size_t stride1,
                  size_t stride2, size_t dst_stride, size_t num_vertices, int morph_normals);

/* concatenateAffineMatrices with the SSE association (t0+t1)+(t2+t3),
 * OgreOptimisedUtilSSE.cpp:1529-1582; rows 0..2 written, row 3 set to (0,0,0,1) */
void oracle_concat_affine_sse(const float* base16, const float* src, float* dst, size_t n);
/* General: Affine3*Affine3, OgreMain/include/OgreMatrix4.h:522-539 */
void oracle_concat_affine_general(const float* base16, const float* src, float* dst, size_t n);

/* calculateFaceNormals, OgreOptimisedUtilGeneral.cpp:285-307 + OgreVector.h:870-883
 * (SSE :1584-1726 is bit-identical up to the sign of zero) */
void oracle_face_normals(const float* positions, const void* triangles32, float* face_normals, size_t n);

/* calculateLightFacing, SSE summation (x+z)+(y+w) and !(dp<=0), OgreOptimisedUtilSSE.cpp:1728-1872 */
void oracle_light_facing_sse(const float* light4, const float* face_normals, char* facings, size_t n);
/* General: ((0+x)+y)+z)+w > 0, OgreOptimisedUtilGeneral.cpp:309-319 */
void oracle_light_facing_general(const float* light4, const float* face_normals, char* facings, size_t n);

/* extrudeVertices, General (exact normalise), OgreOptimisedUtilGeneral.cpp:321-365 */
void oracle_extrude(const float* light4, float extrude_dist, const float* src, float* dst, size_t n);

/* ---- Tiny rasteriser (oracle A) ---- */

typedef struct oracle_draw {
    const uint8_t* positions; uint32_t pos_stride;
    const uint8_t* uvs;       uint32_t uv_stride;      /* NULL => no uv      */
    const uint8_t* normals;   uint32_t normal_stride;  /* NULL => no normals */
    const void* indices;      uint32_t index_type;     /* 0, 2, 4            */
    uint32_t draw_count;
    uint32_t is_strip;
    uint32_t depth_test, depth_write, blend_add, lighting;
    const uint8_t* tex_rgba; uint32_t tex_w, tex_h;    /* NULL => white      */
    float mvp[16], tex_matrix[16], mvit[16], viewport[16];
    float light_dir[3];
    float ambient[4];
} oracle_draw;

/* TinyRenderSystem::clearFrameBuffer, RenderSystems/Tiny/src/OgreTinyRenderSystem.cpp:421-433
 * (Image::setTo OgreMain/src/OgreImage.cpp:121-136, Bitwise::floatToFixed OgreBitwise.h:237-242) */
void oracle_tiny_clear(uint8_t* colour_rgb, float* depth, uint32_t w, uint32_t h, unsigned buffers,
                       const float colour_rgba[4], float depth_value);

/* TinyRenderSystem::_render + DefaultShader + triangle(),
 * OgreTinyRenderSystem.cpp:299-333,349-414, src/tinyrenderer.h:29-95, SURVEY Appendix A.
 * Triangles whose screen-space vertices are not finite are skipped (the reference's
 * behaviour there is undefined: out-of-bounds writes).  Returns fragments written. */
uint64_t oracle_tiny_draw(const oracle_draw* d, uint8_t* colour_rgb, float* depth, uint32_t w, uint32_t h);

#ifdef __cplusplus
}
#endif
#endif
