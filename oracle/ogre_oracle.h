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

/* "Beyond Tiny" render state: state Tiny's setters ignore (OgreTinyRenderSystem.cpp:208-297),
 * honoured only when its ORACLE_EXT_* bit is set.  NOT a restatement of reference behaviour — the reference has none for
 * these — but the definition the CUDA path (ogrecu_draw_with_state) is tested against; with enable == 0 the function IS
 * oracle_tiny_draw.  Semantics:
 *   depth_func     Ogre::CompareFunction (OgreCommon.h:61-71) replacing Tiny's fixed `frag_depth > zbuf` rejection
 *                  (tinyrenderer.h:79) while depth_test is on; CMPF_LESS_EQUAL is that rule;
 *   cull_mode      Ogre::CullingMode (OgreCommon.h:164-172): CULL_CLOCKWISE = Tiny's rule `cross(p2-p0, p2-p1) > 0` is culled
 *                  (tinyrenderer.h:37-39), CULL_ANTICLOCKWISE = `< 0`, CULL_NONE; applies to strips too, whose odd triangles
 *                  count with the opposite sign (Tiny never culls strips, .cpp:410);
 *   scissor        left, top, right, bottom (exclusive): fragments outside are dropped (RenderSystem::setScissorTest);
 *   alpha_func/ref RenderSystem::_setAlphaRejectSettings: the fragment survives iff cmp(alpha, ref) on the texel's alpha byte
 *                  (255 without a texture), after Tiny's own `alpha == 0` discard (.cpp:318);
 *   colours        VES_DIFFUSE as VET_UBYTE4_NORM (bytes r,g,b,a): c_j = byte * (1/255), interpolated as uv is
 *                  ((c0*b0 + c1*b1) + c2*b2), multiplied into rgb after the texture and before lighting. */
#define ORACLE_EXT_DEPTH_FUNC 1u
#define ORACLE_EXT_CULL_MODE 2u
#define ORACLE_EXT_SCISSOR 4u
#define ORACLE_EXT_ALPHA_REJECT 8u
#define ORACLE_EXT_VERTEX_COLOUR 16u
typedef struct oracle_draw_state {
    uint32_t enable;
    uint32_t depth_func;
    uint32_t cull_mode;
    uint32_t alpha_func, alpha_ref;
    uint32_t scissor[4];
    uint32_t colour_stride;
    const uint8_t* colours;
} oracle_draw_state;
uint64_t oracle_tiny_draw_state(const oracle_draw* d, const oracle_draw_state* st, uint8_t* colour_rgb, float* depth, uint32_t w, uint32_t h);


/* ---- Skeleton bone-palette production ---- */

/* Flat description of a Skeleton: bones indexed by handle.  bind[b] = initial (binding-pose) position xyz,
 * orientation wxyz, scale xyz (Node::setInitialState); inherit[b] = {inheritOrientation, inheritScale}. */
typedef struct oracle_skeleton {
    uint32_t num_bones;
    uint32_t blend_mode;            /* 0 = ANIMBLEND_AVERAGE, 1 = ANIMBLEND_CUMULATIVE (OgreSkeleton.h:50-52) */
    const int32_t* parent;          /* [num_bones], -1 = root */
    const float* bind;              /* [num_bones][10] */
    const uint8_t* inherit;         /* [num_bones][2], NULL = all inherit */
} oracle_skeleton;

/* One Animation's node tracks.  keys[k] = time, translate xyz, rotation wxyz, scale xyz (TransformKeyFrame);
 * track t owns keys[key_offset[t] .. key_offset[t+1]) sorted by time. */
typedef struct oracle_animation {
    float length;
    uint32_t interpolation;             /* 0 = IM_LINEAR, 1 = IM_SPLINE (SimpleSpline / RotationalSpline through libm, as RIM_SPHERICAL) */
    uint32_t rotation_interpolation;    /* 0 = RIM_LINEAR, 1 = RIM_SPHERICAL (Quaternion::Slerp through libm: bit-identical to the engine
                                         * only on the same libm; the device path refuses it) */
    uint32_t num_tracks;
    const int32_t* track_bone;          /* [num_tracks] */
    const uint8_t* track_shortest;      /* [num_tracks] NodeAnimationTrack::getUseShortestRotationPath */
    const uint32_t* key_offset;         /* [num_tracks + 1] */
    const float* keys;                  /* [key_offset[num_tracks]][11] */
} oracle_animation;

/* Bone::setBindingPose, OgreMain/src/OgreBone.cpp:66-74 (on the derived transforms of the binding pose,
 * Node::updateFromParentImpl OgreNode.cpp:202-256): out[b] = inverse position xyz, inverse orientation wxyz,
 * inverse scale xyz. */
void oracle_skeleton_bind_inverse(const oracle_skeleton* sk, float* out10);

/* Skeleton::setAnimationState + Skeleton::_getBoneMatrices for ONE skeleton instance:
 * OgreSkeleton.cpp:162-217,402-424; Animation::apply/_getTimeIndex OgreAnimation.cpp:326-342,594-618,620-660;
 * AnimationTrack::getKeyFramesAtTime/_collectKeyFrameTimes/_buildKeyFrameIndexMap OgreAnimationTrack.cpp:57-126,175-206;
 * NodeAnimationTrack::getInterpolatedKeyFrame/applyToNode OgreAnimationTrack.cpp:373-512; Quaternion::nlerp
 * OgreQuaternion.cpp:534-549; Node::translate/rotate/scale OgreNode.cpp:378-428,579-584; Bone::_getOffsetTransform
 * OgreBone.cpp:88-105; TransformBaseReal::makeTransform OgreMatrix4.cpp:205-222.
 * The enabled animation states are given in application order: anim[s] indexes `anims`, time[s] is
 * AnimationState::getTimePosition(), weight[s] its weight.  bind_inverse = oracle_skeleton_bind_inverse's output
 * (or the reference's own).  out = num_bones x 16 floats (Affine3).  Returns 0, or 1 for an interpolation mode
 * that is not restated. */
int oracle_skeleton_palette(const oracle_skeleton* sk, const float* bind_inverse10, const oracle_animation* anims,
                            uint32_t nstates, const int32_t* anim, const float* time, const float* weight, float* out16);
/* ... with AnimationState blend masks ([nstates][num_bones], NULL = none) and manually controlled bones (manual_slot [num_bones]:
 * row of manual10 holding the bone's local transform, -1 = animated normally; NULL = none) */
int oracle_skeleton_palette_ex(const oracle_skeleton* sk, const float* bind_inverse10, const oracle_animation* anims,
                               uint32_t nstates, const int32_t* anim, const float* time, const float* weight, const float* mask,
                               const int32_t* manual_slot, const float* manual10, float* out16);

/* ---- Silhouette extraction / shadow-volume index generation ---- */

/* static isBoundOkForMcGuire + the useMcGuire decision, OgreMain/src/OgreShadowCaster.cpp:134-189,203-204.
 * bounds6 = light-cap bounds min xyz, max xyz; light4 = Light::getAs4DVector (w == 0: directional). */
int oracle_shadow_use_mcguire(uint32_t num_groups, const float* bounds6, const float* light4);

#define ORACLE_SRF_INCLUDE_LIGHT_CAP 1u
#define ORACLE_SRF_INCLUDE_DARK_CAP 2u
#define ORACLE_SRF_EXTRUDE_TO_INFINITY 4u

/* ShadowCaster::generateShadowVolume, OgreShadowCaster.cpp:191-515: the pre-count (:209-296), the wrap of the shared
 * index buffer (:321-324) and the emission (:339-506).
 *   groups[g] = {edge_start, edge_count, tri_start, tri_count, vertex_count}, edges[e] = {tri0, tri1, v0, v1, degenerate},
 *   tris[t] = vertIndex[3], facings[t] = EdgeData::triangleLightFacings.
 * Writes 16-bit indices into `indices` (capacity `buffer_indices`) from *used on (from 0 after a wrap), fills
 * ranges[g] = {indexStart, indexCount, lightCapStart, lightCapCount} and updates *used.  Returns the number of indices
 * emitted, or -1 when they do not fit the buffer at all (the reference then asks the SceneManager for a bigger one). */
long oracle_shadow_volume(uint32_t num_groups, const uint32_t* groups, const uint32_t* edges, const uint32_t* tris, const int8_t* facings,
                          int directional, int use_mcguire, uint32_t flags, int separate_light_cap, uint16_t* indices,
                          uint32_t buffer_indices, uint32_t* used, uint32_t* ranges);

/* ---- Pose blending ---- */
/* Mesh::softwareVertexPoseBlend, OgreMain/src/OgreMesh.cpp:1948-2007, one element (positions or normals): for every map
 * entry dst[idx] += offset * weight, as `*pdst = *pdst + (off * weight)`; nothing at all when weight == 0.
 * `dst` points at the element's first float of vertex 0, stride in floats. */
void oracle_pose_blend(float weight, const uint32_t* idx, const float* off3, uint32_t n, float* dst, uint32_t stride_floats);
/* Entity::finalisePoseNormals, OgreMain/src/OgreEntity.cpp:1207-1261: len = |n|; if (len + 1e-4f < 1) n += base * (1 - len);
 * n.normalise().  dst_norm / src_norm point at the normal of vertex 0 (accumulated pose normals / base mesh normals). */
void oracle_finalise_pose_normals(float* dst_norm, const float* src_norm, uint32_t n, uint32_t dst_stride_floats, uint32_t src_stride_floats);

/* ---- Scaled read-back ---- */
/* Image::scale(FILTER_BILINEAR) for the byte formats = LinearResampler_Byte<channels>::scale,
 * OgreMain/src/OgreImage.cpp:389-458 + OgreMain/src/OgreImageResampler.h:283-339 (2D): 16.48 fixed-point source steps,
 * the half-pixel shift and 12-bit blend weights taken from bits 36..63, 8.24 accumulation, +0x800000 rounding.
 * Pitches are in PIXELS (PixelBox::rowPitch).  src points at the top-left pixel of the source box. */
void oracle_scale_bilinear_bytes(const uint8_t* src, uint32_t src_w, uint32_t src_h, size_t src_pitch_px, uint32_t channels,
                                 uint8_t* dst, uint32_t dst_w, uint32_t dst_h, size_t dst_pitch_px);

#ifdef __cplusplus
}
#endif
#endif
