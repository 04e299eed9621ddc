/*
 * ogrecu.h — C ABI of libogrecu.so, the B200-native (sm_100a) back end for OGRE's
 * software-rendering hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no OGRE, torch or C++ types.
 * Every entry point names the reference interface it serves (paths relative to the
 * OGRE 14.6 tree).  The C++ host plugin (ogre_b200/plugin: RenderSystem_CUDA.so +
 * CudaOptimisedUtil) is a thin adapter from OGRE's virtual interfaces to these calls;
 * INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - every function returns 0 on success, a non-zero ogrecu_status otherwise;
 *    ogrecu_last_error_string() describes the last failure on the calling thread.
 *  - there is NO CPU fallback: with no usable CUDA device ogrecu_init fails.
 *  - one context per device; a context is used from one thread at a time
 *    (OGRE's render thread, SURVEY §8b "Threading").
 *  - "dev" pointers are CUDA device pointers valid on the context's device; "host"
 *    pointers are ordinary process memory.  Functions suffixed _host take host pointers
 *    and perform the H2D/D2H staging themselves (the path OGRE's lock()/unlock() API
 *    forces); functions suffixed _dev work on resident data and are asynchronous on the
 *    context's stream.
 *  - all arithmetic is IEEE binary32 replaying the reference's operation order without
 *    FMA contraction.
 */
#ifndef OGRECU_H
#define OGRECU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define OGRECU_API __attribute__((visibility("default")))
#else
#define OGRECU_API
#endif

#define OGRECU_ABI_VERSION 6

typedef enum ogrecu_status {
    OGRECU_OK = 0,
    OGRECU_ERR_NO_DEVICE = 1,     /* no CUDA device / driver: there is no CPU fallback   */
    OGRECU_ERR_CUDA = 2,          /* a CUDA runtime call failed (see last_error_string)  */
    OGRECU_ERR_INVALID = 3,       /* bad argument (NULL, out of range, unsupported mode) */
    OGRECU_ERR_OOM = 4,           /* device or pinned-host allocation failed             */
    OGRECU_ERR_UNSUPPORTED = 5    /* e.g. operation type other than triangle list/strip  */
} ogrecu_status;

typedef struct ogrecu_ctx ogrecu_ctx;   /* one CUDA device + stream + scratch arenas     */
typedef struct ogrecu_buf ogrecu_buf;   /* HardwareBuffer: device storage + pinned shadow */
typedef struct ogrecu_tex ogrecu_tex;   /* RGBA8 texture, 0 mips (TinyTexture)            */
typedef struct ogrecu_rt  ogrecu_rt;    /* render target: RGB8 colour + f32 depth planes  */

/* THREADING CONTRACT.  A context serves ONE thread at a time: OGRE issues every RenderSystem / OptimisedUtil call from the
 * render thread (Root::renderOneFrame), and so must any other caller — the launch stream, the scratch arenas, the pinned
 * staging block and ring, and the recorded command lists of the targets are not synchronised.  The one exception is the
 * buffer registry: ogrecu_buf_create / ogrecu_buf_destroy / ogrecu_buf_resolve take an internal lock, because OGRE creates
 * and destroys hardware buffers from resource-loading threads (HardwareBufferManager's mutex-protected sets,
 * OgreMain/src/OgreHardwareBufferManager.cpp).  Mapping, unmapping and using a buffer stay render-thread operations.
 * Different contexts (devices) may be driven from different threads. */

/* ------------------------------------------------------------------ context -------- */

OGRECU_API int ogrecu_abi_version(void);
OGRECU_API const char* ogrecu_last_error_string(void);
OGRECU_API int ogrecu_device_count(int* count);

/* Create a context on CUDA device `device`.  Fails with OGRECU_ERR_NO_DEVICE when the
 * driver or device is missing.  (TinyRenderSystem ctor / _createRenderWindow first-time
 * init, RenderSystems/Tiny/src/OgreTinyRenderSystem.cpp:27-46,146-194.) */
OGRECU_API int ogrecu_init(int device, ogrecu_ctx** out_ctx);
OGRECU_API int ogrecu_shutdown(ogrecu_ctx* ctx);
OGRECU_API int ogrecu_device(const ogrecu_ctx* ctx, int* device);

/* DEVICE GROUP: one context that drives `num_devices` GPUs of this node from one process (what a C++ OGRE application that
 * loads the plugin has: RenderSystem config options "Devices" / "Split Frame", OgreMain/src/OgreRenderSystem.cpp:834 is the hook).
 * devices[0] is the ROOT: buffers, textures, every OptimisedUtil / skeleton / shadow entry point and the host-facing copies run
 * there, exactly as on a single-device context.  The other devices are PEERS that only run raster passes:
 *   - ogrecu_buf / ogrecu_tex objects get a replica on every peer; a replica is brought up to date by a peer copy over NVLink
 *     (only the byte range that changed on the root since the last time) when a draw that runs on that peer references it;
 *     device pointers that are not inside an ogrecu_buf are read in place through peer access;
 *   - a render target lives on one device (ogrecu_rt_create_placed, placement = index into `devices`): its draws, clears and
 *     read-backs run on that device, so independent targets (the batched offscreen views of BASELINE config 4,
 *     ogrecu_draw_views) spread over the group without any exchange;
 *   - or it is SPLIT (placement = OGRECU_PLACE_SPLIT): the planes live on the root, the frame's tile rows are dealt out in
 *     interleaved bands ("split_band_tiles" tile rows each, band b -> device b mod num_devices), every device runs the per-triangle
 *     set-up for its own bands and rasterises them STRAIGHT INTO THE ROOT'S PLANES through peer stores — there is no composite
 *     pass and no exchange step; the root's stream continues when all devices have finished (BASELINE config 5).
 * Results are identical to a single-device context's.  Requires peer access between all members (OGRECU_ERR_UNSUPPORTED
 * otherwise).  All members are shut down with the root.  The threading contract is that of a single context. */
OGRECU_API int ogrecu_init_devices(int num_devices, const int* devices, ogrecu_ctx** out_ctx);
OGRECU_API int ogrecu_num_devices(const ogrecu_ctx* ctx, int* num_devices);

/* Launch on an externally owned cudaStream_t (e.g. the bench's timing stream).  Passing
 * NULL restores the context's own stream. */
OGRECU_API int ogrecu_set_stream(ogrecu_ctx* ctx, void* cuda_stream);
OGRECU_API void* ogrecu_get_stream(ogrecu_ctx* ctx);
OGRECU_API int ogrecu_sync(ogrecu_ctx* ctx);
/* number of kernels this context has launched since creation (bench "gpu_launches") */
OGRECU_API uint64_t ogrecu_launch_count(const ogrecu_ctx* ctx);

/* RenderSystem::setConfigOption-style integer knobs (OgreMain/include/OgreRenderSystem.h:246).
 * Kernel-selection switches for A/B measurement only — every value runs on the GPU:
 *   "skin_mode"  0 = auto (tiled stream kernel with TMA where alignment allows), 1 = tiled kernel
 *                with cooperative LDG/STG only, 2 = one-thread-per-vertex kernel
 *   "skin_ipc"   instances per CTA of the tiled kernel, 0 = auto
 *   "optutil_batch" 1 = ogrecu_software_vertex_skinning_host calls whose operands all live in ogrecu_bufs are QUEUED (descriptor +
 *                a copy of the palette) and run as one batched launch at the next point that needs their results: the flush of
 *                a target, a map of one of their buffers, any other OptimisedUtil host call, a memcpy, ogrecu_sync,
 *                ogrecu_optutil_flush.  What Entity::updateAnimation issues for a crowd (7 calls per Sinbad entity,
 *                OgreMain/src/OgreEntity.cpp:747-911) becomes one launch per frame; results are bit-identical.  0 = default:
 *                every call launches at once.  Callers that mix deferred host calls with *_dev entry points on raw device
 *                pointers call ogrecu_optutil_flush in between.
 *   "lane_reuse_repack" 1 = ogrecu_anim_frame_dev skips its palette repack launch and reads the previous call's (valid only while
 *                the palettes are unchanged: lets a benchmark time the lane kernel alone), 0 = default
 *   "split_band_tiles" tile rows (16 pixels each) per interleaved band of the split targets created afterwards (default 4)
 *   "raster_mode" 0 = auto (per tile: order-independent visibility resolve where it is provably
 *                equivalent to Tiny's in-order loop, else in-order), 1 = always in-order
 *   "raster_xf"  1 = transform every vertex once per flush when the draw gives vertex_count (default),
 *                0 = once per index like Tiny; identical results
 *   "skel_mode"  0 = bone-palette kernel in use (16 instances of one bone per half-warp), 1 = first version (a lane per
 *                bone); identical results
 *   "optutil_sync" 0 (default) = an OptimisedUtil *_host call whose outputs all resolve to ogrecu_buf shadows returns without
 *                waiting for the device (the next map / read-back / flush of those buffers is the sync point; small host
 *                inputs such as the bone palette travel through a ring of pinned slots), 1 = every call waits; identical results
 * Unknown names return OGRECU_ERR_INVALID. */
OGRECU_API int ogrecu_set_config_option(ogrecu_ctx* ctx, const char* name, int value);

/* Profiler ranges (RenderSystem::beginProfileEvent / endProfileEvent / markProfileEvent,
 * OgreMain/include/OgreRenderSystem.h:1104-1115; the GL render systems map them to KHR_debug groups): NVTX ranges and markers on
 * the calling thread, visible in Nsight Systems / Compute around the launches they bracket.  The library also opens its own
 * ranges around a target's flush and the batched OptimisedUtil flush.  No-ops when no tool is attached. */
OGRECU_API int ogrecu_profile_begin(ogrecu_ctx* ctx, const char* name);
OGRECU_API int ogrecu_profile_end(ogrecu_ctx* ctx);
OGRECU_API int ogrecu_profile_mark(ogrecu_ctx* ctx, const char* name);

/* raw device / pinned memory for callers that keep data resident */
OGRECU_API int ogrecu_malloc(ogrecu_ctx* ctx, size_t bytes, void** dev_ptr);
OGRECU_API int ogrecu_free(ogrecu_ctx* ctx, void* dev_ptr);
OGRECU_API int ogrecu_host_alloc(size_t bytes, void** host_ptr);    /* pinned */
OGRECU_API int ogrecu_host_free(void* host_ptr);
OGRECU_API int ogrecu_memcpy_h2d(ogrecu_ctx* ctx, void* dev, const void* host, size_t bytes);
OGRECU_API int ogrecu_memcpy_d2h(ogrecu_ctx* ctx, void* host, const void* dev, size_t bytes);   /* waits for the copy */
/* enqueue only (host must be pinned); the data is there after ogrecu_sync or an event the caller records on the stream */
OGRECU_API int ogrecu_memcpy_d2h_async(ogrecu_ctx* ctx, void* host, const void* dev, size_t bytes);
/* device to device on the launch stream (HardwareBuffer::copyData between two of our buffers, OgreHardwareBuffer.h:326-336) */
OGRECU_API int ogrecu_memcpy_d2d(ogrecu_ctx* ctx, void* dst, const void* src, size_t bytes);
/* host -> device on a stream of the CALLER's (a copy stream), not the context's launch stream: lets a renderer upload
 * frame k+1's inputs (bone palettes, weights) while frame k's kernels run.  Ordering against the launch stream is the
 * caller's job (events).  host should be pinned for the copy to be asynchronous. */
OGRECU_API int ogrecu_memcpy_h2d_on(ogrecu_ctx* ctx, void* cuda_stream, void* dev, const void* host, size_t bytes);

/* ------------------------------------------------------------ hardware buffers ----- */
/* HardwareBuffer delegate (OgreMain/include/OgreHardwareBuffer.h:181-186,279-330):
 * lock() returns a pinned host shadow; unlock() after a write uploads the dirty range on
 * the context stream; a lock with read intent first downloads device-side writes
 * (skinning results).  The shadow address range is registered so OptimisedUtil calls that
 * arrive with pointers obtained from lock() resolve to the device copy with no staging. */
enum { OGRECU_MAP_READ = 1, OGRECU_MAP_WRITE = 2, OGRECU_MAP_DISCARD = 4 };

OGRECU_API int ogrecu_buf_create(ogrecu_ctx* ctx, size_t bytes, ogrecu_buf** out);
OGRECU_API int ogrecu_buf_destroy(ogrecu_buf* buf);
OGRECU_API int ogrecu_buf_map(ogrecu_buf* buf, size_t offset, size_t length, int mode, void** host_ptr);
OGRECU_API int ogrecu_buf_unmap(ogrecu_buf* buf);
OGRECU_API int ogrecu_buf_device_ptr(ogrecu_buf* buf, void** dev_ptr);
OGRECU_API size_t ogrecu_buf_size(const ogrecu_buf* buf);
/* mark [offset, offset+length) as written on the device (shadow stale) */
OGRECU_API int ogrecu_buf_mark_device_dirty(ogrecu_buf* buf, size_t offset, size_t length);
/* host shadow pointer -> owning buffer + offset; *out_buf = NULL when not ours */
OGRECU_API int ogrecu_buf_resolve(ogrecu_ctx* ctx, const void* host_ptr, ogrecu_buf** out_buf, size_t* out_offset);

/* ------------------------------------------------ skeleton bone palettes ------ */
/* Bone-palette production for many instances of one skeleton, on the device: what
 * SkeletonInstance::setAnimationState + Skeleton::_getBoneMatrices do per entity on the CPU
 * (OgreMain/src/OgreSkeleton.cpp:162-217,402-424, called from Entity::cacheBoneMatrices OgreEntity.cpp:1307-1330;
 * Animation::apply OgreAnimation.cpp:326-342, NodeAnimationTrack::applyToNode OgreAnimationTrack.cpp:473-512,
 * Node::updateFromParentImpl OgreNode.cpp:202-256, Bone::_getOffsetTransform OgreBone.cpp:88-105).  With it a
 * frame uploads the animation STATES of the instances (a few bytes each) instead of 64 B x bones palettes.
 * Results are bit-identical to the engine's for IM_LINEAR / RIM_LINEAR animations (tests/test_skeleton_gpu.py).  IM_SPLINE and
 * RIM_SPHERICAL (OgreAnimationTrack.cpp:417-450, Quaternion::Slerp / Squad, SimpleSpline / RotationalSpline) are evaluated too:
 * spline tangents are made once on the host at ogrecu_skel_add_animation, translations and scales stay exact, rotations go through
 * the device's sinf / atan2f and agree with the engine within 2e-5 of a matrix's largest element (the stated tolerance of
 * those modes).  Blend masks and manually controlled bones: ogrecu_skel_palettes_ex_dev. */
typedef struct ogrecu_skel ogrecu_skel;

typedef struct ogrecu_skeleton_desc {
    uint32_t num_bones;             /* <= 256 (OGRE_MAX_NUM_BONES), indexed by bone handle                           */
    uint32_t blend_mode;            /* 0 = ANIMBLEND_AVERAGE, 1 = ANIMBLEND_CUMULATIVE (OgreSkeleton.h:50-52)        */
    const int32_t* parent;          /* [num_bones] parent handle, -1 = root bone                                     */
    const float*   bind;            /* [num_bones][10] initial position xyz, orientation wxyz, scale xyz             */
    const uint8_t* inherit;         /* [num_bones][2] inheritOrientation, inheritScale; NULL = all inherit           */
    const float*   bind_inverse;    /* [num_bones][10] Bone::setBindingPose products (inverse position, orientation, */
                                    /* scale); NULL = computed here the way Bone::setBindingPose does                */
} ogrecu_skeleton_desc;

typedef struct ogrecu_animation_desc {
    float length;                       /* Animation::getLength                                                     */
    uint32_t interpolation;             /* 0 = IM_LINEAR, 1 = IM_SPLINE (OgreAnimation.h:70-76)                     */
    uint32_t rotation_interpolation;    /* 0 = RIM_LINEAR, 1 = RIM_SPHERICAL (OgreAnimation.h:79-90)                */
    uint32_t num_tracks;                /* node tracks, at most one per bone                                        */
    const int32_t*  track_bone;         /* [num_tracks] bone handle                                                 */
    const uint8_t*  track_shortest;     /* [num_tracks] getUseShortestRotationPath                                  */
    const uint32_t* key_offset;         /* [num_tracks + 1] key range of each track                                 */
    const float*    keys;               /* [key_offset[num_tracks]][11] time, translate xyz, rotation wxyz, scale xyz */
} ogrecu_animation_desc;

/* All desc pointers are HOST pointers, copied by the call. */
OGRECU_API int ogrecu_skel_create(ogrecu_ctx* ctx, const ogrecu_skeleton_desc* desc, ogrecu_skel** out);
OGRECU_API int ogrecu_skel_destroy(ogrecu_skel* skel);
OGRECU_API int ogrecu_skel_add_animation(ogrecu_skel* skel, const ogrecu_animation_desc* desc, uint32_t* out_index);
OGRECU_API int ogrecu_skel_set_blend_mode(ogrecu_skel* skel, uint32_t blend_mode);
/* the binding-pose inverses in use (computed or given), [num_bones][10], to a HOST array */
OGRECU_API int ogrecu_skel_get_bind_inverse(ogrecu_skel* skel, float* out10_host);
/* Skeleton::setAnimationState + Skeleton::_getBoneMatrices (OgreMain/src/OgreSkeleton.cpp:162-217,402-424) for
 * num_instances skeleton instances, each with `num_states` animation-state slots applied in slot order:
 * anim_dev[i*num_states + s] = animation index (< 0: slot disabled), time_dev = AnimationState::getTimePosition,
 * weight_dev = AnimationState::getWeight (all DEVICE arrays).  Writes num_bones Affine3 per instance to
 * palettes_dev + i*palette_stride_floats (16-byte aligned, stride a multiple of 4 floats) — the layout
 * ogrecu_skin_instanced_dev / ogrecu_anim_instanced_dev read.  One launch, asynchronous on the stream. */
OGRECU_API int ogrecu_skel_palettes_dev(ogrecu_skel* skel, uint32_t num_instances, uint32_t num_states,
                                        const int32_t* anim_dev, const float* time_dev, const float* weight_dev,
                                        float* palettes_dev, size_t palette_stride_floats);

/* ... with what an AnimationStateSet / SkeletonInstance can carry besides (time, weight):
 *   - AnimationState blend masks (OgreAnimationState.h createBlendMask / setBlendMaskEntry; Animation::apply with a mask,
 *     OgreAnimation.cpp:344-357: a track's weight becomes mask[bone handle] * weight): mask_dev[(i * mask_instance_stride_floats) +
 *     s * num_bones + b] for instance i, state slot s, bone b; stride 0 = one set shared by all instances; a state without a mask
 *     has ones;
 *   - manually controlled bones (Bone::setManuallyControlled, OgreBone.cpp:82-86): Skeleton::reset leaves them alone
 *     (OgreSkeleton.cpp:233-240), so they start from the local transform they hold — manual_dev[(i * num_manual + k) * 10 ..] =
 *     position xyz, orientation wxyz (as the bone holds it: Node::setOrientation normalises), scale xyz of bone manual_bones[k]
 *     (HOST array of handles) in instance i — and tracks that name them still apply on top, as in the reference.
 * Bit-identical to the engine like the plain call (tests/test_skeleton_gpu.py).  opt == NULL: the plain call. */
typedef struct ogrecu_palette_options {
    const float*    mask_dev;
    size_t          mask_instance_stride_floats;
    uint32_t        num_manual;
    const uint32_t* manual_bones;
    const float*    manual_dev;
} ogrecu_palette_options;
OGRECU_API int ogrecu_skel_palettes_ex_dev(ogrecu_skel* skel, uint32_t num_instances, uint32_t num_states,
                                           const int32_t* anim_dev, const float* time_dev, const float* weight_dev,
                                           const ogrecu_palette_options* opt, float* palettes_dev, size_t palette_stride_floats);

/* -------------------------------------------------------------- OptimisedUtil ------ */
/* Ogre::OptimisedUtil, OgreMain/include/OgreOptimisedUtil.h:99-206; arithmetic replays
 * OgreMain/src/OgreOptimisedUtilSSE.cpp (see DESIGN.md §kernels for the per-function
 * operation order and where the exact General formulation is followed instead). */

/* One skinning job = one softwareVertexSkinning call (OgreOptimisedUtil.h:99-108).
 * All pointers are DEVICE pointers; strides are in bytes.  `palette` is a dense array of
 * `num_matrices` Affine3 (16 floats each, row major, OgreMatrix4.h:80-84) already
 * gathered through the reference's pointer palette (OgreMesh.cpp:1782-1790). */
typedef struct ogrecu_skin_job {
    const float*   src_pos;        /* 3 floats per vertex                                 */
    float*         dst_pos;
    const float*   src_norm;       /* NULL => positions only                              */
    float*         dst_norm;
    const float*   blend_weight;   /* num_weights floats per vertex                       */
    const uint8_t* blend_index;    /* num_weights bytes per vertex                        */
    const float*   palette;        /* Affine3 array, 16 floats each, 16-byte aligned      */
    const uint16_t* bone_map;      /* optional: matrix b = palette[bone_map[b]] (the      */
                                   /* blendIndexToBoneIndexMap, OgreMesh.cpp:1782-1790);  */
                                   /* NULL = identity                                     */
    uint32_t src_pos_stride, dst_pos_stride;
    uint32_t src_norm_stride, dst_norm_stride;
    uint32_t blend_weight_stride, blend_index_stride;
    uint32_t num_weights;          /* 1..4                                                */
    uint32_t num_matrices;         /* matrices addressable by blend_index, <= 256         */
    uint32_t num_vertices;
    uint32_t reserved;
} ogrecu_skin_job;

/* Batched, device-resident skinning: `jobs` is a HOST array of n descriptors (copied to
 * the device by the call); one launch covers all jobs.  Asynchronous on the stream. */
OGRECU_API int ogrecu_skin_dev(ogrecu_ctx* ctx, const ogrecu_skin_job* jobs, uint32_t n);

/* Instanced form of the same job (config 2: 10 000 animated instances of one mesh):
 * instance i reads the shared blend streams, the source streams at
 * src + i*src_instance_stride_bytes (0 = one shared bind pose; non-zero = per-instance
 * morphed source, Entity::updateAnimation OgreEntity.cpp:823-864), palette
 * `palettes + i*palette_stride_floats`, and writes dst + i*dst_instance_stride_bytes. */
OGRECU_API int ogrecu_skin_instanced_dev(ogrecu_ctx* ctx, const ogrecu_skin_job* job,
                                         uint32_t num_instances,
                                         size_t palette_stride_floats,
                                         size_t src_instance_stride_bytes,
                                         size_t dst_instance_stride_bytes);

/* Fused vertex animation of many instances of one mesh — the two steps Entity::updateAnimation
 * runs per entity (OgreMain/src/OgreEntity.cpp:747-911) in ONE pass with no intermediate in HBM:
 *   source_i = softwareVertexMorph(t[i], key1, key2)      (OgreMesh.cpp:1899-1946)
 *   dst_i    = softwareVertexSkinning(source_i, palette_i) (OgreMesh.cpp:1792-1897)
 * `job` describes the skinning exactly as for ogrecu_skin_instanced_dev, with src_pos/src_norm
 * pointing at keyframe 1; `key2` is keyframe 2 in the same dense layout (stride 24 pos+normal or
 * 12 positions only); `t_dev` is a DEVICE array of num_instances morph weights.  When
 * `morph_out` is non-NULL the morphed intermediate of instance i is also written to
 * morph_out + i*morph_instance_stride_bytes (what OGRE keeps in mSoftwareVertexAnimVertexData).
 * Results are bit-identical to ogrecu_morph_instanced_dev followed by ogrecu_skin_instanced_dev. */
OGRECU_API int ogrecu_anim_instanced_dev(ogrecu_ctx* ctx, const ogrecu_skin_job* job, const float* key2,
                                         const float* t_dev, float* morph_out, size_t morph_instance_stride_bytes,
                                         uint32_t num_instances, size_t palette_stride_floats,
                                         size_t dst_instance_stride_bytes);

/* Exact-signature host form (what CudaOptimisedUtil::softwareVertexSkinning forwards to):
 * host pointers, pointer palette of Affine3 (`blend_matrices[i]` -> 16 floats).  Pointers
 * that lie inside a mapped ogrecu_buf shadow are resolved to the device copy; anything
 * else is staged through pinned memory.  Synchronous. */
/* One animation frame of an instanced mesh: the fused morph+skin of EACH of its sub-meshes (what Entity::updateAnimation
 * does sub-entity by sub-entity, OgreEntity.cpp:823-864) in one call.  All jobs share t_dev, num_instances and the
 * palette stride (one skeleton instance per entity).  The independent launches are overlapped internally (largest first
 * over two side streams, forked from and joined into the context's stream), so the call is ordered like one launch. */
typedef struct ogrecu_anim_job {
    ogrecu_skin_job skin;                 /* as for ogrecu_anim_instanced_dev: src_pos/src_norm = keyframe 1 */
    const float* key2;                    /* keyframe 2, same dense layout */
    float* morph_out;                     /* optional: the morphed intermediate per instance, NULL = not materialised */
    size_t morph_instance_stride_bytes;
    size_t dst_instance_stride_bytes;
} ogrecu_anim_job;
OGRECU_API int ogrecu_anim_instanced_multi_dev(ogrecu_ctx* ctx, const ogrecu_anim_job* jobs, uint32_t num_jobs, const float* t_dev,
                                               uint32_t num_instances, size_t palette_stride_floats);

/* Skin plan: what the instance-pair-per-lane kernel (csrc/skin_lane.cu) precomputes ONCE per sub-mesh from its static
 * blend stream — the mesh cut into chunks of <= 128 vertices that reference <= 16 palette entries, and per vertex the
 * slots to take (zero weights and unit first weights are skipped where that is bitwise exact, see the kernel header).
 * `job` supplies blend_index / blend_weight (+ strides), bone_map, num_weights, num_matrices and num_vertices (DEVICE
 * pointers; read back synchronously by the call — the plan does not reference them afterwards).  The blend stream is
 * the VES_BLEND_INDICES / VES_BLEND_WEIGHTS buffer Mesh::softwareVertexBlend locks (OgreMain/src/OgreMesh.cpp:1823-1850)
 * and the map is SubMesh::blendIndexToBoneIndexMap (:1782-1790); rebuild the plan when either changes.
 * OGRECU_ERR_UNSUPPORTED: the mesh has no bone locality (four consecutive vertices reference more than 16 palette entries
 * together with the vertex before them) — pass plans[k] = NULL for it and the tile kernels take the job. */
typedef struct ogrecu_skin_plan ogrecu_skin_plan;
OGRECU_API int ogrecu_skin_plan_create(ogrecu_ctx* ctx, const ogrecu_skin_job* job, ogrecu_skin_plan** out);
OGRECU_API int ogrecu_skin_plan_destroy(ogrecu_skin_plan* plan);
OGRECU_API int ogrecu_skin_plan_info(const ogrecu_skin_plan* plan, uint32_t* num_chunks);
/* One animation frame of an instanced mesh, all sub-meshes in ONE launch (Entity::updateAnimation walks the sub-entities,
 * OgreMain/src/OgreEntity.cpp:823-864): the ogrecu_anim_instanced_multi_dev contract, plus plans[k] for jobs[k].
 * t_dev != NULL: fused morph + skin (every jobs[k].key2 set); t_dev == NULL: skinning of the shared bind pose in
 * skin.src_pos (every key2 NULL).  Jobs whose layout the lane kernel does not take (positions only, a destination that is
 * not 16-byte aligned, morph_out requested, plans[k] NULL) run through the tile kernels of ogrecu_anim_instanced_dev /
 * ogrecu_skin_instanced_dev; results are bit-identical either way. */
OGRECU_API int ogrecu_anim_frame_dev(ogrecu_ctx* ctx, const ogrecu_anim_job* jobs, ogrecu_skin_plan* const* plans, uint32_t num_jobs,
                                     const float* t_dev, uint32_t num_instances, size_t palette_stride_floats);

OGRECU_API int ogrecu_software_vertex_skinning_host(
    ogrecu_ctx* ctx,
    const float* src_pos, float* dst_pos, const float* src_norm, float* dst_norm,
    const float* blend_weight, const uint8_t* blend_index,
    const float* const* blend_matrices,
    size_t src_pos_stride, size_t dst_pos_stride, size_t src_norm_stride, size_t dst_norm_stride,
    size_t blend_weight_stride, size_t blend_index_stride,
    size_t num_weights_per_vertex, size_t num_vertices);

/* Run the OptimisedUtil jobs queued under "optutil_batch" now (no-op when nothing is queued); stats: batched launches so far
 * and jobs they carried (queued ones included). */
OGRECU_API int ogrecu_optutil_flush(ogrecu_ctx* ctx);
OGRECU_API int ogrecu_optutil_batch_stats(ogrecu_ctx* ctx, uint64_t* flushes, uint64_t* jobs);

/* softwareVertexMorph (OgreOptimisedUtil.h:125-131): dst = a + t*(b-a); with
 * morph_normals the interleaved normal is lerped then normalised (General formulation,
 * OgreOptimisedUtilGeneral.cpp:236-283; the SSE one mis-sums the length).
 * Strides in bytes; General accepts any stride >= 12 (24 with normals). */
OGRECU_API int ogrecu_morph_dev(ogrecu_ctx* ctx, float t, const float* src1, const float* src2, float* dst,
                                size_t stride1, size_t stride2, size_t dst_stride,
                                size_t num_vertices, int morph_normals);
/* Instanced morph: instance i writes dst + i*dst_instance_stride_bytes = lerp(src1, src2, t[i])
 * from two shared keyframes; `t_dev` is a DEVICE array of num_instances floats. */
OGRECU_API int ogrecu_morph_instanced_dev(ogrecu_ctx* ctx, const float* t_dev, const float* src1, const float* src2,
                                          float* dst, size_t stride1, size_t stride2, size_t dst_stride,
                                          size_t num_vertices, int morph_normals, uint32_t num_instances,
                                          size_t dst_instance_stride_bytes);
OGRECU_API int ogrecu_software_vertex_morph_host(ogrecu_ctx* ctx, float t, const float* src1, const float* src2,
                                                 float* dst, size_t stride1, size_t stride2, size_t dst_stride,
                                                 size_t num_vertices, int morph_normals);

/* concatenateAffineMatrices (OgreOptimisedUtil.h:142-146): dst[i] = base * src[i];
 * 16 floats per matrix.  `base` is always a HOST pointer (it is a by-reference argument
 * in the reference). */
OGRECU_API int ogrecu_concat_affine_dev(ogrecu_ctx* ctx, const float* base_host16, const float* src, float* dst,
                                        size_t num_matrices);
/* The same for many entities in one launch: what Entity::cacheBoneMatrices (OgreMain/src/OgreEntity.cpp:1307-1330) does per
 * entity with its parent node's full transform.  bases_dev[i] = instance i's base matrix (16 floats, row major, DEVICE);
 * instance i's n matrices are at src + i*instance_stride_floats (dst likewise; dst may alias src).  Bit-identical to n calls
 * of ogrecu_concat_affine_dev. */
OGRECU_API int ogrecu_concat_affine_instanced_dev(ogrecu_ctx* ctx, const float* bases_dev, const float* src, float* dst, uint32_t n,
                                                  uint32_t num_instances, size_t instance_stride_floats);
OGRECU_API int ogrecu_concatenate_affine_matrices_host(ogrecu_ctx* ctx, const float* base16, const float* src,
                                                       float* dst, size_t num_matrices);

/* calculateFaceNormals (OgreOptimisedUtil.h:159-163).  `triangles` has the layout of
 * EdgeData::Triangle (OgreEdgeListBuilder.h:59-72): 32-byte records, vertIndex[3] are the
 * uint32 at byte offsets 8,12,16.  positions are packed xyz (stride 12).  Output: one
 * float4 (nx,ny,nz,-n.v0) per triangle, not normalised. */
OGRECU_API int ogrecu_face_normals_dev(ogrecu_ctx* ctx, const float* positions, const void* triangles,
                                       float* face_normals, size_t num_triangles);
OGRECU_API int ogrecu_calculate_face_normals_host(ogrecu_ctx* ctx, const float* positions, size_t num_vertices,
                                                  const void* triangles, float* face_normals,
                                                  size_t num_triangles);

/* calculateLightFacing (OgreOptimisedUtil.h:181-185): facing[i] = !(dot4(light, n[i]) <= 0)
 * with the SSE summation order (x+z)+(y+w).  light4 is a HOST pointer to 4 floats. */
OGRECU_API int ogrecu_light_facing_dev(ogrecu_ctx* ctx, const float* light4_host, const float* face_normals,
                                       char* facings, size_t num_faces);
/* n_lights lights over the same normals in one launch; facings[l*facing_stride + i] */
OGRECU_API int ogrecu_light_facing_multi_dev(ogrecu_ctx* ctx, const float* lights_host, uint32_t n_lights,
                                             const float* face_normals, char* facings, size_t facing_stride,
                                             size_t num_faces);
OGRECU_API int ogrecu_calculate_light_facing_host(ogrecu_ctx* ctx, const float* light4, const float* face_normals,
                                                  char* facings, size_t num_faces);

/* extrudeVertices (OgreOptimisedUtil.h:201-206): packed xyz in/out.  light4[3]==0 =>
 * directional, else point.  Exact 1/sqrt (General formulation; the SSE rsqrt estimate is
 * CPU-vendor dependent). */
OGRECU_API int ogrecu_extrude_dev(ogrecu_ctx* ctx, const float* light4_host, float extrude_dist,
                                  const float* src, float* dst, size_t num_vertices);
/* the same for 1..8 lights in one launch (lights_host = n_lights x 4 floats): the positions are read once, light k
 * writes dst + k*dst_light_stride_floats; per light bit-identical to ogrecu_extrude_dev */
OGRECU_API int ogrecu_extrude_multi_dev(ogrecu_ctx* ctx, const float* lights_host, uint32_t n_lights, float extrude_dist,
                                        const float* src, float* dst, size_t dst_light_stride_floats, size_t num_vertices);
OGRECU_API int ogrecu_extrude_vertices_host(ogrecu_ctx* ctx, const float* light4, float extrude_dist,
                                            const float* src, float* dst, size_t num_vertices);

/* ---------------------------------------------------------- pose blending ------ */
/* Mesh::softwareVertexPoseBlend (OgreMain/src/OgreMesh.cpp:1948-2007, called per pose by VertexAnimationTrack::applyPoseToVertexData,
 * OgreAnimationTrack.cpp:942) for ONE element — call it with the pose's vertex offsets and the position element, and again
 * with the pose's normals and the normal element: dst[vertex_index[e]] += offsets[e] * weight as `d + (o * w)`, nothing
 * at all when weight == 0.  vertex_index / offsets (3 floats per entry) are the pose's std::map flattened, on the device;
 * dst_dev points at the element's first float of vertex 0.  The instanced form blends the pose into num_instances
 * buffers with a weight per instance (device array). */
OGRECU_API int ogrecu_pose_blend_dev(ogrecu_ctx* ctx, float weight, const uint32_t* vertex_index_dev, const float* offsets_dev,
                                     uint32_t num_entries, float* dst_dev, uint32_t dst_stride_bytes);
OGRECU_API int ogrecu_pose_blend_instanced_dev(ogrecu_ctx* ctx, const float* weights_dev, const uint32_t* vertex_index_dev,
                                               const float* offsets_dev, uint32_t num_entries, float* dst_dev,
                                               uint32_t dst_stride_bytes, uint32_t num_instances,
                                               size_t dst_instance_stride_bytes);

/* The two normal passes around the pose blends of a frame (Entity::_updateAnimation, OgreEntity.cpp:823-864), for
 * num_instances buffers at once: reset = the zeroing of initialisePoseVertexData (OgreEntity.cpp:1180-1204) when the poses
 * carry normals; finalise = Entity::finalisePoseNormals (OgreEntity.cpp:1207-1261): where the accumulated normal is shorter
 * than 1 the base-mesh normal (shared by all instances) fills in with weight 1 - length, then Vector3::normalise.
 * Pointers name the normal of vertex 0. */
OGRECU_API int ogrecu_pose_normals_reset_dev(ogrecu_ctx* ctx, float* dst_norm_dev, uint32_t num_vertices, uint32_t dst_stride_bytes,
                                             uint32_t num_instances, size_t dst_instance_stride_bytes);
OGRECU_API int ogrecu_pose_normals_finalise_dev(ogrecu_ctx* ctx, float* dst_norm_dev, const float* base_norm_dev,
                                                uint32_t num_vertices, uint32_t dst_stride_bytes, uint32_t base_stride_bytes,
                                                uint32_t num_instances, size_t dst_instance_stride_bytes);

/* ------------------------------------------ silhouette / shadow-volume indices ------ */
/* ShadowCaster::generateShadowVolume (OgreMain/src/OgreShadowCaster.cpp:191-515) as an order-preserving stream
 * compaction on the device.  The edge list of a mesh (EdgeData, OgreEdgeListBuilder.h:50-160) is uploaded once:
 *   groups5[g] = {edge_start, edge_count, tri_start, tri_count, vertex_count}   (EdgeGroup; edge ranges consecutive)
 *   edges5[e]  = {triIndex[0], triIndex[1], vertIndex[0], vertIndex[1], degenerate}
 *   tris3[t]   = EdgeData::Triangle::vertIndex */
typedef struct ogrecu_edges ogrecu_edges;
#define OGRECU_SRF_INCLUDE_LIGHT_CAP 1u      /* ShadowRenderableFlags, OgreShadowCaster.h:110-120 */
#define OGRECU_SRF_INCLUDE_DARK_CAP 2u
#define OGRECU_SRF_EXTRUDE_TO_INFINITY 4u
OGRECU_API int ogrecu_edges_create(ogrecu_ctx* ctx, uint32_t num_groups, const uint32_t* groups5, const uint32_t* edges5,
                                   const uint32_t* tris3, uint32_t num_tris, ogrecu_edges** out);
OGRECU_API int ogrecu_edges_destroy(ogrecu_edges* edges);
/* the reference's useMcGuire decision (host arithmetic): bounds6 = light-cap bounds min xyz, max xyz; light4 = Light::getAs4DVector */
OGRECU_API int ogrecu_shadow_use_mcguire(uint32_t num_groups, const float* bounds6, const float* light4);
/* facings_dev = EdgeData::triangleLightFacings on the device (what ogrecu_light_facing_dev wrote).  Writes the indices
 * (index_type 2 = 16-bit as the reference, 4 = 32-bit: lifts the 65 536-vertex limit of OgreShadowCaster.cpp:387) into
 * indices_dev from *used_inout on — or from 0 when they would not fit behind it, as the reference wraps its shared
 * buffer — and updates *used_inout.  ranges4_host[g] = {indexStart, indexCount, lightCapStart, lightCapCount} of the
 * group's ShadowRenderable (the last two only with separate_light_cap and a light cap).  Synchronous, like the
 * reference.  OGRECU_ERR_INVALID when the indices do not fit the buffer at all. */
OGRECU_API int ogrecu_shadow_volume_dev(ogrecu_edges* edges, const char* facings_dev, int directional, int use_mcguire,
                                        uint32_t flags, int separate_light_cap, uint32_t index_type, void* indices_dev,
                                        uint32_t buffer_indices, uint32_t* used_inout, uint32_t* ranges4_host,
                                        uint32_t* emitted);

/* The same for up to 8 lights in ONE submission without a host round trip (BASELINE config 3: face normals -> light facing
 * -> extrusion -> silhouettes of 8 lights back to back on the stream).  Light l reads facings_dev + l * facing_stride_bytes
 * (ogrecu_light_facing_multi_dev's layout), has its own directional / use_mcguire / flags, and writes the region
 * [l * region_indices, (l + 1) * region_indices) of indices_dev; a device pass between the count and scatter kernels does what
 * the host does in the single-light call (CTA bases, group totals, McGuire fan apex, section offsets).  ranges4_host
 * [light][group][4] (index values are absolute positions in indices_dev) and result2_host [light][2] = {indices emitted — or
 * needed, when status is 1 —, status: 0 ok, 1 did not fit the region: nothing written} are copied asynchronously: valid after
 * ogrecu_sync (pass pinned memory, ogrecu_host_alloc, for the copies to be asynchronous).  Index bytes per light equal the
 * single-light call's. */
OGRECU_API int ogrecu_shadow_volumes_multi_dev(ogrecu_edges* edges, uint32_t num_lights, const char* facings_dev, size_t facing_stride_bytes,
                                               const int* directional, const int* use_mcguire, const uint32_t* flags,
                                               int separate_light_cap, uint32_t index_type, void* indices_dev, uint32_t region_indices,
                                               uint32_t* ranges4_host, uint32_t* result2_host);

/* ------------------------------------------------------------------ rasteriser ----- */
/* TinyTexture / TinyHardwarePixelBuffer (RenderSystems/Tiny/src/OgreTinyTexture.cpp:30-57):
 * always RGBA8, tightly packed, 0 mips. */
OGRECU_API int ogrecu_tex_create(ogrecu_ctx* ctx, uint32_t width, uint32_t height, ogrecu_tex** out);
OGRECU_API int ogrecu_tex_destroy(ogrecu_tex* tex);
OGRECU_API int ogrecu_tex_upload(ogrecu_tex* tex, const uint8_t* rgba_host, size_t row_pitch_bytes);
OGRECU_API int ogrecu_tex_download(ogrecu_tex* tex, uint8_t* rgba_host, size_t row_pitch_bytes);
OGRECU_API int ogrecu_tex_device_ptr(ogrecu_tex* tex, void** dev_ptr);
/* Render-to-texture resolve: copy a target's RGB8 colour plane into a texture of the same size as
 * RGBA8 with alpha 255, on the device (what sampling a TU_RENDERTARGET texture needs; Tiny has no
 * render textures).  Flushes the target first.  Ordered on the stream. */
OGRECU_API int ogrecu_tex_copy_from_rt(ogrecu_tex* tex, ogrecu_rt* rt);

/* Render target = TinyWindow colour Image (PF_BYTE_RGB, tightly packed,
 * OgreTinyWindow.cpp:38-42) + TinyDepthBuffer Image (PF_FLOAT32_R, OgreTinyDepthBuffer.h:20). */
OGRECU_API int ogrecu_rt_create(ogrecu_ctx* ctx, uint32_t width, uint32_t height, ogrecu_rt** out);
/* ... on device `placement` of the context's device group (0 = the root = what ogrecu_rt_create does), or split over all of its
 * devices (see ogrecu_init_devices).  On a single-device context OGRECU_PLACE_SPLIT and 0 both give an ordinary target. */
#define OGRECU_PLACE_SPLIT (-1)
OGRECU_API int ogrecu_rt_create_placed(ogrecu_ctx* ctx, uint32_t width, uint32_t height, int placement, ogrecu_rt** out);
OGRECU_API int ogrecu_rt_destroy(ogrecu_rt* rt);
OGRECU_API int ogrecu_rt_size(const ogrecu_rt* rt, uint32_t* width, uint32_t* height);

enum { OGRECU_CLEAR_COLOUR = 1, OGRECU_CLEAR_DEPTH = 2 };
/* clearFrameBuffer (OgreTinyRenderSystem.cpp:421-433): whole target, ignores viewport;
 * colour packed as (uint)(v*256) clamped to 255 (PixelUtil::packColour / Bitwise::floatToFixed,
 * OgreBitwise.h:237-242); depth plane filled with `depth`.  Ordered with draws. */
OGRECU_API int ogrecu_rt_clear(ogrecu_rt* rt, unsigned buffers, const float colour_rgba[4], float depth);

enum { OGRECU_PRIM_TRIANGLE_LIST = 4, OGRECU_PRIM_TRIANGLE_STRIP = 5 };  /* RenderOperation::OperationType values */
enum { OGRECU_INDEX_NONE = 0, OGRECU_INDEX_U16 = 2, OGRECU_INDEX_U32 = 4 };
enum {
    OGRECU_DRAW_DEPTH_TEST  = 1,   /* _setDepthBufferParams depthTest  (.cpp:270-274) */
    OGRECU_DRAW_DEPTH_WRITE = 2,   /* _setDepthBufferParams depthWrite               */
    OGRECU_DRAW_BLEND_ADD   = 4,   /* setColourBlendState destFactor==SBF_ONE (.cpp:284-287) */
    OGRECU_DRAW_LIGHTING    = 8    /* setLightingEnabled && mesh has normals (.cpp:369) */
};

/* One TinyRenderSystem::_render(RenderOperation) (OgreTinyRenderSystem.cpp:349-414) with
 * the render state and DefaultShader uniforms (include/OgreTinyRenderSystem.h:57-75) that
 * were current when it was issued.  Vertex/index pointers are DEVICE pointers already
 * offset the way getData() does (`base + elemOffset + vertexStart*stride`, .cpp:335-347;
 * indices + indexStart).  Matrices are row-major float[16] (Ogre::Matrix4). */
typedef struct ogrecu_draw_desc {
    const void* positions;  uint32_t pos_stride;     /* float3 per vertex                 */
    const void* uvs;        uint32_t uv_stride;      /* float2 or NULL                    */
    const void* normals;    uint32_t normal_stride;  /* float3 or NULL                    */
    const void* indices;    uint32_t index_type;     /* OGRECU_INDEX_*                    */
    uint32_t draw_count;        /* indexCount, or vertexCount when not indexed (.cpp:373,387) */
    uint32_t prim_type;         /* OGRECU_PRIM_*; lists are culled (cross>0), strips not (.cpp:410) */
    uint32_t flags;             /* OGRECU_DRAW_*                                          */
    uint32_t vertex_count;      /* vertexData->vertexCount: indices are < vertex_count.  Optional (0 = unknown):
                                 * when given, every vertex is transformed ONCE per flush instead of once per
                                 * index as Tiny does (.cpp:400-408) — same operations per vertex, same bits */
    const ogrecu_tex* texture;  /* unit 0 or NULL (white)  (.cpp:208-220)                 */
    float mvp[16];              /* uniform_MVP                                            */
    float tex_matrix[16];       /* uniform_Tex                                            */
    float mvit[16];             /* uniform_MVIT (upper-left 3x3 used)                     */
    float viewport[16];         /* mVP from _setViewport (.cpp:230-262)                   */
    float light_dir[3];         /* uniform_lightDir                                       */
    float ambient[4];           /* uniform_ambientCol                                     */
    /* Instancing (triangle lists only; 0 = a plain draw).  RenderOperation::numberOfInstances
     * (OgreMain/include/OgreRenderOperation.h:133) is ignored by Tiny (.cpp:349-414); here instance i draws the same
     * draw_count indices with i * instance_vertex_count added to every vertex index (non-indexed: to the running vertex
     * number) — the vertex streams hold instance_count copies of the mesh back to back, as ogrecu_anim_frame_dev writes
     * them.  Equal to instance_count separate draws in instance order.  vertex_count then counts ALL instances. */
    uint32_t instance_count;
    uint32_t instance_vertex_count;
} ogrecu_draw_desc;

/* Record a draw on `rt` (cheap; nothing runs yet).  Draws and clears keep submission order. */
OGRECU_API int ogrecu_draw(ogrecu_rt* rt, const ogrecu_draw_desc* desc);
/* ---- "Beyond Tiny" render state (opt-in) ----
 * State TinyRenderSystem's setters ignore (OgreTinyRenderSystem.cpp:208-297: depth function, culling mode, scissor,
 * alpha rejection, vertex colours).  A field is honoured only when its OGRECU_EXT_* bit is set; with enable == 0 (or through
 * plain ogrecu_draw) the draw is Tiny's, bit for bit.  Definitions (the reference has no behaviour of its own here):
 *   depth_func   replaces Tiny's fixed `frag_depth > zbuf` rejection (tinyrenderer.h:79) while OGRECU_DRAW_DEPTH_TEST is set;
 *                CMPF_LESS_EQUAL is that rule (kept in Tiny's form, under which a NaN depth passes);
 *   cull_mode    CULL_CLOCKWISE = Tiny's rule, `cross(p2-p0, p2-p1) > 0` is culled (tinyrenderer.h:37-39); CULL_ANTICLOCKWISE =
 *                `< 0`; CULL_NONE.  Applies to strips too, whose odd triangles count with the opposite sign (Tiny never culls
 *                strips, OgreTinyRenderSystem.cpp:410);
 *   scissor      fragments outside [left,right) x [top,bottom) are dropped; clears still cover the whole target (as Tiny's);
 *   alpha        the fragment survives iff cmp(alpha, ref) on the texel's alpha byte (255 without a texture), after Tiny's own
 *                `alpha == 0` discard (OgreTinyRenderSystem.cpp:318);
 *   colours      c_j = byte * (1/255) per vertex, interpolated as uv is ((c0*b0 + c1*b1) + c2*b2) and multiplied into rgb
 *                after the texture and before lighting.
 * A depth function other than CMPF_LESS_EQUAL or an alpha test other than CMPF_ALWAYS_PASS sends the tiles the draw touches
 * down the in-order raster path. */
enum {   /* Ogre::CompareFunction, OgreMain/include/OgreCommon.h:61-71 */
    OGRECU_CMPF_ALWAYS_FAIL = 0, OGRECU_CMPF_ALWAYS_PASS = 1, OGRECU_CMPF_LESS = 2, OGRECU_CMPF_LESS_EQUAL = 3, OGRECU_CMPF_EQUAL = 4,
    OGRECU_CMPF_NOT_EQUAL = 5, OGRECU_CMPF_GREATER_EQUAL = 6, OGRECU_CMPF_GREATER = 7
};
enum { OGRECU_CULL_NONE = 1, OGRECU_CULL_CLOCKWISE = 2, OGRECU_CULL_ANTICLOCKWISE = 3 };   /* Ogre::CullingMode, OgreCommon.h:164-172 */
enum {
    OGRECU_EXT_DEPTH_FUNC = 1,      /* RenderSystem::_setDepthBufferParams depthFunction (OgreRenderSystem.h:709)     */
    OGRECU_EXT_CULL_MODE = 2,       /* RenderSystem::_setCullingMode (:692); CULL_CLOCKWISE is Tiny's built-in rule   */
    OGRECU_EXT_SCISSOR = 4,         /* RenderSystem::setScissorTest (:916)                                            */
    OGRECU_EXT_ALPHA_REJECT = 8,    /* RenderSystem::_setAlphaRejectSettings (:643), on the texel's alpha byte        */
    OGRECU_EXT_VERTEX_COLOUR = 16   /* VES_DIFFUSE stream, VET_UBYTE4_NORM (bytes r,g,b,a), modulates rgb             */
};
typedef struct ogrecu_draw_state {
    uint32_t enable;            /* OGRECU_EXT_* bits                                  */
    uint32_t depth_func;        /* OGRECU_CMPF_*                                      */
    uint32_t cull_mode;         /* OGRECU_CULL_*                                      */
    uint32_t alpha_func;        /* OGRECU_CMPF_*                                      */
    uint32_t alpha_ref;         /* 0..255                                             */
    uint32_t scissor[4];        /* left, top, right, bottom (exclusive), pixels       */
    uint32_t colour_stride;     /* bytes between vertices of `colours`                */
    const void* colours;        /* DEVICE pointer, offset like the other streams      */
} ogrecu_draw_state;
/* ogrecu_draw with the extra state; state == NULL is ogrecu_draw. */
OGRECU_API int ogrecu_draw_with_state(ogrecu_rt* rt, const ogrecu_draw_desc* desc, const ogrecu_draw_state* state);
/* The same for a caller that knows which ogrecu_bufs the draw's pointers lie in (the plugin: every stream of a RenderOperation is a
 * CudaHardwareBuffer): refs[0..num_refs) — NULL entries allowed — are marked as read by the recorded draw directly, instead of
 * being looked up from the five device pointers in the context's registry (five ordered-map look-ups under a lock per _render:
 * 0.3 us of the 1.5 us a _render cost on the host).  A pointer that lies in a buffer NOT named here is not tracked: name them all. */
OGRECU_API int ogrecu_draw_refs(ogrecu_rt* rt, const ogrecu_draw_desc* desc, const ogrecu_draw_state* state,
                                const ogrecu_buf* const* refs, uint32_t num_refs);
/* Run everything recorded on `rt`: vertex transform + triangle setup, order-preserving
 * tile binning, per-tile raster/shade/blend, coalesced flush to the colour/depth planes.
 * Asynchronous on the stream (_endFrame / readback / buffer lock are the sync points). */
OGRECU_API int ogrecu_flush(ogrecu_rt* rt);

/* Batched offscreen views (config 4): the same draw list (everything but mvp/mvit/light)
 * rendered into n_views targets of identical size in one pipeline pass — n_views x TinyRenderSystem::_render
 * (OgreTinyRenderSystem.cpp:349-414) with the per-camera uniforms of applyFixedFunctionParams (:49-82).
 * per_view_mvp / per_view_mvit: n_views*16 floats (HOST); per_view_light_dir: n_views*3. */
OGRECU_API int ogrecu_draw_views(ogrecu_rt* const* rts, uint32_t n_views, const ogrecu_draw_desc* desc,
                                 const float* per_view_mvp, const float* per_view_mvit,
                                 const float* per_view_light_dir);

/* Split-frame (config 5; no counterpart in the reference, whose triangle() walks the whole bounding box,
 * tinyrenderer.h:67-69): restrict rasterisation on this context to screen tile rows
 * [y0, y1) (multiples of the tile height); planes outside are left untouched.  The owner
 * of a strip is the only writer, so the composite is a disjoint gather. */
OGRECU_API int ogrecu_rt_set_strip(ogrecu_rt* rt, uint32_t y0, uint32_t y1);
/* The load-balanced form: the frame is cut into bands of `band_rows` pixel rows (a multiple of the tile height) and this
 * context rasterises the bands b with b % period == phase (rank and world size of the split).  Dense and empty
 * regions of the screen then spread over all GPUs.  ogrecu_rt_set_strip(rt, 0, height) returns to the whole frame. */
OGRECU_API int ogrecu_rt_set_bands(ogrecu_rt* rt, uint32_t band_rows, uint32_t phase, uint32_t period);

/* Sort-last split of one frame (config 5, the scalable form): rank r draws the r-th slice of the TRIANGLES (in submission
 * order) into its own full-size target, and the targets are merged per pixel: smallest depth wins, the later rank on ties,
 * pixels a rank did not touch (still the cleared colour AND depth) never win (the per-pixel outcome of Tiny's in-order
 * `frag_depth > zbuf` rule, tinyrenderer.h:79, when the submission order is the rank order).  That is Tiny's result whenever every
 * triangle is opaque, depth-tested and depth-writing (ogrecu_raster_stats.tiles_ordered == 0 on every rank); blended or
 * depth-write-off content must use the screen-space split (ogrecu_rt_set_strip / _set_bands).  The call merges
 * num_sources source planes (RGB8 + f32 depth, num_pixels each, source s at src + s*src_stride_pixels pixels), which must
 * be those of the ranks AFTER the one `dst` holds, in rank order, into `dst`. */
OGRECU_API int ogrecu_composite_depth_dev(ogrecu_ctx* ctx, uint8_t* dst_rgb, float* dst_depth, const uint8_t* src_rgb,
                                          const float* src_depth, size_t num_pixels, uint32_t num_sources,
                                          size_t src_stride_pixels, const float clear_rgba[4], float clear_depth);

/* Readback (TinyWindow::copyContentsToMemory, OgreTinyWindow.cpp:68-80;
 * HardwarePixelBuffer::blitToMemory).  Flushes pending work, copies the box
 * [left,right) x [top,bottom) to host as tightly described by dst_pitch_bytes.
 * Colour is RGB8 (3 bytes/pixel); depth is float32. */
OGRECU_API int ogrecu_rt_read_colour(ogrecu_rt* rt, uint32_t left, uint32_t top, uint32_t right, uint32_t bottom,
                                     uint8_t* dst_host, size_t dst_pitch_bytes);
/* The same readback converted on the device to a 32-bit, 8-bits-per-channel destination format: every pixel becomes the
 * native-endian word r << r_shift | g << g_shift | b << b_shift | fill (fill = 0xFF in the alpha or X byte) — the bytes PixelUtil::bulkPixelConversion produces from PF_BYTE_RGB on the CPU in the reference
 * (PF_A8R8G8B8: shifts 16, 8, 0, fill 0xFF000000). */
OGRECU_API int ogrecu_rt_read_colour_packed32(ogrecu_rt* rt, uint32_t left, uint32_t top, uint32_t right, uint32_t bottom,
                                              void* dst_host, size_t dst_pitch_bytes, uint32_t r_shift, uint32_t g_shift,
                                              uint32_t b_shift, uint32_t fill);
/* Scaled read-back: the sub-box resampled to dst_width x dst_height with Image::scale(FILTER_BILINEAR)'s byte arithmetic
 * (LinearResampler_Byte, OgreMain/src/OgreImageResampler.h:283-339) on the device — what TinyHardwarePixelBuffer::blitToMemory
 * does on the CPU when the sizes differ (RenderSystems/Tiny/src/OgreTinyHardwarePixelBuffer.cpp:43-47, the read-back of a
 * RenderTexture, OgreMain/src/OgreRenderTexture.cpp:50-56).  dst_px_bytes = 3: RGB8 rows; 4: packed words as
 * ogrecu_rt_read_colour_packed32 (Image::scale's temp buffer + bulkPixelConversion tail, OgreImage.cpp:453-457). */
OGRECU_API int ogrecu_rt_read_colour_scaled(ogrecu_rt* rt, uint32_t left, uint32_t top, uint32_t right, uint32_t bottom, uint32_t dst_width,
                                            uint32_t dst_height, void* dst_host, size_t dst_pitch_bytes, uint32_t dst_px_bytes,
                                            uint32_t r_shift, uint32_t g_shift, uint32_t b_shift, uint32_t fill);
OGRECU_API int ogrecu_rt_read_depth(ogrecu_rt* rt, uint32_t left, uint32_t top, uint32_t right, uint32_t bottom,
                                    float* dst_host, size_t dst_pitch_bytes);
/* upload into the planes (render-to-texture sources, tests) */
OGRECU_API int ogrecu_rt_write_colour(ogrecu_rt* rt, const uint8_t* src_host, size_t src_pitch_bytes);
OGRECU_API int ogrecu_rt_write_depth(ogrecu_rt* rt, const float* src_host, size_t src_pitch_bytes);
OGRECU_API int ogrecu_rt_colour_device_ptr(ogrecu_rt* rt, void** dev_ptr);
OGRECU_API int ogrecu_rt_depth_device_ptr(ogrecu_rt* rt, void** dev_ptr);

/* statistics of the last flush on this target (triangles submitted / surviving cull+clip,
 * (triangle,tile) pairs binned, tiles rasterised) */
typedef struct ogrecu_raster_stats {
    uint64_t triangles_in, triangles_binned, tile_pairs, tiles_active;
    uint64_t tiles_visibility;   /* tiles resolved by the order-independent visibility path            */
    uint64_t tiles_ordered;      /* tiles replayed in submission order (state or NaN/depth fallback)   */
} ogrecu_raster_stats;
OGRECU_API int ogrecu_rt_last_stats(ogrecu_rt* rt, ogrecu_raster_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* OGRECU_H */
