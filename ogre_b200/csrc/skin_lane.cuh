// skin_lane.cuh — parameters and plan of the instance-pair-per-lane skinning kernel (skin_lane.cu).
#pragma once
#include <cuda.h>
#include "common.cuh"

#ifndef LN_MAXB_N
#define LN_MAXB_N 16
#endif

namespace ogrecu {

#ifndef LN_WARPS_N
#define LN_WARPS_N 15
#endif
#ifndef LN_NBUF_N
#define LN_NBUF_N 2
#endif
constexpr int LN_WARPS = LN_WARPS_N;          // consumer warps; one more warp produces their inputs (16 warps: 128 registers each)
#ifndef LN_UNIT_BATCH_N
#define LN_UNIT_BATCH_N 2
#endif
constexpr int LN_UNIT_BATCH = LN_UNIT_BATCH_N;   // consecutive work units a CTA takes from the global counter at a time
constexpr int LN_NBUF = LN_NBUF_N;            // unit buffers: the producer runs this many units ahead of the slowest consumer warp
constexpr int LN_THREADS = (LN_WARPS + 1) * 32;
constexpr int LN_MAXB = LN_MAXB_N;            // palette entries resident per chunk (a chunk whose first quad plus the vertex before it needs more cannot be planned)
#ifndef LN_CHUNK_QUADS_N
#define LN_CHUNK_QUADS_N 30
#endif
constexpr int LN_CHUNK_QUADS = LN_CHUNK_QUADS_N;            // vertex quads per chunk at most (one vertex per thread in the prologue)
constexpr int LN_CHUNK_VERTS = 4 * LN_CHUNK_QUADS;
constexpr int LN_SLOT_BYTES = 3072;           // 64 instances x 48 bytes
constexpr int LN_BOX_BYTES = 32 * 96;         // one TMA box: 32 instance rows x 4 vertices
constexpr int LN_REC_BYTES = 32;              // plan record per vertex: 4 slot offsets (+ flags in the low bits of the first) and 4 weights
constexpr int LN_MAX_JOBS = 8;
static_assert(LN_MAXB >= 4 && LN_MAXB <= 16 && LN_MAXB * LN_SLOT_BYTES < 65536 * 16, "palette window");

struct LaneJob {
    CUtensorMap mapA, mapB;          // destination as [instance pair][words]: even instances / odd instances
    const uint8_t* key1;             // keyframe 1 or the shared bind pose, pos+normal dense (24 B per vertex)
    const uint8_t* key2;             // keyframe 2 (MORPH)
    const uint4* recs;               // plan records [num_vertices][2]
    const uint4* pre_recs;           // [num_chunks][2]: the vertex before each chunk, in that chunk's slot numbering
    uint8_t* dst;                    // instance 0 destination
    size_t dst_stride;               // bytes between instances
    const uint4* chunks;             // plan chunks [num_chunks][3]: {v_begin, nquads, nslots, -}, then the palette matrix (after bone_map) of each of 16 slots
    uint32_t nw, nverts, nchunks, shift, first_item, pad;   // shift: odd instances start 8 mod 16 (stride = 8 mod 16)
};

struct LaneParams {
    LaneJob job[LN_MAX_JOBS];
    const float* palette;            // instance-0 palette (Affine3, 16 floats each)
    const float* t;                  // morph weight per instance (MORPH)
    size_t pal_stride_floats;
    uint32_t* counter;               // work-unit counter (zeroed before the launch)
    const uint8_t* ipal;             // repacked palettes [group][matrix][3072 B]
    const uint32_t* iflags;          // hazard flags [group][matrix]
    uint32_t nmat, pad;
    uint32_t njobs, ninst, groups, total_units;
};

}  // namespace ogrecu

struct ogrecu_skin_plan {
    ogrecu_ctx* ctx;
    // what the plan was built from (a job must match to use it)
    uint32_t num_vertices, num_weights, num_matrices;
    uint32_t num_chunks, max_matrix;
    void* d_recs;
    void* d_pre_recs;
    void* d_chunks;
};

namespace ogrecu {

int build_skin_plan(ogrecu_ctx* ctx, const uint8_t* blend_index, size_t index_stride, const uint8_t* blend_weight, size_t weight_stride,
                    const uint16_t* bone_map, uint32_t nw, uint32_t num_matrices, uint32_t nverts, ogrecu_skin_plan* plan);
bool skin_lane_eligible(const ogrecu_skin_plan* plan, const ogrecu_anim_job& J, uint32_t num_instances, size_t palette_stride_floats);
// jobs all eligible; morph = keyframe pairs + t_dev, else the shared bind pose of skin.src_pos
// the frame's palettes -> pair layout + hazard flags in ctx->lane_ipal (matrices 0 .. nmat-1 of every instance)
int repack_palettes(ogrecu_ctx* ctx, const float* palette, size_t palette_stride_floats, uint32_t num_instances, uint32_t nmat);
int launch_skin_lane(ogrecu_ctx* ctx, const ogrecu_anim_job* jobs, ogrecu_skin_plan* const* plans, uint32_t num_jobs, const float* t_dev,
                     uint32_t num_instances, size_t palette_stride_floats, bool morph, uint32_t nmat);

}  // namespace ogrecu
