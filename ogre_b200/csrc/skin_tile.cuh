// skin_tile.cuh — parameters of the tiled skinning stream kernel (skin_tile.cu).
#pragma once
#include "common.cuh"

namespace ogrecu {

struct SkinTileParams {
    const uint8_t* src;          // !FUSED: instance-0 source (pos[+normal] dense, VS bytes/vertex); FUSED: keyframe 1 (shared)
    const uint8_t* key2;         // FUSED: keyframe 2 (shared)
    const float* tarr;           // FUSED: morph weight per instance (device)
    uint8_t* morph_out;          // FUSED, optional: the morphed intermediate per instance (NULL = not materialised)
    uint8_t* dst;                // instance-0 destination, dense VS bytes/vertex
    const float* blend_weight;
    const uint8_t* blend_index;
    const float* palette;        // instance-0 palette (Affine3, 16 floats each)
    const uint16_t* bone_map;    // optional
    size_t palette_stride_floats, src_inst_stride, dst_inst_stride, morph_inst_stride;
    uint32_t blend_weight_stride, blend_index_stride;
    uint32_t num_matrices, num_vertices, num_instances, inst_per_cta;
    uint32_t tma_load, tma_store;   // host verified 16-byte alignment of bases and instance strides
    uint32_t load_unit, store_unit; // cooperative fallback granularity: 8 or 4 bytes
};

// Fills the derived fields (TMA eligibility, cooperative unit, instances per CTA) and launches.
int launch_skin_tile(ogrecu_ctx* ctx, SkinTileParams P, uint32_t num_weights, bool with_normals, bool fused_morph);

// skin_pair.cu: same parameters (alignment fields already derived), two instances per pass.
int launch_skin_pair(ogrecu_ctx* ctx, SkinTileParams P, uint32_t num_weights, bool with_normals, bool fused_morph);

// skin_stream.cu: one instance, one long stream (persistent CTAs, resident palette, TMA pipeline over tiles).  Covers the full
// 256-vertex tiles and reports how many vertices that was; the caller skins the tail with the tile kernel.
bool skin_stream_eligible(const SkinTileParams& P, uint32_t num_weights, bool fused_morph);
int launch_skin_stream(ogrecu_ctx* ctx, const SkinTileParams& P, uint32_t num_weights, bool with_normals, uint32_t* vertices_done);

}  // namespace ogrecu
