// shadow.cu — silhouette extraction / shadow-volume index generation as an order-preserving stream compaction
// (SURVEY §8f rank 2).
//
// Reference work replaced: ShadowCaster::generateShadowVolume, OgreMain/src/OgreShadowCaster.cpp:191-515 — the serial
// CPU loop that follows calculateLightFacing / extrudeVertices in the stencil-shadow chain: per edge group it emits
// the side quads of every silhouette edge (two triangles' light facings differ, or a degenerate edge of a lit
// triangle), then the dark cap, then the light cap, as 16-bit indices into a shared index buffer.  The reference is
// limited to 65 536 vertices per group (:387); 32-bit output here lifts that.
//
// Shape: the output position of an item depends only on how many flagged items precede it in its group, so
//   pass 1  one CTA per 256 items (edges of a group, or triangles of a group): count the flagged ones;
//   host    the per-CTA counts (a few KB) give every CTA its base, every group its ranges, and the wrap decision of
//           the shared buffer (the call is synchronous, as the reference's is);
//   pass 2  the same CTAs flag again, rank inside the CTA (ballot + shared scan) and scatter the indices.
// Integer work only; results are identical to the reference's bytes (tests/test_shadow_gpu.py).
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

using namespace ogrecu;

namespace {

constexpr int SH_THREADS = 256;

struct BlockDesc { uint32_t group, kind, start, count; };      // kind 0: edges [start, start+count), 1: triangles

struct GroupLayout {                                            // filled by the host between the passes
    uint32_t edge_base, dark_base, light_base, vertex_count;    // index offsets of the three sections, original vertex count
    uint32_t dark_start;                                        // McGuire fan apex (index value)
    uint32_t pad[3];
};

struct ShadowParams {
    const BlockDesc* blocks;
    const uint32_t* edges;          // [e][5] tri0, tri1, v0, v1, degenerate
    const uint32_t* tris;           // [t][3]
    const int8_t* facings;          // [t]
    uint32_t* block_count;          // [nblocks] flagged items (pass 1)
    uint32_t* block_first;          // [nblocks] first flagged item of the CTA (pass 1; 0xffffffff = none)
    const uint32_t* block_base;     // [nblocks] flagged items of the same group and kind before this CTA (pass 2)
    const GroupLayout* layout;      // [groups] (pass 2)
    void* indices;
    uint32_t two_tris, fan, dark_cap, light_cap;
};

__device__ __forceinline__ bool silhouette(const uint32_t* e, const int8_t* __restrict__ facings)
{
    const int lf = facings[e[0]];
    return e[4] ? lf != 0 : lf != facings[e[1]];
}

__device__ __forceinline__ bool flagged(const ShadowParams& P, const BlockDesc& D, uint32_t i)
{
    if (i >= D.count) return false;
    if (D.kind == 0) return silhouette(P.edges + 5 * (size_t)(D.start + i), P.facings);
    return P.facings[D.start + i] != 0;
}

__global__ void __launch_bounds__(SH_THREADS) shadow_count_kernel(ShadowParams P)
{
    __shared__ uint32_t s_cnt[SH_THREADS / 32], s_first[SH_THREADS / 32];
    const BlockDesc D = P.blocks[blockIdx.x];
    const bool f = flagged(P, D, threadIdx.x);
    const uint32_t b = __ballot_sync(0xffffffffu, f);
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_cnt[wid] = __popc(b); s_first[wid] = b ? wid * 32 + (__ffs(b) - 1) : 0xffffffffu; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t c = 0, first = 0xffffffffu;
        for (int w = 0; w < SH_THREADS / 32; ++w) { c += s_cnt[w]; first = min(first, s_first[w]); }
        P.block_count[blockIdx.x] = c;
        P.block_first[blockIdx.x] = first;
    }
}

template <typename IndexT>
__global__ void __launch_bounds__(SH_THREADS) shadow_scatter_kernel(ShadowParams P)
{
    __shared__ uint32_t s_warp[SH_THREADS / 32];
    const BlockDesc D = P.blocks[blockIdx.x];
    const GroupLayout G = P.layout[D.group];
    const bool f = flagged(P, D, threadIdx.x);
    const uint32_t b = __ballot_sync(0xffffffffu, f);
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_warp[wid] = __popc(b);
    __syncthreads();
    uint32_t rank = P.block_base[blockIdx.x] + __popc(b & ((1u << lane) - 1u));
    for (unsigned w = 0; w < wid; ++w) rank += s_warp[w];
    if (!f) return;
    IndexT* out = static_cast<IndexT*>(P.indices);
    const uint32_t ovc = G.vertex_count;
    if (D.kind == 0) {
        const uint32_t* e = P.edges + 5 * (size_t)(D.start + threadIdx.x);
        uint32_t v0 = e[2], v1 = e[3];
        if (!P.facings[e[0]]) { const uint32_t t = v0; v0 = v1; v1 = t; }     // run anticlockwise seen from the lit triangle
        const uint32_t per = 3u + 3u * P.two_tris + 3u * P.fan;
        IndexT* p = out + G.edge_base + rank * per - (rank ? 3u * P.fan : 0u);     // the first silhouette edge only sets the fan apex
        *p++ = (IndexT)v1; *p++ = (IndexT)v0; *p++ = (IndexT)(v0 + ovc);
        if (P.two_tris) { *p++ = (IndexT)(v0 + ovc); *p++ = (IndexT)(v1 + ovc); *p++ = (IndexT)v1; }
        if (P.fan && rank) { *p++ = (IndexT)G.dark_start; *p++ = (IndexT)(v1 + ovc); *p++ = (IndexT)(v0 + ovc); }
    } else {
        const uint32_t* T = P.tris + 3 * (size_t)(D.start + threadIdx.x);
        if (P.dark_cap) {
            IndexT* p = out + G.dark_base + 3 * rank;
            p[0] = (IndexT)(T[1] + ovc); p[1] = (IndexT)(T[0] + ovc); p[2] = (IndexT)(T[2] + ovc);
        }
        if (P.light_cap) {
            IndexT* p = out + G.light_base + 3 * rank;
            p[0] = (IndexT)T[0]; p[1] = (IndexT)T[1]; p[2] = (IndexT)T[2];
        }
    }
}


/* ---- several lights per call, no host round trip ------------------------------------------------------------------------
 * The same two passes with the light as blockIdx.y, and the host's work between them — CTA bases inside each (group, kind)
 * run, per-group totals, the McGuire fan apex, the section offsets and the ShadowRenderable ranges — done by ONE CTA per
 * light on the device (shadow_layout_kernel).  Every light writes its own region of the index buffer. */
constexpr int SH_MAX_LIGHTS = 8;
struct LightCfg { uint32_t two_tris, fan, dark_cap, light_cap; };
struct GroupStatic { uint32_t edge_blk0, edge_blk1, tri_blk0, tri_blk1, vertex_count; };     // block ranges of the group's edges / triangles
struct ShadowMulti {
    const BlockDesc* blocks;
    const uint32_t* edges;
    const uint32_t* tris;
    const int8_t* facings;          // light l: facings + l * facing_stride
    size_t facing_stride;
    uint32_t* block_count;          // [L][nblocks]
    uint32_t* block_first;          // [L][nblocks]
    uint32_t* block_base;           // [L][nblocks]
    GroupLayout* layout;            // [L][groups]
    const GroupStatic* gstatic;     // [groups]
    const uint32_t* seg_first;      // [nblocks] first block of the block's (group, kind) run
    uint32_t* ranges;               // [L][groups][4]
    uint32_t* result;               // [L][2]: indices emitted (or needed), status 0 ok / 1 did not fit
    void* indices;
    uint32_t nblocks, ngroups, region_indices, separate_light_cap;
    LightCfg cfg[SH_MAX_LIGHTS];
};

// the item's table entry is read ONCE and tested against every light's facings (light loop inside the CTA)
struct ItemData { uint32_t a, b, c, d, e; bool valid; };
__device__ __forceinline__ ItemData load_item(const ShadowMulti& M, const BlockDesc& D, uint32_t i)
{
    ItemData it{0, 0, 0, 0, 0, i < D.count};
    if (!it.valid) return it;
    if (D.kind == 0) { const uint32_t* e = M.edges + 5 * (size_t)(D.start + i); it.a = e[0]; it.b = e[1]; it.c = e[2]; it.d = e[3]; it.e = e[4]; }
    else it.a = D.start + i;
    return it;
}
__device__ __forceinline__ bool flagged_item(const ItemData& it, uint32_t kind, const int8_t* __restrict__ facings)
{
    if (!it.valid) return false;
    const int lf = facings[it.a];
    if (kind != 0) return lf != 0;
    return it.e ? lf != 0 : lf != facings[it.b];
}

__global__ void __launch_bounds__(SH_THREADS) shadow_count_multi_kernel(ShadowMulti M, uint32_t num_lights)
{
    __shared__ uint32_t s_cnt[SH_MAX_LIGHTS][SH_THREADS / 32], s_first[SH_MAX_LIGHTS][SH_THREADS / 32];
    const BlockDesc D = M.blocks[blockIdx.x];
    const ItemData it = load_item(M, D, threadIdx.x);
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t l = 0; l < num_lights; ++l) {
        const bool f = flagged_item(it, D.kind, M.facings + l * M.facing_stride);
        const uint32_t b = __ballot_sync(0xffffffffu, f);
        if (lane == 0) { s_cnt[l][wid] = __popc(b); s_first[l][wid] = b ? wid * 32 + (__ffs(b) - 1) : 0xffffffffu; }
    }
    __syncthreads();
    if (threadIdx.x < num_lights) {
        const uint32_t l = threadIdx.x;
        uint32_t c = 0, first = 0xffffffffu;
        for (int w = 0; w < SH_THREADS / 32; ++w) { c += s_cnt[l][w]; first = min(first, s_first[l][w]); }
        M.block_count[(size_t)l * M.nblocks + blockIdx.x] = c;
        M.block_first[(size_t)l * M.nblocks + blockIdx.x] = first;
    }
}

// one CTA of 1024 threads per light: exclusive scan of the light's CTA counts (block_base = flagged items before the CTA over ALL
// groups and kinds; the scatter pass subtracts the prefix of its (group, kind) run's first CTA), then per group the totals, the
// first silhouette edge and the layout
__global__ void __launch_bounds__(1024) shadow_layout_kernel(ShadowMulti M)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const uint32_t l = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t* cnt = M.block_count + (size_t)l * M.nblocks;
    const uint32_t* first = M.block_first + (size_t)l * M.nblocks;
    uint32_t* base = M.block_base + (size_t)l * M.nblocks;
    const int8_t* facings = M.facings + l * M.facing_stride;
    const LightCfg C = M.cfg[l];
    if (tid == 0) s_carry = 0;
    __syncthreads();
    constexpr int PER = 8;                                  // consecutive CTA counts per thread and round
    for (uint32_t b0 = 0; b0 < M.nblocks; b0 += 1024 * PER) {
        const uint32_t b = b0 + tid * PER;
        uint32_t v[PER], sum = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { v[i] = b + i < M.nblocks ? cnt[b + i] : 0u; sum += v[i]; }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t a = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (unsigned)o) inc += a; }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        uint32_t wb = 0, tot = 0;
        for (int w = 0; w < 32; ++w) { if (w < (int)wid) wb += s_warp[w]; tot += s_warp[w]; }
        uint32_t run = s_carry + wb + inc - sum;
#pragma unroll
        for (int i = 0; i < PER; ++i) { if (b + i < M.nblocks) base[b + i] = run; run += v[i]; }
        __syncthreads();
        if (tid == 0) s_carry += tot;
        __syncthreads();
    }
    // ---- per group (a handful): serial over the groups, the thread-parallel part was the scan
    if (tid == 0) {
        const uint32_t total = s_carry;
        auto prefix = [&](uint32_t blk) { return blk < M.nblocks ? base[blk] : total; };      // flagged items before block `blk`
        const uint32_t per = 3u + 3u * C.two_tris + 3u * C.fan;
        unsigned long long need = 0;
        for (uint32_t g = 0; g < M.ngroups; ++g) {
            const GroupStatic S = M.gstatic[g];
            const uint32_t nsil = prefix(S.edge_blk1) - prefix(S.edge_blk0), nlit = prefix(S.tri_blk1) - prefix(S.tri_blk0);
            need += (unsigned long long)nsil * per - ((C.fan && nsil) ? 3u : 0u);
            need += (unsigned long long)nlit * 3u * (C.dark_cap + C.light_cap);
        }
        const bool fits = need <= M.region_indices;
        M.result[2 * l] = (uint32_t)min(need, 0xffffffffull);
        M.result[2 * l + 1] = fits ? 0u : 1u;
        uint32_t num = l * M.region_indices;
        for (uint32_t g = 0; g < M.ngroups; ++g) {
            const GroupStatic S = M.gstatic[g];
            const uint32_t nsil = prefix(S.edge_blk1) - prefix(S.edge_blk0), nlit = prefix(S.tri_blk1) - prefix(S.tri_blk0);
            GroupLayout L;
            uint32_t* R = M.ranges + ((size_t)l * M.ngroups + g) * 4;
            L.vertex_count = S.vertex_count;
            R[0] = num; R[1] = 0; R[2] = 0; R[3] = 0;
            L.edge_base = num;
            num += nsil * per - ((C.fan && nsil) ? 3u : 0u);
            L.dark_start = 0;
            if (C.fan && nsil) {
                uint32_t blk = S.edge_blk0;                                  // the first CTA of the group's edges that flagged one
                while (blk < S.edge_blk1 && cnt[blk] == 0) ++blk;
                const uint32_t* e = M.edges + 5 * (size_t)(M.blocks[blk].start + first[blk]);
                L.dark_start = (facings[e[0]] ? e[2] : e[3]) + L.vertex_count;
            }
            L.dark_base = num;
            if (C.dark_cap) num += 3 * nlit;
            if (C.light_cap) {
                if (M.separate_light_cap) { R[1] = num - R[0]; R[2] = num; }
                L.light_base = num;
                num += 3 * nlit;
                if (M.separate_light_cap) R[3] = num - R[2]; else R[1] = num - R[0];
            } else {
                L.light_base = num;
                R[1] = num - R[0];
            }
            L.pad[0] = L.pad[1] = L.pad[2] = 0;
            M.layout[(size_t)l * M.ngroups + g] = L;
        }
    }
}

template <typename IndexT>
__global__ void __launch_bounds__(SH_THREADS) shadow_scatter_multi_kernel(ShadowMulti M, uint32_t num_lights)
{
    __shared__ uint32_t s_warp[SH_MAX_LIGHTS][SH_THREADS / 32];
    const BlockDesc D = M.blocks[blockIdx.x];
    const ItemData it = load_item(M, D, threadIdx.x);
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t mine = 0, before[SH_MAX_LIGHTS];
#pragma unroll
    for (uint32_t l = 0; l < SH_MAX_LIGHTS; ++l) {
        before[l] = 0;
        if (l < num_lights) {
            const bool f = flagged_item(it, D.kind, M.facings + l * M.facing_stride);
            const uint32_t b = __ballot_sync(0xffffffffu, f);
            if (lane == 0) s_warp[l][wid] = __popc(b);
            before[l] = __popc(b & ((1u << lane) - 1u));
            mine |= (f ? 1u : 0u) << l;
        }
    }
    __syncthreads();
    if (!mine) return;
    uint32_t T0 = 0, T1 = 0, T2 = 0;
    if (D.kind != 0) { const uint32_t* T = M.tris + 3 * (size_t)it.a; T0 = T[0]; T1 = T[1]; T2 = T[2]; }
    IndexT* out = static_cast<IndexT*>(M.indices);
#pragma unroll
    for (uint32_t l = 0; l < SH_MAX_LIGHTS; ++l) {
        if (!((mine >> l) & 1u) || M.result[2 * l + 1]) continue;     // (a light whose indices do not fit its region writes nothing)
        uint32_t rank = M.block_base[(size_t)l * M.nblocks + blockIdx.x] - M.block_base[(size_t)l * M.nblocks + M.seg_first[blockIdx.x]] + before[l];
        for (unsigned w = 0; w < wid; ++w) rank += s_warp[l][w];
        const GroupLayout G = M.layout[(size_t)l * M.ngroups + D.group];
        const LightCfg C = M.cfg[l];
        const uint32_t ovc = G.vertex_count;
        if (D.kind == 0) {
            uint32_t v0 = it.c, v1 = it.d;
            if (!(M.facings + l * M.facing_stride)[it.a]) { const uint32_t t = v0; v0 = v1; v1 = t; }
            const uint32_t per = 3u + 3u * C.two_tris + 3u * C.fan;
            IndexT* p = out + G.edge_base + rank * per - (rank ? 3u * C.fan : 0u);
            *p++ = (IndexT)v1; *p++ = (IndexT)v0; *p++ = (IndexT)(v0 + ovc);
            if (C.two_tris) { *p++ = (IndexT)(v0 + ovc); *p++ = (IndexT)(v1 + ovc); *p++ = (IndexT)v1; }
            if (C.fan && rank) { *p++ = (IndexT)G.dark_start; *p++ = (IndexT)(v1 + ovc); *p++ = (IndexT)(v0 + ovc); }
        } else {
            if (C.dark_cap) {
                IndexT* p = out + G.dark_base + 3 * rank;
                p[0] = (IndexT)(T1 + ovc); p[1] = (IndexT)(T0 + ovc); p[2] = (IndexT)(T2 + ovc);
            }
            if (C.light_cap) {
                IndexT* p = out + G.light_base + 3 * rank;
                p[0] = (IndexT)T0; p[1] = (IndexT)T1; p[2] = (IndexT)T2;
            }
        }
    }
}

}  // namespace

struct ogrecu_edges {
    ogrecu_ctx* ctx;
    uint32_t num_groups, num_edges, num_tris, num_blocks;
    std::vector<uint32_t> groups;           // [g][5]
    std::vector<BlockDesc> blocks;
    std::vector<uint32_t> edges_host;       // kept for the one edge the McGuire fan apex needs
    BlockDesc* d_blocks; uint32_t* d_edges; uint32_t* d_tris;
    uint32_t *d_count, *d_first, *d_base; GroupLayout* d_layout;
    std::vector<uint32_t> h_count, h_first, h_base;
    std::vector<GroupLayout> h_layout;
    // several lights per call (ogrecu_shadow_volumes_multi_dev): static block tables, [L][...] work arrays made on first use
    GroupStatic* d_gstatic = nullptr; uint32_t* d_seg_first = nullptr;
    uint32_t *m_count = nullptr, *m_first = nullptr, *m_base = nullptr, *m_ranges = nullptr, *m_result = nullptr;
    GroupLayout* m_layout = nullptr;
};

extern "C" {

// The reference's choice between the McGuire fan and full caps (OgreShadowCaster.cpp:134-189,203-204): host arithmetic
// on the caster's light-cap bounds, float32 in the reference's order.
int ogrecu_shadow_use_mcguire(uint32_t num_groups, const float* b, const float* light4)
{
    if (!b || !light4 || num_groups > 1) return 0;
    if (light4[3] == 0.0f) return 1;
    const float lx = light4[0], ly = light4[1], lz = light4[2];
    if (b[0] <= lx && lx <= b[3] && b[1] <= ly && ly <= b[4] && b[2] <= lz && lz <= b[5]) return 0;      // light inside the bounds
    auto len = [](float x, float y, float z) { return sqrtf(((0.0f + x * x) + y * y) + z * z); };
    const float half = len((b[3] - b[0]) * 0.5f, (b[4] - b[1]) * 0.5f, (b[5] - b[2]) * 0.5f);
    const float dist = len((b[3] + b[0]) * 0.5f - lx, (b[4] + b[1]) * 0.5f - ly, (b[5] + b[2]) * 0.5f - lz);
    if (!(half / dist > 0.5f)) return 1;                                                               // bounding sphere under 60 degrees
    // corners on the lit / unlit border of the box, then the widest angle between two of them seen from the light
    const unsigned mask = (lx < b[0] ? 1u : 0u) | (lx > b[3] ? 2u : 0u) | (ly < b[1] ? 4u : 0u) | (ly > b[4] ? 8u : 0u) |
                          (lz < b[2] ? 16u : 0u) | (lz > b[5] ? 32u : 0u);
    float cx[8], cy[8], cz[8];
    unsigned n = 0;
    for (unsigned k = 0; k < 8; ++k) {
        const unsigned sides = ((k & 1) ? 2u : 1u) | ((k & 2) ? 8u : 4u) | ((k & 4) ? 32u : 16u);      // x: left/right, y: bottom/top, z: far/near
        if ((mask & sides) != 0 && (mask & sides) != sides) { cx[n] = b[(k & 1) ? 3 : 0]; cy[n] = b[(k & 2) ? 4 : 1]; cz[n] = b[(k & 4) ? 5 : 2]; ++n; }
    }
    float cos_angle = 1.0f;
    for (unsigned i = 0; i + 1 < n; ++i)
        for (unsigned j = i + 1; j < n; ++j) {
            float ax = cx[i] - lx, ay = cy[i] - ly, az = cz[i] - lz, bx = cx[j] - lx, by = cy[j] - ly, bz = cz[j] - lz;
            const float la = len(ax, ay, az), lb = len(bx, by, bz);
            if (la > 0.0f) { const float inv = 1.0f / la; ax *= inv; ay *= inv; az *= inv; }
            if (lb > 0.0f) { const float inv = 1.0f / lb; bx *= inv; by *= inv; bz *= inv; }
            const float c = ((0.0f + ax * bx) + ay * by) + az * bz;
            if (cos_angle > c) cos_angle = c;
        }
    return cos_angle < 0.5f ? 0 : 1;
}

int ogrecu_edges_create(ogrecu_ctx* ctx, uint32_t num_groups, const uint32_t* groups5, const uint32_t* edges5, const uint32_t* tris3,
                        uint32_t num_tris, ogrecu_edges** out)
{
    OGRECU_REQUIRE(ctx && out && (num_groups == 0 || groups5), "NULL argument");
    *out = nullptr;
    uint32_t ne = 0;
    for (uint32_t g = 0; g < num_groups; ++g) {
        const uint32_t* G = groups5 + 5 * g;
        OGRECU_REQUIRE(G[0] == ne, "edge ranges of the groups must be consecutive");
        OGRECU_REQUIRE((uint64_t)G[2] + G[3] <= num_tris, "triangle range outside the triangle list");
        ne += G[1];
    }
    OGRECU_REQUIRE((ne == 0 || edges5) && (num_tris == 0 || tris3), "NULL edge or triangle table");
    for (uint32_t e = 0; e < ne; ++e)
        OGRECU_REQUIRE(edges5[5 * (size_t)e] < num_tris && (edges5[5 * (size_t)e + 4] || edges5[5 * (size_t)e + 1] < num_tris), "edge names a triangle outside the list");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    ogrecu_edges* E = new ogrecu_edges();
    E->ctx = ctx; E->num_groups = num_groups; E->num_edges = ne; E->num_tris = num_tris;
    E->groups.assign(groups5, groups5 + 5 * (size_t)num_groups);
    E->edges_host.assign(edges5, edges5 + 5 * (size_t)ne);
    for (uint32_t g = 0; g < num_groups; ++g) {
        const uint32_t* G = groups5 + 5 * g;
        for (uint32_t s = 0; s < G[1]; s += SH_THREADS) E->blocks.push_back(BlockDesc{g, 0, G[0] + s, std::min<uint32_t>(SH_THREADS, G[1] - s)});
        for (uint32_t s = 0; s < G[3]; s += SH_THREADS) E->blocks.push_back(BlockDesc{g, 1, G[2] + s, std::min<uint32_t>(SH_THREADS, G[3] - s)});
    }
    E->num_blocks = (uint32_t)E->blocks.size();
    const size_t nb = std::max<size_t>(E->num_blocks, 1);
    E->d_blocks = nullptr; E->d_edges = nullptr; E->d_tris = nullptr; E->d_count = E->d_first = E->d_base = nullptr; E->d_layout = nullptr;
    cudaError_t e = cudaMalloc((void**)&E->d_blocks, nb * sizeof(BlockDesc));
    if (e == cudaSuccess) e = cudaMalloc((void**)&E->d_edges, std::max<size_t>(ne, 1) * 20);
    if (e == cudaSuccess) e = cudaMalloc((void**)&E->d_tris, std::max<size_t>(num_tris, 1) * 12);
    if (e == cudaSuccess) e = cudaMalloc((void**)&E->d_count, nb * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&E->d_first, nb * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&E->d_base, nb * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&E->d_layout, std::max<size_t>(num_groups, 1) * sizeof(GroupLayout));
    if (e == cudaSuccess && E->num_blocks) e = cudaMemcpyAsync(E->d_blocks, E->blocks.data(), E->num_blocks * sizeof(BlockDesc), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && ne) e = cudaMemcpyAsync(E->d_edges, edges5, (size_t)ne * 20, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && num_tris) e = cudaMemcpyAsync(E->d_tris, tris3, (size_t)num_tris * 12, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ogrecu_edges_destroy(E); return cuda_fail(e, "ogrecu_edges_create", __FILE__, __LINE__); }
    E->h_count.resize(nb); E->h_first.resize(nb); E->h_base.resize(nb); E->h_layout.resize(std::max<size_t>(num_groups, 1));
    {
        std::vector<GroupStatic> gs(std::max<size_t>(num_groups, 1));
        std::vector<uint32_t> seg(nb, 0);
        for (uint32_t g = 0; g < num_groups; ++g) { gs[g].edge_blk0 = gs[g].edge_blk1 = gs[g].tri_blk0 = gs[g].tri_blk1 = 0; gs[g].vertex_count = groups5[5 * g + 4]; }
        uint32_t b = 0;
        for (uint32_t g = 0; g < num_groups; ++g) {
            gs[g].edge_blk0 = b;
            while (b < E->num_blocks && E->blocks[b].group == g && E->blocks[b].kind == 0) { seg[b] = gs[g].edge_blk0; ++b; }
            gs[g].edge_blk1 = gs[g].tri_blk0 = b;
            while (b < E->num_blocks && E->blocks[b].group == g && E->blocks[b].kind == 1) { seg[b] = gs[g].tri_blk0; ++b; }
            gs[g].tri_blk1 = b;
        }
        e = cudaMalloc((void**)&E->d_gstatic, gs.size() * sizeof(GroupStatic));
        if (e == cudaSuccess) e = cudaMalloc((void**)&E->d_seg_first, nb * 4);
        if (e == cudaSuccess) e = cudaMemcpy(E->d_gstatic, gs.data(), gs.size() * sizeof(GroupStatic), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(E->d_seg_first, seg.data(), nb * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { ogrecu_edges_destroy(E); return cuda_fail(e, "ogrecu_edges_create", __FILE__, __LINE__); }
    }
    *out = E;
    return OGRECU_OK;
}

int ogrecu_edges_destroy(ogrecu_edges* E)
{
    if (!E) return OGRECU_OK;
    cudaSetDevice(E->ctx->device);
    cudaStreamSynchronize(E->ctx->stream);
    for (void* p : {(void*)E->d_blocks, (void*)E->d_edges, (void*)E->d_tris, (void*)E->d_count, (void*)E->d_first, (void*)E->d_base, (void*)E->d_layout,
                    (void*)E->d_gstatic, (void*)E->d_seg_first, (void*)E->m_count, (void*)E->m_first, (void*)E->m_base, (void*)E->m_ranges, (void*)E->m_result,
                    (void*)E->m_layout})
        if (p) cudaFree(p);
    delete E;
    return OGRECU_OK;
}

int ogrecu_shadow_volume_dev(ogrecu_edges* E, const char* facings_dev, int directional, int use_mcguire, uint32_t flags,
                             int separate_light_cap, uint32_t index_type, void* indices_dev, uint32_t buffer_indices,
                             uint32_t* used_inout, uint32_t* ranges4_host, uint32_t* emitted)
{
    OGRECU_REQUIRE(E && used_inout && ranges4_host, "NULL argument");
    OGRECU_REQUIRE(index_type == 2 || index_type == 4, "index_type is 2 or 4");
    if (emitted) *emitted = 0;
    ogrecu_ctx* ctx = E->ctx;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    const bool dark = (flags & OGRECU_SRF_INCLUDE_DARK_CAP) != 0, light = (flags & OGRECU_SRF_INCLUDE_LIGHT_CAP) != 0;
    const bool two = !(directional && (flags & OGRECU_SRF_EXTRUDE_TO_INFINITY));
    const bool fan = use_mcguire && dark;
    ShadowParams P{};
    P.blocks = E->d_blocks; P.edges = E->d_edges; P.tris = E->d_tris; P.facings = reinterpret_cast<const int8_t*>(facings_dev);
    P.block_count = E->d_count; P.block_first = E->d_first; P.block_base = E->d_base; P.layout = E->d_layout; P.indices = indices_dev;
    P.two_tris = two; P.fan = fan; P.dark_cap = dark && !use_mcguire; P.light_cap = light;
    if (E->num_blocks) {
        OGRECU_REQUIRE(facings_dev && indices_dev, "NULL device pointer");
        shadow_count_kernel<<<E->num_blocks, SH_THREADS, 0, ctx->stream>>>(P);
        OGRECU_CUDA(cudaGetLastError());
        count_launch(ctx);
        OGRECU_CUDA(cudaMemcpyAsync(E->h_count.data(), E->d_count, E->num_blocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
        OGRECU_CUDA(cudaMemcpyAsync(E->h_first.data(), E->d_first, E->num_blocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
        OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    // ---- per group: silhouette edges, lit triangles, the first silhouette edge; CTA bases
    std::vector<uint32_t> nsil(E->num_groups, 0), nlit(E->num_groups, 0), first_edge(E->num_groups, 0xffffffffu);
    for (uint32_t b = 0; b < E->num_blocks; ++b) {
        const BlockDesc& D = E->blocks[b];
        uint32_t& acc = D.kind == 0 ? nsil[D.group] : nlit[D.group];
        E->h_base[b] = acc;
        if (D.kind == 0 && acc == 0 && E->h_count[b]) first_edge[D.group] = D.start + E->h_first[b];
        acc += E->h_count[b];
    }
    uint64_t pre = 0;
    const uint32_t per = 3u + 3u * two + 3u * fan;
    for (uint32_t g = 0; g < E->num_groups; ++g) {
        pre += (uint64_t)nsil[g] * per - (fan && nsil[g] ? 3u : 0u);
        pre += (uint64_t)nlit[g] * 3u * ((P.dark_cap ? 1u : 0u) + (light ? 1u : 0u));
    }
    if (pre > buffer_indices) {
        set_error("ogrecu_shadow_volume_dev: %llu indices do not fit the index buffer of %u (the reference grows the SceneManager's shadow index buffer here)",
                  (unsigned long long)pre, buffer_indices);
        return OGRECU_ERR_INVALID;
    }
    if ((uint64_t)*used_inout + pre > buffer_indices) *used_inout = 0;           // wrap, as OgreShadowCaster.cpp:321-324
    uint32_t num = *used_inout;
    for (uint32_t g = 0; g < E->num_groups; ++g) {
        GroupLayout& L = E->h_layout[g];
        uint32_t* R = ranges4_host + 4 * g;
        L.vertex_count = E->groups[5 * g + 4];
        R[0] = num; R[2] = 0; R[3] = 0;
        L.edge_base = num;
        num += nsil[g] * per - (fan && nsil[g] ? 3u : 0u);
        L.dark_start = 0;
        if (fan && first_edge[g] != 0xffffffffu) {
            // apex of the McGuire fan: v0 (after the facing swap) of the group's first silhouette edge + vertex count.  The
            // swap needs that edge's light facing: for a silhouette edge it is "tri 0 lit" unless tri 0 is unlit, in which
            // case the edge is non-degenerate and tri 1 is lit — one byte from the device decides.
            const uint32_t* e = &E->edges_host[5 * (size_t)first_edge[g]];
            char lf = 0;
            OGRECU_CUDA(cudaMemcpyAsync(&lf, facings_dev + e[0], 1, cudaMemcpyDeviceToHost, ctx->stream));
            OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
            L.dark_start = (lf ? e[2] : e[3]) + L.vertex_count;
        }
        L.dark_base = num;
        if (P.dark_cap) num += 3 * nlit[g];
        if (light) {
            if (separate_light_cap) { R[1] = num - R[0]; R[2] = num; }
            L.light_base = num;
            num += 3 * nlit[g];
            if (separate_light_cap) R[3] = num - R[2]; else R[1] = num - R[0];
        } else {
            L.light_base = num;
            R[1] = num - R[0];
        }
    }
    if (E->num_blocks && pre) {
        OGRECU_CUDA(cudaMemcpyAsync(E->d_base, E->h_base.data(), E->num_blocks * 4, cudaMemcpyHostToDevice, ctx->stream));
        OGRECU_CUDA(cudaMemcpyAsync(E->d_layout, E->h_layout.data(), E->num_groups * sizeof(GroupLayout), cudaMemcpyHostToDevice, ctx->stream));
        if (index_type == 2) shadow_scatter_kernel<uint16_t><<<E->num_blocks, SH_THREADS, 0, ctx->stream>>>(P);
        else shadow_scatter_kernel<uint32_t><<<E->num_blocks, SH_THREADS, 0, ctx->stream>>>(P);
        OGRECU_CUDA(cudaGetLastError());
        count_launch(ctx);
        OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));      // h_base / h_layout are reused by the next call
    }
    *used_inout = num;
    if (emitted) *emitted = (uint32_t)pre;
    return OGRECU_OK;
}

int ogrecu_shadow_volumes_multi_dev(ogrecu_edges* E, uint32_t num_lights, const char* facings_dev, size_t facing_stride_bytes,
                                    const int* directional, const int* use_mcguire, const uint32_t* flags, int separate_light_cap,
                                    uint32_t index_type, void* indices_dev, uint32_t region_indices, uint32_t* ranges4_host,
                                    uint32_t* result2_host)
{
    OGRECU_REQUIRE(E && directional && use_mcguire && flags, "NULL argument");
    OGRECU_REQUIRE(num_lights >= 1 && num_lights <= SH_MAX_LIGHTS, "1..8 lights per call");
    OGRECU_REQUIRE(index_type == 2 || index_type == 4, "index_type is 2 or 4");
    OGRECU_REQUIRE((uint64_t)region_indices * num_lights <= 0xffffffffull, "index regions exceed 2^32 indices");
    ogrecu_ctx* ctx = E->ctx;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    const size_t nb = std::max<size_t>(E->num_blocks, 1), ng = std::max<size_t>(E->num_groups, 1);
    if (!E->m_count) {
        cudaError_t e = cudaMalloc((void**)&E->m_count, SH_MAX_LIGHTS * nb * 4);
        if (e == cudaSuccess) e = cudaMalloc((void**)&E->m_first, SH_MAX_LIGHTS * nb * 4);
        if (e == cudaSuccess) e = cudaMalloc((void**)&E->m_base, SH_MAX_LIGHTS * nb * 4);
        if (e == cudaSuccess) e = cudaMalloc((void**)&E->m_layout, SH_MAX_LIGHTS * ng * sizeof(GroupLayout));
        if (e == cudaSuccess) e = cudaMalloc((void**)&E->m_ranges, SH_MAX_LIGHTS * ng * 16);
        if (e == cudaSuccess) e = cudaMalloc((void**)&E->m_result, SH_MAX_LIGHTS * 8);
        if (e != cudaSuccess) return cuda_fail(e, "ogrecu_shadow_volumes_multi_dev", __FILE__, __LINE__);
    }
    ShadowMulti M{};
    M.blocks = E->d_blocks; M.edges = E->d_edges; M.tris = E->d_tris;
    M.facings = reinterpret_cast<const int8_t*>(facings_dev); M.facing_stride = facing_stride_bytes;
    M.block_count = E->m_count; M.block_first = E->m_first; M.block_base = E->m_base; M.layout = E->m_layout;
    M.gstatic = E->d_gstatic; M.seg_first = E->d_seg_first; M.ranges = E->m_ranges; M.result = E->m_result;
    M.indices = indices_dev; M.nblocks = E->num_blocks; M.ngroups = E->num_groups; M.region_indices = region_indices;
    M.separate_light_cap = separate_light_cap ? 1u : 0u;
    for (uint32_t l = 0; l < num_lights; ++l) {
        const bool dark = (flags[l] & OGRECU_SRF_INCLUDE_DARK_CAP) != 0, light = (flags[l] & OGRECU_SRF_INCLUDE_LIGHT_CAP) != 0;
        M.cfg[l].two_tris = !(directional[l] && (flags[l] & OGRECU_SRF_EXTRUDE_TO_INFINITY));
        M.cfg[l].fan = use_mcguire[l] && dark;
        M.cfg[l].dark_cap = dark && !use_mcguire[l];
        M.cfg[l].light_cap = light;
    }
    if (E->num_blocks) {
        OGRECU_REQUIRE(facings_dev && indices_dev, "NULL device pointer");
        shadow_count_multi_kernel<<<E->num_blocks, SH_THREADS, 0, ctx->stream>>>(M, num_lights);
    }
    shadow_layout_kernel<<<num_lights, 1024, 0, ctx->stream>>>(M);
    if (E->num_blocks) {
        if (index_type == 2) shadow_scatter_multi_kernel<uint16_t><<<E->num_blocks, SH_THREADS, 0, ctx->stream>>>(M, num_lights);
        else shadow_scatter_multi_kernel<uint32_t><<<E->num_blocks, SH_THREADS, 0, ctx->stream>>>(M, num_lights);
    }
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx, E->num_blocks ? 3 : 1);
    // the ShadowRenderable ranges and the per-light totals follow the kernels on the stream; the host reads them after its next sync
    if (ranges4_host) OGRECU_CUDA(cudaMemcpyAsync(ranges4_host, E->m_ranges, (size_t)num_lights * E->num_groups * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (result2_host) OGRECU_CUDA(cudaMemcpyAsync(result2_host, E->m_result, (size_t)num_lights * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return OGRECU_OK;
}

}  // extern "C"
