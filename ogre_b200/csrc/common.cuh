// common.cuh — shared internals of libogrecu.so (not part of the public C ABI).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stddef.h>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <map>
#include <string>
#include <vector>

#include "../../include/ogrecu.h"

namespace ogrecu {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define OGRECU_CUDA(expr)                                                            \
    do {                                                                             \
        cudaError_t e__ = (expr);                                                    \
        if (e__ != cudaSuccess) return ::ogrecu::cuda_fail(e__, #expr, __FILE__, __LINE__); \
    } while (0)

#define OGRECU_REQUIRE(cond, msg)                                     \
    do {                                                              \
        if (!(cond)) { ::ogrecu::set_error("%s: %s", __func__, msg); return OGRECU_ERR_INVALID; } \
    } while (0)

#define OGRECU_CHECK(expr)             \
    do {                               \
        int rc__ = (expr);             \
        if (rc__) return rc__;         \
    } while (0)

// Growable device scratch (never shrinks; freed at shutdown).
struct Scratch {
    void* ptr = nullptr;
    size_t cap = 0;
};

}  // namespace ogrecu

constexpr int OGRECU_MAX_PEERS = 7;      // a device group is at most 8 devices: the root and seven peers

struct ogrecu_buf {
    ogrecu_ctx* ctx;
    size_t size;
    void* dev;          // device storage
    // device group (ogrecu_init_devices): one replica per peer device, brought up to date over NVLink when a draw that runs on
    // that peer references the buffer.  [peer_lo, peer_hi) = bytes of the root copy the replica has not seen yet.
    void* peer_dev[OGRECU_MAX_PEERS];
    size_t peer_lo[OGRECU_MAX_PEERS], peer_hi[OGRECU_MAX_PEERS];
    uint8_t* shadow;    // pinned host shadow (allocated lazily on first map)
    bool mapped;
    int map_mode;
    size_t map_off, map_len;
    // device-side writes not yet visible in the shadow
    bool dev_dirty;
    size_t dev_dirty_lo, dev_dirty_hi;
    // last shadow->device upload still in flight on the stream (the shadow must not be
    // rewritten by the host before it completes)
    cudaEvent_t upload_ev;
    bool upload_pending;
    // hazard tracking of deferred work (see "optutil_batch"): the value of the context's epoch at the time the buffer was
    // referenced; a flush bumps the epoch, which clears every mark at once
    uint64_t batch_r = 0, batch_w = 0;      // read / written by a queued OptimisedUtil job
    uint64_t draw_ref = 0;                  // referenced by a recorded, unflushed draw
    // highest blend index of a (range, stride, weights, vertices) of this buffer, kept until the host next writes it
    struct MaxIdx { size_t off, stride, nw, n; unsigned max_idx; };
    std::vector<MaxIdx> maxidx_cache;
};

struct ogrecu_tex {
    ogrecu_ctx* ctx;
    uint32_t w, h;
    uint8_t* dev;       // RGBA8, tightly packed
    int has_zero_alpha; // any texel with alpha == 0 (enables the discard test)
    uint8_t* peer_dev[OGRECU_MAX_PEERS];     // device group: replicas, whole-texture refresh when stale
    bool peer_stale[OGRECU_MAX_PEERS];
};

struct ogrecu_ctx {
    int device;
    // device group (ogrecu_init_devices): the root context owns the buffers, textures, OptimisedUtil work and the host-facing
    // API; each peer is a full context of its own device (stream, scratch arenas) that only ever runs raster passes
    ogrecu_ctx* root = nullptr;             // peers: the group's root; the root and single-device contexts: NULL
    std::vector<ogrecu_ctx*> peers;         // root: the other devices, peers[k - 1] = device index k of the group
    int dev_index = 0;                      // position in the group (0 = root)
    cudaEvent_t group_ev = nullptr;         // "this context's stream got here": cross-device ordering
    int split_band_tiles = 4;               // split targets: tile rows per interleaved band
    struct GroupPool* pool = nullptr;       // root: one host thread per peer, so that the devices' passes are submitted side by side
    int split_mode = 0;                     // split targets: 0 = sort-last where the content allows, else bands | 1 = always bands
    cudaStream_t own_stream;
    cudaStream_t stream;        // current launch stream (own or external)
    uint64_t launches;
    int sm_count;
    // scratch arenas
    ogrecu::Scratch jobs;       // skin job descriptors
    ogrecu::Scratch stage_in;   // host-path staging (device side)
    ogrecu::Scratch stage_out;
    ogrecu::Scratch lane_counter;   // work-unit counter of skin_lane_kernel
    ogrecu::Scratch lane_ipal;      // the frame's palettes repacked for skin_lane_kernel
    int profile_depth = 0;          // open ogrecu_profile_begin ranges
    int lane_reuse_repack = 0;      // measurement switch: ogrecu_anim_frame_dev keeps the previous call's repacked palettes
    void* pinned;               // pinned staging (host side)
    size_t pinned_cap;
    // the raster passes' draw / target tables: two pinned blocks used in turn, each guarded by an event, so that recording a pass
    // does not have to wait for the stream (the single block needed a cudaStreamSynchronize before every pass)
    uint8_t* tab_pin[2] = {nullptr, nullptr};
    size_t tab_cap[2] = {0, 0};
    cudaEvent_t tab_ev[2] = {nullptr, nullptr};
    bool tab_used[2] = {false, false};
    int tab_slot = 0;
    // shadow registry: shadow base address -> buffer
    std::map<uintptr_t, ogrecu_buf*> shadows;
    std::mutex shadows_mutex;       // buffers may be created / destroyed from resource threads (see the threading contract in ogrecu.h)
    std::vector<ogrecu_rt*> targets;
    // development switches read once from the environment (A/B measurements, never a fallback):
    //   OGRECU_SKIN_MODE  0 auto | 1 tile kernels without TMA (cooperative copies) | 2 one-thread-per-vertex kernel |
    //                     3 single-instance tile kernel even for >= 2 instances (skin_tile.cu instead of skin_pair.cu)
    //   OGRECU_SKIN_IPC   instances per CTA of the tile kernel (0 = auto)
    //   OGRECU_RASTER_MODE 0 auto (visibility path where it is equivalent) | 1 always the in-order path
    //   OGRECU_OPTUTIL_SYNC 1 = every OptimisedUtil host call waits for the device (default 0: deferred, see optutil_sync)
    //   OGRECU_RASTER_XF   1 (default) transform each vertex once per flush when the draw gives a vertex count | 0 per index
    int skin_mode, skin_ipc, raster_mode, raster_xf;
    // OptimisedUtil host entry points whose results all land in resolved device buffers do not wait for the device
    // (config option "optutil_sync" = 0, the default): the next lock / read-back / flush of those buffers is the sync point.
    // Their small host inputs (bone palettes) go through a ring of pinned slots, each guarded by an event.
    int optutil_sync = 0;
    static constexpr int RING_SLOTS = 8;
    static constexpr size_t RING_SLOT_BYTES = 64 * 1024;
    uint8_t* ring = nullptr;
    cudaEvent_t ring_ev[RING_SLOTS] = {};
    bool ring_used[RING_SLOTS] = {};
    int ring_next = 0;
    int skel_mode = 0;          // 0: lanes = instances (skel_palette_kernel2) | 1: first version, a lane per bone (A/B)
    // deferred, batched OptimisedUtil ("optutil_batch" = 1): softwareVertexSkinning host calls whose operands all live in
    // ogrecu_bufs are queued (descriptor + a copy of the palette) and run as ONE batched launch at the next point that needs
    // their results (a draw flush, a map of one of their buffers, any other entry point that may read them, ogrecu_sync)
    int optutil_batch = 0;
    uint64_t batch_epoch = 1, draw_epoch = 1;
    std::vector<ogrecu_skin_job> batch_jobs;
    float* batch_palpin[2] = {nullptr, nullptr};     // pinned, per arena: the queued jobs' palettes are gathered straight into it
    size_t batch_pal_cap[2] = {0, 0}, batch_pal_floats = 0;
    uint8_t* batch_pin[2] = {nullptr, nullptr};
    size_t batch_pin_cap[2] = {0, 0};
    cudaEvent_t batch_ev[2] = {nullptr, nullptr};
    bool batch_ev_used[2] = {false, false};
    int batch_slot = 0;
    ogrecu::Scratch batch_dev;
    uint64_t batch_flushes = 0, batch_jobs_run = 0;
    std::map<uintptr_t, ogrecu_buf*> dev_registry;      // device base address -> buffer (guarded by shadows_mutex)
    // side streams for independent launches of one call (ogrecu_anim_instanced_multi_dev): created on first use
    cudaStream_t side_stream[2] = {nullptr, nullptr};
    cudaEvent_t side_fork = nullptr, side_join[2] = {nullptr, nullptr};
};

namespace ogrecu {
inline ogrecu_ctx* root_of(ogrecu_ctx* ctx) { return ctx->root ? ctx->root : ctx; }
// fn(0) on the calling thread, fn(1) .. fn(n-1) on the group's worker threads, all at once; the first failure's status (and its
// error text) is returned.  Each fn(k) must touch only task k's state and set its own device.  Serial when there is no pool.
int group_parallel(ogrecu_ctx* root, size_t n, const std::function<int(size_t)>& fn);
void group_pool_create(ogrecu_ctx* root);
void group_pool_destroy(ogrecu_ctx* root);
// `other`'s stream continues only after everything `ctx`'s stream has been given so far
int stream_wait(ogrecu_ctx* other, ogrecu_ctx* ctx);
// the address `p` of the root's memory as device `peer` should read it: inside an ogrecu_buf -> the peer's replica, refreshed on
// the peer's stream if stale (call stream_wait(peer, root) first); anything else -> unchanged (peer access over NVLink)
int replica_ptr(ogrecu_ctx* peer, const void* p, const void** out);
int replica_tex(ogrecu_ctx* peer, ogrecu_tex* t, const uint8_t** out);
int scratch_reserve(ogrecu_ctx* ctx, Scratch& s, size_t bytes);
int pinned_reserve(ogrecu_ctx* ctx, size_t bytes);
inline void count_launch(ogrecu_ctx* ctx, unsigned n = 1) { ctx->launches += n; }
// the current device for a scope (targets placed on a peer of a device group are served on that peer); restored on exit
struct DeviceScope {
    int prev = -1;
    explicit DeviceScope(int d) { int cur = -1; cudaGetDevice(&cur); if (cur != d) { prev = cur; cudaSetDevice(d); } }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};
// NVTX range over a scope (the library's own flushes)
struct ProfileScope { explicit ProfileScope(const char* n) { nvtxRangePushA(n); } ~ProfileScope() { nvtxRangePop(); } };

inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
int rt_release(ogrecu_rt* rt);
void free_raster_scratch(ogrecu_ctx* ctx);                   // the context's raster arenas (no-op when it has none)
// deferred work (optutil_host.cu / raster.cu)
int flush_skin_batch(ogrecu_ctx* ctx);                       // run the queued OptimisedUtil jobs now (no-op when empty)
int flush_all_targets(ogrecu_ctx* ctx);                      // run every target's recorded draws (the queued jobs first)
ogrecu_buf* buf_of_device_ptr(ogrecu_ctx* ctx, const void* dev_ptr);
int launch_skin_batched(ogrecu_ctx* ctx, const ogrecu_skin_job* djobs, const uint32_t* dprefix, uint32_t n, unsigned blocks);
int validate_skin_job_public(const ogrecu_skin_job& j);
}  // namespace ogrecu
