// ctx.cu — context, memory and HardwareBuffer delegate of libogrecu.so.
//
// Reference roles served here (OGRE 14.6 tree):
//   - first-window initialisation of TinyRenderSystem (RenderSystems/Tiny/src/OgreTinyRenderSystem.cpp:146-194)
//   - the HardwareBuffer delegate contract (OgreMain/include/OgreHardwareBuffer.h:181-186,279-330):
//     lock()/unlock() become map()/unmap() on a pinned host shadow mirrored in HBM.
// There is no CPU fallback: without a CUDA device ogrecu_init fails.
#include "common.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace ogrecu {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    if (e == cudaErrorMemoryAllocation) return OGRECU_ERR_OOM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return OGRECU_ERR_NO_DEVICE;
    return OGRECU_ERR_CUDA;
}

int scratch_reserve(ogrecu_ctx* ctx, Scratch& s, size_t bytes)
{
    if (bytes <= s.cap) return OGRECU_OK;
    // growing: the old block may still be in use by queued kernels on the stream
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    if (s.ptr) OGRECU_CUDA(cudaFree(s.ptr));
    s.ptr = nullptr; s.cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    OGRECU_CUDA(cudaMalloc(&s.ptr, want));
    s.cap = want;
    return OGRECU_OK;
}

int pinned_reserve(ogrecu_ctx* ctx, size_t bytes)
{
    if (bytes <= ctx->pinned_cap) return OGRECU_OK;
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->pinned) OGRECU_CUDA(cudaFreeHost(ctx->pinned));
    ctx->pinned = nullptr; ctx->pinned_cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    OGRECU_CUDA(cudaHostAlloc(&ctx->pinned, want, cudaHostAllocDefault));
    ctx->pinned_cap = want;
    return OGRECU_OK;
}

}  // namespace ogrecu

using namespace ogrecu;


// bytes [lo, hi) of the root copy changed: every peer replica has to fetch them before its next use
static void replicas_stale(ogrecu_buf* b, size_t lo, size_t hi)
{
    const size_t n = b->ctx->peers.size();
    for (size_t k = 0; k < n; ++k) {
        if (b->peer_lo[k] >= b->peer_hi[k]) { b->peer_lo[k] = lo; b->peer_hi[k] = hi; }
        else { b->peer_lo[k] = std::min(b->peer_lo[k], lo); b->peer_hi[k] = std::max(b->peer_hi[k], hi); }
    }
}

namespace ogrecu {
int stream_wait(ogrecu_ctx* other, ogrecu_ctx* ctx)
{
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    OGRECU_CUDA(cudaEventRecord(ctx->group_ev, ctx->stream));
    OGRECU_CUDA(cudaSetDevice(other->device));
    OGRECU_CUDA(cudaStreamWaitEvent(other->stream, ctx->group_ev, 0));
    return OGRECU_OK;
}

int replica_ptr(ogrecu_ctx* peer, const void* p, const void** out)
{
    *out = p;
    ogrecu_ctx* root = peer->root;
    if (!root || !p) return OGRECU_OK;
    ogrecu_buf* b = buf_of_device_ptr(root, p);
    if (!b) return OGRECU_OK;                                   // not one of our buffers: the peer reads the root's memory directly
    const int k = peer->dev_index - 1;
    if (b->peer_lo[k] < b->peer_hi[k]) {
        OGRECU_CUDA(cudaSetDevice(peer->device));
        OGRECU_CUDA(cudaMemcpyPeerAsync((uint8_t*)b->peer_dev[k] + b->peer_lo[k], peer->device, (const uint8_t*)b->dev + b->peer_lo[k], root->device,
                                        b->peer_hi[k] - b->peer_lo[k], peer->stream));
        b->peer_lo[k] = b->peer_hi[k] = 0;
    }
    *out = (const uint8_t*)b->peer_dev[k] + ((const uint8_t*)p - (const uint8_t*)b->dev);
    return OGRECU_OK;
}

int replica_tex(ogrecu_ctx* peer, ogrecu_tex* t, const uint8_t** out)
{
    *out = t ? t->dev : nullptr;
    if (!t || !peer->root) return OGRECU_OK;
    const int k = peer->dev_index - 1;
    if (!t->peer_dev[k]) return OGRECU_OK;                      // created before the group existed: read it over NVLink
    if (t->peer_stale[k]) {
        OGRECU_CUDA(cudaSetDevice(peer->device));
        OGRECU_CUDA(cudaMemcpyPeerAsync(t->peer_dev[k], peer->device, t->dev, peer->root->device, (size_t)t->w * t->h * 4, peer->stream));
        t->peer_stale[k] = false;
    }
    *out = t->peer_dev[k];
    return OGRECU_OK;
}
}  // namespace ogrecu


/* ---- host threads of a device group: submitting the devices' passes one after the other from one thread costs ~60 us per device
 * (table building, a few dozen CUDA calls) and staggers them; with a worker per peer the passes start together ---- */
struct GroupPool {
    struct W {
        std::thread th;
        std::mutex m;
        std::condition_variable cv;
        const std::function<int(size_t)>* fn = nullptr;
        size_t arg = 0;
        int rc = OGRECU_OK;
        bool stop = false;
        char err[512] = "";
    };
    std::vector<W*> ws;
};

namespace ogrecu {
static void pool_worker(GroupPool::W* w)
{
    std::unique_lock<std::mutex> lk(w->m);
    for (;;) {
        w->cv.wait(lk, [&] { return w->fn != nullptr || w->stop; });
        if (w->stop) return;
        const std::function<int(size_t)>* fn = w->fn;
        lk.unlock();
        const int rc = (*fn)(w->arg);
        lk.lock();
        w->rc = rc;
        if (rc != OGRECU_OK) { strncpy(w->err, g_err, sizeof w->err - 1); w->err[sizeof w->err - 1] = 0; }
        w->fn = nullptr;
        w->cv.notify_all();
    }
}

void group_pool_create(ogrecu_ctx* root)
{
    if (root->pool || root->peers.empty()) return;
    root->pool = new GroupPool();
    for (size_t k = 0; k < root->peers.size(); ++k) {
        GroupPool::W* w = new GroupPool::W();
        w->th = std::thread(pool_worker, w);
        root->pool->ws.push_back(w);
    }
}

void group_pool_destroy(ogrecu_ctx* root)
{
    if (!root->pool) return;
    for (GroupPool::W* w : root->pool->ws) {
        { std::lock_guard<std::mutex> g(w->m); w->stop = true; }
        w->cv.notify_all();
        w->th.join();
        delete w;
    }
    delete root->pool;
    root->pool = nullptr;
}

int group_parallel(ogrecu_ctx* root, size_t n, const std::function<int(size_t)>& fn)
{
    if (n == 0) return OGRECU_OK;
    GroupPool* P = root->pool;
    const size_t nw = P ? std::min(P->ws.size(), n - 1) : 0;
    for (size_t i = 0; i < nw; ++i) {
        GroupPool::W* w = P->ws[i];
        { std::lock_guard<std::mutex> g(w->m); w->arg = i + 1; w->rc = OGRECU_OK; w->fn = &fn; }
        w->cv.notify_all();
    }
    int rc = fn(0);
    for (size_t k = nw + 1; k < n; ++k) { const int r = fn(k); if (rc == OGRECU_OK) rc = r; }      // more tasks than workers: the rest here
    for (size_t i = 0; i < nw; ++i) {
        GroupPool::W* w = P->ws[i];
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->fn == nullptr; });
        if (rc == OGRECU_OK && w->rc != OGRECU_OK) { rc = w->rc; set_error("%s", w->err); }
    }
    return rc;
}
}  // namespace ogrecu

extern "C" {

int ogrecu_abi_version(void) { return OGRECU_ABI_VERSION; }
const char* ogrecu_last_error_string(void) { return g_err; }

int ogrecu_profile_begin(ogrecu_ctx* ctx, const char* name)
{
    OGRECU_REQUIRE(ctx && name, "NULL argument");
    nvtxRangePushA(name);
    ++ctx->profile_depth;
    return OGRECU_OK;
}
int ogrecu_profile_end(ogrecu_ctx* ctx)
{
    OGRECU_REQUIRE(ctx, "ctx is NULL");
    OGRECU_REQUIRE(ctx->profile_depth > 0, "ogrecu_profile_end without a matching begin");
    nvtxRangePop();
    --ctx->profile_depth;
    return OGRECU_OK;
}
int ogrecu_profile_mark(ogrecu_ctx* ctx, const char* name)
{
    OGRECU_REQUIRE(ctx && name, "NULL argument");
    nvtxMarkA(name);
    return OGRECU_OK;
}

int ogrecu_device_count(int* count)
{
    OGRECU_REQUIRE(count, "count is NULL");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__); }
    return OGRECU_OK;
}

int ogrecu_init(int device, ogrecu_ctx** out_ctx)
{
    OGRECU_REQUIRE(out_ctx, "out_ctx is NULL");
    *out_ctx = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("ogrecu_init: no usable CUDA device (%s); libogrecu has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return OGRECU_ERR_NO_DEVICE;
    }
    OGRECU_REQUIRE(device >= 0 && device < n, "device index out of range");
    OGRECU_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    OGRECU_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("ogrecu_init: device %d is sm_%d%d; this library is built for sm_100a only", device,
                  prop.major, prop.minor);
        return OGRECU_ERR_NO_DEVICE;
    }
    ogrecu_ctx* ctx = new ogrecu_ctx();
    ctx->device = device;
    ctx->launches = 0;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->pinned = nullptr;
    ctx->pinned_cap = 0;
    const char* ev = getenv("OGRECU_SKIN_MODE");
    ctx->skin_mode = ev ? atoi(ev) : 0;
    ev = getenv("OGRECU_SKIN_IPC");
    ctx->skin_ipc = ev ? atoi(ev) : 0;
    ev = getenv("OGRECU_RASTER_MODE");
    ctx->raster_mode = ev ? atoi(ev) : 0;
    ev = getenv("OGRECU_RASTER_XF");
    ctx->raster_xf = ev ? atoi(ev) : 1;
    ev = getenv("OGRECU_OPTUTIL_SYNC");
    ctx->optutil_sync = ev ? (atoi(ev) != 0) : 0;
    e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    ctx->stream = ctx->own_stream;
    *out_ctx = ctx;
    return OGRECU_OK;
}

int ogrecu_init_devices(int num_devices, const int* devices, ogrecu_ctx** out_ctx)
{
    OGRECU_REQUIRE(out_ctx && devices, "NULL argument");
    *out_ctx = nullptr;
    OGRECU_REQUIRE(num_devices >= 1 && num_devices <= OGRECU_MAX_PEERS + 1, "1..8 devices per group");
    for (int a = 0; a < num_devices; ++a)
        for (int b = 0; b < a; ++b) OGRECU_REQUIRE(devices[a] != devices[b], "a device appears twice in the group");
    std::vector<ogrecu_ctx*> all;
    auto undo = [&]() { for (ogrecu_ctx* c : all) { c->root = nullptr; c->peers.clear(); ogrecu_shutdown(c); } };
    for (int k = 0; k < num_devices; ++k) {
        ogrecu_ctx* c = nullptr;
        const int rc = ogrecu_init(devices[k], &c);
        if (rc != OGRECU_OK) { undo(); return rc; }
        c->dev_index = k;
        all.push_back(c);
    }
    // every device reads and writes every other's memory directly (NVLink / NVSwitch): the split targets' planes live on the root
    // and the peers rasterise into them, replicas are refreshed by peer copies, raw device pointers stay valid group-wide
    for (int a = 0; a < num_devices; ++a)
        for (int b = 0; b < num_devices; ++b) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            if (!can) {
                set_error("ogrecu_init_devices: device %d cannot access device %d's memory (no peer access)", devices[a], devices[b]);
                undo();
                return OGRECU_ERR_UNSUPPORTED;
            }
            cudaSetDevice(devices[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { undo(); return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__); }
            cudaGetLastError();
        }
    for (int k = 0; k < num_devices; ++k) {
        cudaSetDevice(devices[k]);
        const cudaError_t e = cudaEventCreateWithFlags(&all[k]->group_ev, cudaEventDisableTiming);
        if (e != cudaSuccess) { undo(); return cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__); }
        if (k) { all[k]->root = all[0]; all[0]->peers.push_back(all[k]); }
    }
    cudaSetDevice(devices[0]);
    group_pool_create(all[0]);
    *out_ctx = all[0];
    return OGRECU_OK;
}

int ogrecu_num_devices(const ogrecu_ctx* ctx, int* num_devices)
{
    OGRECU_REQUIRE(ctx && num_devices, "NULL argument");
    *num_devices = 1 + (int)root_of(const_cast<ogrecu_ctx*>(ctx))->peers.size();
    return OGRECU_OK;
}

int ogrecu_shutdown(ogrecu_ctx* ctx)
{
    if (!ctx) return OGRECU_OK;
    if (ctx->root) { set_error("ogrecu_shutdown: shut the group's root context down, not a peer"); return OGRECU_ERR_INVALID; }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    group_pool_destroy(ctx);
    while (!ctx->targets.empty()) rt_release(ctx->targets.back());      // split targets first: their aliases name the peers
    {
        std::vector<ogrecu_ctx*> peers;
        peers.swap(ctx->peers);
        for (ogrecu_ctx* p : peers) { p->root = nullptr; ogrecu_shutdown(p); }
    }
    cudaSetDevice(ctx->device);
    free_raster_scratch(ctx);
    for (Scratch* s : {&ctx->jobs, &ctx->stage_in, &ctx->stage_out, &ctx->lane_counter, &ctx->lane_ipal, &ctx->batch_dev})
        if (s->ptr) cudaFree(s->ptr);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (int k = 0; k < 2; ++k) { if (ctx->tab_pin[k]) cudaFreeHost(ctx->tab_pin[k]); if (ctx->tab_ev[k]) cudaEventDestroy(ctx->tab_ev[k]); }
    for (int k = 0; k < 2; ++k) {
        if (ctx->batch_pin[k]) cudaFreeHost(ctx->batch_pin[k]);
        if (ctx->batch_palpin[k]) cudaFreeHost(ctx->batch_palpin[k]);
        if (ctx->batch_ev[k]) cudaEventDestroy(ctx->batch_ev[k]);
    }
    if (ctx->ring) cudaFreeHost(ctx->ring);
    for (cudaEvent_t e : ctx->ring_ev) if (e) cudaEventDestroy(e);
    for (int k = 0; k < 2; ++k) {
        if (ctx->side_stream[k]) { cudaStreamSynchronize(ctx->side_stream[k]); cudaStreamDestroy(ctx->side_stream[k]); }
        if (ctx->side_join[k]) cudaEventDestroy(ctx->side_join[k]);
    }
    if (ctx->side_fork) cudaEventDestroy(ctx->side_fork);
    if (ctx->group_ev) cudaEventDestroy(ctx->group_ev);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return OGRECU_OK;
}

int ogrecu_device(const ogrecu_ctx* ctx, int* device)
{
    OGRECU_REQUIRE(ctx && device, "NULL argument");
    *device = ctx->device;
    return OGRECU_OK;
}

int ogrecu_set_stream(ogrecu_ctx* ctx, void* cuda_stream)
{
    OGRECU_REQUIRE(ctx, "ctx is NULL");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return OGRECU_OK;
}

void* ogrecu_get_stream(ogrecu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int ogrecu_sync(ogrecu_ctx* ctx)
{
    OGRECU_REQUIRE(ctx, "ctx is NULL");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    OGRECU_CHECK(flush_skin_batch(root_of(ctx)));
    for (ogrecu_ctx* p : ctx->peers) {
        OGRECU_CUDA(cudaSetDevice(p->device));
        OGRECU_CUDA(cudaStreamSynchronize(p->stream));
    }
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    return OGRECU_OK;
}

int ogrecu_set_config_option(ogrecu_ctx* ctx, const char* name, int value)
{
    OGRECU_REQUIRE(ctx && name, "NULL argument");
    if (!strcmp(name, "skin_mode")) { OGRECU_REQUIRE(value >= 0 && value <= 3, "skin_mode is 0..3"); ctx->skin_mode = value; return OGRECU_OK; }
    if (!strcmp(name, "raster_xf")) { OGRECU_REQUIRE(value >= 0 && value <= 1, "raster_xf is 0..1"); ctx->raster_xf = value; return OGRECU_OK; }
    if (!strcmp(name, "raster_mode")) { OGRECU_REQUIRE(value >= 0 && value <= 1, "raster_mode is 0..1"); ctx->raster_mode = value; return OGRECU_OK; }
    if (!strcmp(name, "skel_mode")) { OGRECU_REQUIRE(value >= 0 && value <= 1, "skel_mode is 0..1"); ctx->skel_mode = value; return OGRECU_OK; }
    if (!strcmp(name, "optutil_sync")) { OGRECU_REQUIRE(value >= 0 && value <= 1, "optutil_sync is 0..1"); ctx->optutil_sync = value; return OGRECU_OK; }
    if (!strcmp(name, "optutil_batch")) {
        OGRECU_REQUIRE(value >= 0 && value <= 1, "optutil_batch is 0..1");
        if (!value) OGRECU_CHECK(flush_skin_batch(ctx));
        ctx->optutil_batch = value;
        return OGRECU_OK;
    }
    if (!strcmp(name, "lane_reuse_repack")) { OGRECU_REQUIRE(value >= 0 && value <= 1, "lane_reuse_repack is 0..1"); ctx->lane_reuse_repack = value; return OGRECU_OK; }
    if (!strcmp(name, "split_mode")) { OGRECU_REQUIRE(value >= 0 && value <= 1, "split_mode is 0..1"); ctx->split_mode = value; return OGRECU_OK; }
    if (!strcmp(name, "split_band_tiles")) { OGRECU_REQUIRE(value >= 1 && value <= 1024, "split_band_tiles is 1..1024"); ctx->split_band_tiles = value; return OGRECU_OK; }
    if (!strcmp(name, "skin_ipc")) { OGRECU_REQUIRE(value >= 0 && value <= 4096, "skin_ipc is 0..4096"); ctx->skin_ipc = value; return OGRECU_OK; }
    set_error("ogrecu_set_config_option: unknown option '%s'", name);
    return OGRECU_ERR_INVALID;
}

uint64_t ogrecu_launch_count(const ogrecu_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ogrecu_malloc(ogrecu_ctx* ctx, size_t bytes, void** dev_ptr)
{
    OGRECU_REQUIRE(ctx && dev_ptr, "NULL argument");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    *dev_ptr = nullptr;
    OGRECU_CUDA(cudaMalloc(dev_ptr, bytes ? bytes : 1));
    return OGRECU_OK;
}

int ogrecu_free(ogrecu_ctx* ctx, void* dev_ptr)
{
    OGRECU_REQUIRE(ctx, "ctx is NULL");
    if (!dev_ptr) return OGRECU_OK;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    OGRECU_CUDA(cudaFree(dev_ptr));
    return OGRECU_OK;
}

int ogrecu_host_alloc(size_t bytes, void** host_ptr)
{
    OGRECU_REQUIRE(host_ptr, "host_ptr is NULL");
    *host_ptr = nullptr;
    OGRECU_CUDA(cudaHostAlloc(host_ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return OGRECU_OK;
}

int ogrecu_host_free(void* host_ptr)
{
    if (host_ptr) OGRECU_CUDA(cudaFreeHost(host_ptr));
    return OGRECU_OK;
}

int ogrecu_memcpy_h2d(ogrecu_ctx* ctx, void* dev, const void* host, size_t bytes)
{
    OGRECU_REQUIRE(ctx && (bytes == 0 || (dev && host)), "NULL argument");
    if (!bytes) return OGRECU_OK;
    OGRECU_CHECK(flush_skin_batch(ctx));          // queued OptimisedUtil jobs come first (program order)
    OGRECU_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return OGRECU_OK;
}

int ogrecu_memcpy_h2d_on(ogrecu_ctx* ctx, void* cuda_stream, void* dev, const void* host, size_t bytes)
{
    OGRECU_REQUIRE(ctx && (bytes == 0 || (dev && host)), "NULL argument");
    if (!bytes) return OGRECU_OK;
    OGRECU_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)cuda_stream));
    return OGRECU_OK;
}

int ogrecu_memcpy_d2h(ogrecu_ctx* ctx, void* host, const void* dev, size_t bytes)
{
    OGRECU_REQUIRE(ctx && (bytes == 0 || (dev && host)), "NULL argument");
    if (!bytes) return OGRECU_OK;
    OGRECU_CHECK(flush_skin_batch(ctx));          // queued OptimisedUtil jobs come first (program order)
    OGRECU_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    return OGRECU_OK;
}

int ogrecu_memcpy_d2h_async(ogrecu_ctx* ctx, void* host, const void* dev, size_t bytes)
{
    OGRECU_REQUIRE(ctx && (bytes == 0 || (dev && host)), "NULL argument");
    if (!bytes) return OGRECU_OK;
    OGRECU_CHECK(flush_skin_batch(ctx));          // queued OptimisedUtil jobs come first (program order)
    OGRECU_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return OGRECU_OK;
}

int ogrecu_memcpy_d2d(ogrecu_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    OGRECU_REQUIRE(ctx && (bytes == 0 || (dst && src)), "NULL argument");
    if (!bytes) return OGRECU_OK;
    OGRECU_CHECK(flush_skin_batch(ctx));          // queued OptimisedUtil jobs come first (program order)
    OGRECU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return OGRECU_OK;
}

/* ------------------------------------------------------------ hardware buffers ----- */

int ogrecu_buf_create(ogrecu_ctx* ctx, size_t bytes, ogrecu_buf** out)
{
    OGRECU_REQUIRE(ctx && out, "NULL argument");
    *out = nullptr;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    ogrecu_buf* b = new ogrecu_buf();
    memset(b, 0, sizeof *b);
    b->ctx = ctx;
    b->size = bytes;
    cudaError_t e = cudaMalloc(&b->dev, bytes ? bytes : 1);
    if (e != cudaSuccess) { delete b; return cuda_fail(e, "cudaMalloc(buffer)", __FILE__, __LINE__); }
    // OGRE hands lock() pointers straight to OptimisedUtil and Tiny keeps them after
    // unlock(), so the shadow lives as long as the buffer.
    e = cudaHostAlloc((void**)&b->shadow, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaFree(b->dev); delete b; return cuda_fail(e, "cudaHostAlloc(shadow)", __FILE__, __LINE__); }
    e = cudaEventCreateWithFlags(&b->upload_ev, cudaEventDisableTiming);
    if (e != cudaSuccess) { cudaFree(b->dev); cudaFreeHost(b->shadow); delete b; return cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__); }
    for (size_t k = 0; k < ctx->peers.size(); ++k) {             // device group: a replica on every peer, filled on first use
        cudaSetDevice(ctx->peers[k]->device);
        e = cudaMalloc(&b->peer_dev[k], bytes ? bytes : 1);
        if (e != cudaSuccess) {
            for (size_t q = 0; q < k; ++q) { cudaSetDevice(ctx->peers[q]->device); cudaFree(b->peer_dev[q]); }
            cudaSetDevice(ctx->device);
            cudaFree(b->dev); cudaFreeHost(b->shadow); cudaEventDestroy(b->upload_ev); delete b;
            return cuda_fail(e, "cudaMalloc(buffer replica)", __FILE__, __LINE__);
        }
    }
    if (!ctx->peers.empty()) cudaSetDevice(ctx->device);
    { std::lock_guard<std::mutex> g(ctx->shadows_mutex); ctx->shadows[(uintptr_t)b->shadow] = b; ctx->dev_registry[(uintptr_t)b->dev] = b; }
    *out = b;
    return OGRECU_OK;
}

int ogrecu_buf_destroy(ogrecu_buf* b)
{
    if (!b) return OGRECU_OK;
    ogrecu_ctx* ctx = b->ctx;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    if (b->draw_ref == ctx->draw_epoch) flush_all_targets(ctx);          // recorded draws still read it
    if (b->batch_r == ctx->batch_epoch || b->batch_w == ctx->batch_epoch) flush_skin_batch(ctx);
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t k = 0; k < ctx->peers.size(); ++k) {
        cudaSetDevice(ctx->peers[k]->device);
        cudaStreamSynchronize(ctx->peers[k]->stream);
        cudaFree(b->peer_dev[k]);
    }
    if (!ctx->peers.empty()) cudaSetDevice(ctx->device);
    { std::lock_guard<std::mutex> g(ctx->shadows_mutex); ctx->shadows.erase((uintptr_t)b->shadow); ctx->dev_registry.erase((uintptr_t)b->dev); }
    cudaFree(b->dev);
    cudaFreeHost(b->shadow);
    cudaEventDestroy(b->upload_ev);
    delete b;
    return OGRECU_OK;
}

size_t ogrecu_buf_size(const ogrecu_buf* b) { return b ? b->size : 0; }

int ogrecu_buf_device_ptr(ogrecu_buf* b, void** dev_ptr)
{
    OGRECU_REQUIRE(b && dev_ptr, "NULL argument");
    *dev_ptr = b->dev;
    return OGRECU_OK;
}

int ogrecu_buf_mark_device_dirty(ogrecu_buf* b, size_t offset, size_t length)
{
    OGRECU_REQUIRE(b && offset <= b->size && length <= b->size - offset, "range out of bounds");
    if (!length) return OGRECU_OK;
    replicas_stale(b, offset, offset + length);
    if (!b->dev_dirty) { b->dev_dirty = true; b->dev_dirty_lo = offset; b->dev_dirty_hi = offset + length; }
    else {
        if (offset < b->dev_dirty_lo) b->dev_dirty_lo = offset;
        if (offset + length > b->dev_dirty_hi) b->dev_dirty_hi = offset + length;
    }
    return OGRECU_OK;
}

int ogrecu_buf_map(ogrecu_buf* b, size_t offset, size_t length, int mode, void** host_ptr)
{
    OGRECU_REQUIRE(b && host_ptr, "NULL argument");
    OGRECU_REQUIRE(!b->mapped, "buffer is already mapped");          // HardwareBuffer::lock asserts the same
    OGRECU_REQUIRE(offset <= b->size && length <= b->size - offset, "range out of bounds");
    ogrecu_ctx* ctx = b->ctx;
    // deferred work that the program order puts before this map: a host write must not reach a buffer that recorded draws or
    // queued jobs still read (or that a queued job would overwrite afterwards); a host read needs what queued jobs write
    if (mode & (OGRECU_MAP_WRITE | OGRECU_MAP_DISCARD)) {
        if (b->draw_ref == ctx->draw_epoch) OGRECU_CHECK(flush_all_targets(ctx));
        if (b->batch_r == ctx->batch_epoch || b->batch_w == ctx->batch_epoch) OGRECU_CHECK(flush_skin_batch(ctx));
        b->maxidx_cache.clear();
    } else if (b->batch_w == ctx->batch_epoch) {
        OGRECU_CHECK(flush_skin_batch(ctx));
    }
    if (b->upload_pending) {   // a previous unmap's upload still reads the shadow
        OGRECU_CUDA(cudaEventSynchronize(b->upload_ev));
        b->upload_pending = false;
    }
    // Bring device-side writes (skinning output) back before the host looks at or partially
    // overwrites the shadow.  Only a DISCARD map of the whole buffer skips the download; after
    // this block dev_dirty is always clear, so anything marked while the buffer is mapped is a
    // device write made DURING the map (OptimisedUtil writing into a locked destination).
    if (b->dev_dirty) {
        const bool whole_discard = (mode & OGRECU_MAP_DISCARD) && offset == 0 && length == b->size;
        if (!whole_discard) {
            OGRECU_CUDA(cudaMemcpyAsync(b->shadow + b->dev_dirty_lo, (uint8_t*)b->dev + b->dev_dirty_lo,
                                        b->dev_dirty_hi - b->dev_dirty_lo, cudaMemcpyDeviceToHost, ctx->stream));
            OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        b->dev_dirty = false;
    }
    b->mapped = true;
    b->map_mode = mode;
    b->map_off = offset;
    b->map_len = length;
    *host_ptr = b->shadow + offset;
    return OGRECU_OK;
}

int ogrecu_buf_unmap(ogrecu_buf* b)
{
    OGRECU_REQUIRE(b, "NULL argument");
    OGRECU_REQUIRE(b->mapped, "buffer is not mapped");
    ogrecu_ctx* ctx = b->ctx;
    b->mapped = false;
    if ((b->map_mode & (OGRECU_MAP_WRITE | OGRECU_MAP_DISCARD)) && b->map_len) {
        // pinned -> device, ordered on the stream before any later kernel.  A range written on the
        // device while the buffer was mapped (Mesh::softwareVertexBlend locks the destination, then
        // calls OptimisedUtil with the locked pointer, OgreMesh.cpp:1827-1896) holds the newer data:
        // it is left alone and stays device-dirty.
        size_t lo = b->map_off, hi = b->map_off + b->map_len;
        size_t piece[2][2] = {{lo, hi}, {0, 0}};
        if (b->dev_dirty && b->dev_dirty_lo < hi && b->dev_dirty_hi > lo) {
            piece[0][0] = lo; piece[0][1] = std::max(lo, std::min(hi, b->dev_dirty_lo));
            piece[1][0] = std::min(hi, std::max(lo, b->dev_dirty_hi)); piece[1][1] = hi;
        }
        bool any = false;
        for (auto& pc : piece)
            if (pc[1] > pc[0]) {
                OGRECU_CUDA(cudaMemcpyAsync((uint8_t*)b->dev + pc[0], b->shadow + pc[0], pc[1] - pc[0], cudaMemcpyHostToDevice, ctx->stream));
                replicas_stale(b, pc[0], pc[1]);
                any = true;
            }
        if (any) {
            OGRECU_CUDA(cudaEventRecord(b->upload_ev, ctx->stream));
            b->upload_pending = true;
        }
    }
    return OGRECU_OK;
}

}  // extern "C"

namespace ogrecu {
ogrecu_buf* buf_of_device_ptr(ogrecu_ctx* ctx, const void* dev_ptr)
{
    if (!dev_ptr) return nullptr;
    const uintptr_t p = (uintptr_t)dev_ptr;
    std::lock_guard<std::mutex> g(ctx->shadows_mutex);
    auto it = ctx->dev_registry.upper_bound(p);
    if (it == ctx->dev_registry.begin()) return nullptr;
    --it;
    ogrecu_buf* b = it->second;
    return (p < it->first + (b->size ? b->size : 1)) ? b : nullptr;
}
}  // namespace ogrecu

extern "C" {

int ogrecu_buf_resolve(ogrecu_ctx* ctx, const void* host_ptr, ogrecu_buf** out_buf, size_t* out_offset)
{
    OGRECU_REQUIRE(ctx && out_buf && out_offset, "NULL argument");
    *out_buf = nullptr;
    *out_offset = 0;
    uintptr_t p = (uintptr_t)host_ptr;
    std::lock_guard<std::mutex> g(ctx->shadows_mutex);
    auto it = ctx->shadows.upper_bound(p);
    if (it == ctx->shadows.begin()) return OGRECU_OK;
    --it;
    ogrecu_buf* b = it->second;
    if (p >= it->first && p < it->first + (b->size ? b->size : 1)) {
        *out_buf = b;
        *out_offset = (size_t)(p - it->first);
    }
    return OGRECU_OK;
}

}  // extern "C"
