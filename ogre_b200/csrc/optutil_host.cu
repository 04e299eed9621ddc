// optutil_host.cu — exact-signature host entry points of the OptimisedUtil back end.
//
// OGRE calls OptimisedUtil with raw host pointers obtained from HardwareBuffer::lock()
// (OgreMain/src/OgreMesh.cpp:1827-1896, OgreEdgeListBuilder.cpp:400-436, OgreShadowCaster.cpp:517-538).
// Pointers that fall inside an ogrecu_buf shadow are resolved to the device copy (no
// transfer; the destination is marked device-dirty).  Anything else is staged through
// pinned memory: H2D, kernel, D2H.  There is no CPU compute path.
#include "common.cuh"

#include <algorithm>
#include <cstring>

using namespace ogrecu;

namespace {

struct InReq { const uint8_t* host; size_t len; size_t dev_off; void* dev; bool resolved; };
struct OutReq {
    uint8_t* host; size_t len; size_t dev_off; void* dev; bool resolved;
    size_t piece, stride, count;   // copy-back pattern inside [host, host+len)
    bool preload;                  // extent also uploaded first (in-place / gaps preserved)
};

constexpr size_t ALIGN = 256;
inline size_t up(size_t v) { return (v + ALIGN - 1) & ~(ALIGN - 1); }

// Plans all transfers of one host call, then performs them with one H2D and one D2H.
class Stager {
public:
    explicit Stager(ogrecu_ctx* c) : ctx(c) {}
    size_t add_in(const void* host, size_t len) { ins.push_back({(const uint8_t*)host, len, 0, nullptr, false}); return ins.size() - 1; }
    size_t add_out(void* host, size_t len, size_t piece, size_t stride, size_t count, bool preload = false)
    {
        outs.push_back({(uint8_t*)host, len, 0, nullptr, false, piece, stride, count, preload});
        return outs.size() - 1;
    }
    void* in_dev(size_t k) const { return ins[k].dev; }
    void* out_dev(size_t k) const { return outs[k].dev; }

    int commit()
    {
        size_t in_total = 0, out_total = 0;
        for (auto& r : ins) {
            ogrecu_buf* b = nullptr; size_t off = 0;
            OGRECU_CHECK(ogrecu_buf_resolve(ctx, r.host, &b, &off));
            if (b && off + r.len <= b->size) { r.resolved = true; r.dev = (uint8_t*)b->dev + off; }
            else { r.dev_off = in_total; in_total += up(r.len); }
        }
        for (auto& r : outs) {
            ogrecu_buf* b = nullptr; size_t off = 0;
            OGRECU_CHECK(ogrecu_buf_resolve(ctx, r.host, &b, &off));
            if (b && off + r.len <= b->size) {
                r.resolved = true; r.dev = (uint8_t*)b->dev + off;
                OGRECU_CHECK(ogrecu_buf_mark_device_dirty(b, off, r.len));
            } else { r.dev_off = out_total; out_total += up(r.len); }
        }
        in_bytes = in_total; out_bytes = out_total;
        if (in_total + out_total == 0) return OGRECU_OK;          // everything already lives on the device: nothing to stage
        OGRECU_CHECK(scratch_reserve(ctx, ctx->stage_in, in_total + out_total + ALIGN));
        uint8_t* dev = (uint8_t*)ctx->stage_in.ptr;
        if (out_total == 0 && !ctx->optutil_sync && in_total <= ogrecu_ctx::RING_SLOT_BYTES) {
            // deferred completion: the inputs go through the next slot of the pinned ring; only that slot's previous copy is
            // waited for, never the stream
            if (!ctx->ring) OGRECU_CUDA(cudaHostAlloc((void**)&ctx->ring, ogrecu_ctx::RING_SLOTS * ogrecu_ctx::RING_SLOT_BYTES, cudaHostAllocDefault));
            const int slot = ctx->ring_next;
            ctx->ring_next = (slot + 1) % ogrecu_ctx::RING_SLOTS;
            if (!ctx->ring_ev[slot]) OGRECU_CUDA(cudaEventCreateWithFlags(&ctx->ring_ev[slot], cudaEventDisableTiming));
            if (ctx->ring_used[slot]) OGRECU_CUDA(cudaEventSynchronize(ctx->ring_ev[slot]));
            uint8_t* pin = ctx->ring + (size_t)slot * ogrecu_ctx::RING_SLOT_BYTES;
            for (auto& r : ins)
                if (!r.resolved) { memcpy(pin + r.dev_off, r.host, r.len); r.dev = dev + r.dev_off; }
            OGRECU_CUDA(cudaMemcpyAsync(dev, pin, in_total, cudaMemcpyHostToDevice, ctx->stream));
            OGRECU_CUDA(cudaEventRecord(ctx->ring_ev[slot], ctx->stream));
            ctx->ring_used[slot] = true;
            deferred = true;
            return OGRECU_OK;
        }
        OGRECU_CHECK(pinned_reserve(ctx, in_total + out_total + ALIGN));
        OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));   // pinned block free for reuse
        uint8_t* pin = (uint8_t*)ctx->pinned;
        for (auto& r : ins)
            if (!r.resolved) { memcpy(pin + r.dev_off, r.host, r.len); r.dev = dev + r.dev_off; }
        bool preload_any = false;
        for (auto& r : outs)
            if (!r.resolved) {
                r.dev = dev + in_total + r.dev_off;
                if (r.preload) { memcpy(pin + in_total + r.dev_off, r.host, r.len); preload_any = true; }
            }
        size_t h2d = in_total + (preload_any ? out_total : 0);
        if (h2d) OGRECU_CUDA(cudaMemcpyAsync(dev, pin, h2d, cudaMemcpyHostToDevice, ctx->stream));
        return OGRECU_OK;
    }

    int finish()
    {
        if (out_bytes) {
            uint8_t* pin = (uint8_t*)ctx->pinned + in_bytes;
            uint8_t* dev = (uint8_t*)ctx->stage_in.ptr + in_bytes;
            OGRECU_CUDA(cudaMemcpyAsync(pin, dev, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
            OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
            for (auto& r : outs) {
                if (r.resolved) continue;
                const uint8_t* s = pin + r.dev_off;
                if (r.piece == r.stride || r.count <= 1) memcpy(r.host, s, r.count ? (r.count - 1) * r.stride + r.piece : 0);
                else for (size_t i = 0; i < r.count; ++i) memcpy(r.host + i * r.stride, s + i * r.stride, r.piece);
            }
        } else if (ctx->optutil_sync) {
            OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));      // "optutil_sync" = 1: every call waits
        }
        // else: the results live in resolved device buffers, marked device-dirty: ogrecu_buf_map downloads them after a stream
        // sync when the host next locks them, and draws / later OptimisedUtil calls consume them in stream order
        return OGRECU_OK;
    }

private:
    ogrecu_ctx* ctx;
    std::vector<InReq> ins;
    std::vector<OutReq> outs;
    size_t in_bytes = 0, out_bytes = 0;
    bool deferred = false;
};

// [host, host + len) inside one ogrecu_buf's shadow -> the buffer and the device address, else NULL
static ogrecu_buf* resolve_range(ogrecu_ctx* ctx, const void* host, size_t len, const uint8_t** dev)
{
    ogrecu_buf* b = nullptr; size_t off = 0;
    if (ogrecu_buf_resolve(ctx, host, &b, &off) != OGRECU_OK || !b || off + len > b->size) return nullptr;
    *dev = (const uint8_t*)b->dev + off;
    return b;
}

// the same when the range probably lies in a buffer already found (normals beside positions, weights beside indices): one
// comparison instead of a locked map look-up
static ogrecu_buf* resolve_near(ogrecu_ctx* ctx, ogrecu_buf* hint, const void* host, size_t len, const uint8_t** dev)
{
    const uint8_t* p = (const uint8_t*)host;
    if (hint && p >= hint->shadow && p + len <= hint->shadow + hint->size) { *dev = (const uint8_t*)hint->dev + (p - hint->shadow); return hint; }
    return resolve_range(ctx, host, len, dev);
}

// highest blend index of the stream, remembered per buffer until the host next writes it (ogrecu_buf_map clears the cache)
static unsigned max_blend_index(ogrecu_buf* b, const uint8_t* blend_index, size_t stride, size_t nw, size_t n)
{
    const size_t off = b ? (size_t)(blend_index - b->shadow) : 0;
    if (b)
        for (const auto& c : b->maxidx_cache)
            if (c.off == off && c.stride == stride && c.nw == nw && c.n == n) return c.max_idx;
    unsigned max_idx = 0;
    for (size_t v = 0; v < n; ++v) {
        const uint8_t* bi = blend_index + v * stride;
        for (size_t k = 0; k < nw; ++k) max_idx = std::max<unsigned>(max_idx, bi[k]);
    }
    if (b) {
        if (b->maxidx_cache.size() >= 16) b->maxidx_cache.clear();
        b->maxidx_cache.push_back({off, stride, nw, n, max_idx});
    }
    return max_idx;
}

}  // namespace

namespace ogrecu {

// Run the queued softwareVertexSkinning jobs: descriptors, block prefix and palettes travel in ONE copy from a pinned
// arena (two arenas alternate, each guarded by an event: no stream synchronisation), one batched launch skins them all.
int flush_skin_batch(ogrecu_ctx* ctx)
{
    const uint32_t n = (uint32_t)ctx->batch_jobs.size();
    if (n == 0) return OGRECU_OK;
    ProfileScope range("ogrecu: batched OptimisedUtil flush");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    const size_t jobs_bytes = up((size_t)n * sizeof(ogrecu_skin_job)), prefix_bytes = up(((size_t)n + 1) * 4);
    const size_t pal_bytes = ctx->batch_pal_floats * 4, head = jobs_bytes + prefix_bytes, total = head + pal_bytes;
    const int s = ctx->batch_slot;
    ctx->batch_slot ^= 1;
    // (the arena's previous upload was waited for when the batch's first job was queued: its palettes already sit in batch_palpin[s])
    if (ctx->batch_pin_cap[s] < head) {
        if (ctx->batch_pin[s]) cudaFreeHost(ctx->batch_pin[s]);
        ctx->batch_pin[s] = nullptr; ctx->batch_pin_cap[s] = 0;
        const size_t cap = std::max(head * 2, (size_t)1 << 16);
        if (cudaHostAlloc((void**)&ctx->batch_pin[s], cap, cudaHostAllocDefault) != cudaSuccess) {
            (void)cudaGetLastError(); ctx->batch_jobs.clear(); ctx->batch_pal_floats = 0; ++ctx->batch_epoch;
            set_error("pinned allocation of %zu bytes failed", cap); return OGRECU_ERR_OOM;
        }
        ctx->batch_pin_cap[s] = cap;
    }
    if (!ctx->batch_ev[s]) OGRECU_CUDA(cudaEventCreateWithFlags(&ctx->batch_ev[s], cudaEventDisableTiming));
    int rc = scratch_reserve(ctx, ctx->batch_dev, total);
    uint64_t blocks = 0;
    if (rc == OGRECU_OK) {
        uint8_t* pin = ctx->batch_pin[s];
        uint8_t* dev = (uint8_t*)ctx->batch_dev.ptr;
        ogrecu_skin_job* hj = (ogrecu_skin_job*)pin;
        uint32_t* prefix = (uint32_t*)(pin + jobs_bytes);
        const float* dpal = (const float*)(dev + head);
        for (uint32_t k = 0; k < n; ++k) {
            hj[k] = ctx->batch_jobs[k];
            hj[k].palette = dpal + (size_t)(uintptr_t)ctx->batch_jobs[k].palette;            // queued as an offset in floats
            prefix[k] = (uint32_t)blocks;
            blocks += div_up(hj[k].num_vertices, 256);
        }
        prefix[n] = (uint32_t)blocks;
        cudaError_t e = cudaMemcpyAsync(dev, pin, head, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && pal_bytes) e = cudaMemcpyAsync(dev + head, ctx->batch_palpin[s], pal_bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->batch_ev[s], ctx->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "batch upload", __FILE__, __LINE__);
        else {
            ctx->batch_ev_used[s] = true;
            if (blocks) rc = launch_skin_batched(ctx, (const ogrecu_skin_job*)dev, (const uint32_t*)(dev + jobs_bytes), n, (unsigned)blocks);
        }
    }
    ctx->batch_flushes++; ctx->batch_jobs_run += n;
    ctx->batch_jobs.clear(); ctx->batch_pal_floats = 0;
    ++ctx->batch_epoch;                                           // every buffer's batch_r / batch_w mark is stale now
    return rc;
}

}  // namespace ogrecu

extern "C" {

int ogrecu_optutil_flush(ogrecu_ctx* ctx)
{
    OGRECU_REQUIRE(ctx, "ctx is NULL");
    return flush_skin_batch(ctx);
}

int ogrecu_optutil_batch_stats(ogrecu_ctx* ctx, uint64_t* flushes, uint64_t* jobs)
{
    OGRECU_REQUIRE(ctx && flushes && jobs, "NULL argument");
    *flushes = ctx->batch_flushes; *jobs = ctx->batch_jobs_run + ctx->batch_jobs.size();
    return OGRECU_OK;
}

int ogrecu_software_vertex_skinning_host(ogrecu_ctx* ctx, const float* src_pos, float* dst_pos, const float* src_norm,
                                         float* dst_norm, const float* blend_weight, const uint8_t* blend_index,
                                         const float* const* blend_matrices, size_t src_pos_stride, size_t dst_pos_stride,
                                         size_t src_norm_stride, size_t dst_norm_stride, size_t blend_weight_stride,
                                         size_t blend_index_stride, size_t nw, size_t n)
{
    OGRECU_REQUIRE(ctx, "ctx is NULL");
    if (n == 0) return OGRECU_OK;
    OGRECU_REQUIRE(src_pos && dst_pos && blend_weight && blend_index && blend_matrices, "NULL pointer");
    OGRECU_REQUIRE(nw >= 1 && nw <= 4, "num_weights_per_vertex must be 1..4");
    OGRECU_REQUIRE(!src_norm || dst_norm, "src_norm without dst_norm");
    OGRECU_REQUIRE(n <= 0xffffffffull, "too many vertices");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    const size_t q_sp_len = (n - 1) * src_pos_stride + 12, q_sn_len = src_norm ? (n - 1) * src_norm_stride + 12 : 0;
    const size_t q_w_len = (n - 1) * blend_weight_stride + 4 * nw, q_i_len = (n - 1) * blend_index_stride + nw;
    const size_t q_dp_len = (n - 1) * dst_pos_stride + 12, q_dn_len = src_norm ? (n - 1) * dst_norm_stride + 12 : 0;
    const uint8_t* d_bi = nullptr;
    ogrecu_buf* b_bi = resolve_range(ctx, blend_index, q_i_len, &d_bi);
    // highest palette slot referenced (the reference takes a bare pointer palette); scanned once per blend buffer
    const unsigned nmat = max_blend_index(b_bi, blend_index, blend_index_stride, nw, n) + 1;
    if (ctx->optutil_batch && !ctx->optutil_sync && b_bi) {
        // ---- deferred: every operand lives in an ogrecu_buf -> queue the job, copy the palette, return
        const uint8_t *d_sp = nullptr, *d_sn = nullptr, *d_w = nullptr, *d_dp = nullptr, *d_dn = nullptr;
        ogrecu_buf* b_sp = resolve_range(ctx, src_pos, q_sp_len, &d_sp);
        ogrecu_buf* b_sn = src_norm ? resolve_near(ctx, b_sp, src_norm, q_sn_len, &d_sn) : b_sp;
        ogrecu_buf* b_w = resolve_near(ctx, b_bi, blend_weight, q_w_len, &d_w);
        ogrecu_buf* b_dp = resolve_range(ctx, dst_pos, q_dp_len, &d_dp);
        ogrecu_buf* b_dn = src_norm ? resolve_near(ctx, b_dp, dst_norm, q_dn_len, &d_dn) : b_dp;
        if (b_sp && b_sn && b_w && b_dp && b_dn) {
            ogrecu_skin_job j;
            memset(&j, 0, sizeof j);
            j.src_pos = (const float*)d_sp; j.src_norm = (const float*)d_sn;
            j.dst_pos = (float*)d_dp; j.dst_norm = src_norm ? (float*)d_dn : nullptr;
            j.blend_weight = (const float*)d_w; j.blend_index = d_bi;
            j.src_pos_stride = (uint32_t)src_pos_stride; j.dst_pos_stride = (uint32_t)dst_pos_stride;
            j.src_norm_stride = (uint32_t)src_norm_stride; j.dst_norm_stride = (uint32_t)dst_norm_stride;
            j.blend_weight_stride = (uint32_t)blend_weight_stride; j.blend_index_stride = (uint32_t)blend_index_stride;
            j.num_weights = (uint32_t)nw; j.num_matrices = nmat; j.num_vertices = (uint32_t)n;
            j.palette = (const float*)(uintptr_t)64;                  // placeholder that passes the alignment check
            OGRECU_CHECK(validate_skin_job_public(j));
            // the jobs of one batch run concurrently in one launch, recorded draws run after the batch: whatever the program
            // order makes depend on an earlier queued job or an unflushed draw goes first
            if (b_dp->draw_ref == ctx->draw_epoch || b_dn->draw_ref == ctx->draw_epoch) OGRECU_CHECK(flush_all_targets(ctx));
            const uint64_t ep = ctx->batch_epoch;
            if (b_sp->batch_w == ep || b_sn->batch_w == ep || b_w->batch_w == ep || b_bi->batch_w == ep || b_dp->batch_w == ep ||
                b_dp->batch_r == ep || b_dn->batch_w == ep || b_dn->batch_r == ep)
                OGRECU_CHECK(flush_skin_batch(ctx));
            const uint64_t now = ctx->batch_epoch;
            b_sp->batch_r = b_sn->batch_r = b_w->batch_r = b_bi->batch_r = now;
            b_dp->batch_w = b_dn->batch_w = now;
            OGRECU_CHECK(ogrecu_buf_mark_device_dirty(b_dp, (size_t)(d_dp - (const uint8_t*)b_dp->dev), q_dp_len));
            if (src_norm) OGRECU_CHECK(ogrecu_buf_mark_device_dirty(b_dn, (size_t)(d_dn - (const uint8_t*)b_dn->dev), q_dn_len));
            // the palette is gathered straight into the pinned block the flush uploads from (a vector + one more copy of all the
            // frame's palettes at the flush cost 1.5 ms per 7 000 jobs)
            const int slot = ctx->batch_slot;
            if (ctx->batch_jobs.empty() && ctx->batch_ev_used[slot]) OGRECU_CUDA(cudaEventSynchronize(ctx->batch_ev[slot]));   // its last upload has been read
            const size_t pal_off = ctx->batch_pal_floats;            // multiple of 16 floats: keeps 16-byte alignment
            if (pal_off + (size_t)nmat * 16 > ctx->batch_pal_cap[slot]) {
                const size_t cap = std::max<size_t>(2 * (pal_off + (size_t)nmat * 16), (size_t)1 << 18);
                float* fresh = nullptr;
                if (cudaHostAlloc((void**)&fresh, cap * 4, cudaHostAllocDefault) != cudaSuccess) {
                    (void)cudaGetLastError();
                    set_error("pinned allocation of %zu bytes failed", cap * 4);
                    return OGRECU_ERR_OOM;
                }
                if (pal_off) memcpy(fresh, ctx->batch_palpin[slot], pal_off * 4);
                if (ctx->batch_palpin[slot]) cudaFreeHost(ctx->batch_palpin[slot]);
                ctx->batch_palpin[slot] = fresh; ctx->batch_pal_cap[slot] = cap;
            }
            float* pal_dst = ctx->batch_palpin[slot] + pal_off;
            for (unsigned m = 0; m < nmat; ++m) memcpy(pal_dst + (size_t)m * 16, blend_matrices[m], 64);
            ctx->batch_pal_floats = pal_off + (size_t)nmat * 16;
            j.palette = (const float*)(uintptr_t)pal_off;            // an offset in floats until the flush
            ctx->batch_jobs.push_back(j);
            if (ctx->batch_jobs.size() >= 65536) OGRECU_CHECK(flush_skin_batch(ctx));
            return OGRECU_OK;
        }
    }
    OGRECU_CHECK(flush_skin_batch(ctx));          // immediate path: queued jobs come first (program order)
    std::vector<float> palette((size_t)nmat * 16);
    for (unsigned m = 0; m < nmat; ++m) memcpy(&palette[(size_t)m * 16], blend_matrices[m], 64);

    Stager st(ctx);
    // source vertex stream(s): positions and normals interleaved in one buffer are one range
    const uint8_t* sp = (const uint8_t*)src_pos;
    const uint8_t* sn = (const uint8_t*)src_norm;
    size_t sp_len = (n - 1) * src_pos_stride + 12;
    size_t sn_len = src_norm ? (n - 1) * src_norm_stride + 12 : 0;
    size_t k_pos, k_norm = 0;
    bool shared_src = src_norm && !(sn + sn_len <= sp || sp + sp_len <= sn);
    const uint8_t* s_lo = sp; size_t s_len = sp_len;
    if (shared_src) { s_lo = std::min(sp, sn); s_len = (size_t)(std::max(sp + sp_len, sn + sn_len) - s_lo); }
    k_pos = st.add_in(s_lo, s_len);
    if (src_norm && !shared_src) k_norm = st.add_in(sn, sn_len);
    // blend stream(s)
    const uint8_t* wp = (const uint8_t*)blend_weight;
    size_t w_len = (n - 1) * blend_weight_stride + 4 * nw;
    size_t i_len = (n - 1) * blend_index_stride + nw;
    bool shared_blend = !(blend_index + i_len <= wp || wp + w_len <= blend_index);
    const uint8_t* b_lo = wp; size_t b_len = w_len;
    if (shared_blend) { b_lo = std::min(wp, blend_index); b_len = (size_t)(std::max(wp + w_len, blend_index + i_len) - b_lo); }
    size_t k_w = st.add_in(b_lo, b_len), k_i = 0;
    if (!shared_blend) k_i = st.add_in(blend_index, i_len);
    size_t k_pal = st.add_in(palette.data(), palette.size() * 4);
    // destination(s)
    uint8_t* dp = (uint8_t*)dst_pos;
    uint8_t* dn = (uint8_t*)dst_norm;
    size_t dp_len = (n - 1) * dst_pos_stride + 12;
    size_t dn_len = dst_norm && src_norm ? (n - 1) * dst_norm_stride + 12 : 0;
    bool shared_dst = src_norm && !(dn + dn_len <= dp || dp + dp_len <= dn);
    size_t k_dp, k_dn = 0;
    if (shared_dst) {
        uint8_t* d_lo = std::min(dp, dn);
        size_t d_len = (size_t)(std::max(dp + dp_len, dn + dn_len) - d_lo);
        bool packed = dst_pos_stride == 24 && dst_norm_stride == 24 && dn == dp + 12;
        // packed pos+normal: every byte of the extent is produced; otherwise keep the gaps
        k_dp = packed ? st.add_out(d_lo, d_len, 24, 24, n) : st.add_out(d_lo, d_len, d_len, d_len, 1, /*preload=*/true);
    } else {
        k_dp = st.add_out(dp, dp_len, 12, dst_pos_stride, n);
        if (src_norm) k_dn = st.add_out(dn, dn_len, 12, dst_norm_stride, n);
    }
    OGRECU_CHECK(st.commit());

    ogrecu_skin_job j;
    memset(&j, 0, sizeof j);
    const uint8_t* dsrc = (const uint8_t*)st.in_dev(k_pos);
    j.src_pos = (const float*)(dsrc + (sp - s_lo));
    if (src_norm) j.src_norm = shared_src ? (const float*)(dsrc + (sn - s_lo)) : (const float*)st.in_dev(k_norm);
    const uint8_t* dbl = (const uint8_t*)st.in_dev(k_w);
    j.blend_weight = (const float*)(dbl + (wp - b_lo));
    j.blend_index = shared_blend ? dbl + (blend_index - b_lo) : (const uint8_t*)st.in_dev(k_i);
    j.palette = (const float*)st.in_dev(k_pal);
    uint8_t* ddst = (uint8_t*)st.out_dev(k_dp);
    if (shared_dst) {
        uint8_t* d_lo = std::min(dp, dn);
        j.dst_pos = (float*)(ddst + (dp - d_lo));
        j.dst_norm = (float*)(ddst + (dn - d_lo));
    } else {
        j.dst_pos = (float*)ddst;
        if (src_norm) j.dst_norm = (float*)st.out_dev(k_dn);
    }
    j.src_pos_stride = (uint32_t)src_pos_stride; j.dst_pos_stride = (uint32_t)dst_pos_stride;
    j.src_norm_stride = (uint32_t)src_norm_stride; j.dst_norm_stride = (uint32_t)dst_norm_stride;
    j.blend_weight_stride = (uint32_t)blend_weight_stride; j.blend_index_stride = (uint32_t)blend_index_stride;
    j.num_weights = (uint32_t)nw; j.num_matrices = nmat; j.num_vertices = (uint32_t)n;
    OGRECU_CHECK(ogrecu_skin_instanced_dev(ctx, &j, 1, 0, 0, 0));
    return st.finish();
}

int ogrecu_software_vertex_morph_host(ogrecu_ctx* ctx, float t, const float* src1, const float* src2, float* dst,
                                      size_t stride1, size_t stride2, size_t dst_stride, size_t n, int morph_normals)
{
    OGRECU_REQUIRE(ctx, "ctx is NULL");
    if (n == 0) return OGRECU_OK;
    OGRECU_REQUIRE(src1 && src2 && dst, "NULL pointer");
    const size_t vs = morph_normals ? 24 : 12;
    OGRECU_REQUIRE(stride1 >= vs && stride2 >= vs && dst_stride >= vs, "stride smaller than the vertex");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    OGRECU_CHECK(flush_skin_batch(ctx));          // may read what a queued job writes
    Stager st(ctx);
    size_t k1 = st.add_in(src1, (n - 1) * stride1 + vs);
    size_t k2 = st.add_in(src2, (n - 1) * stride2 + vs);
    size_t kd = st.add_out(dst, (n - 1) * dst_stride + vs, vs, dst_stride, n);
    OGRECU_CHECK(st.commit());
    OGRECU_CHECK(ogrecu_morph_dev(ctx, t, (const float*)st.in_dev(k1), (const float*)st.in_dev(k2), (float*)st.out_dev(kd),
                                  stride1, stride2, dst_stride, n, morph_normals));
    return st.finish();
}

int ogrecu_concatenate_affine_matrices_host(ogrecu_ctx* ctx, const float* base16, const float* src, float* dst, size_t n)
{
    OGRECU_REQUIRE(ctx && base16, "NULL argument");
    if (n == 0) return OGRECU_OK;
    OGRECU_REQUIRE(src && dst, "NULL pointer");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    OGRECU_CHECK(flush_skin_batch(ctx));          // may read what a queued job writes
    Stager st(ctx);
    size_t ks = st.add_in(src, n * 64);
    size_t kd = st.add_out(dst, n * 64, 64, 64, n);
    OGRECU_CHECK(st.commit());
    OGRECU_CHECK(ogrecu_concat_affine_dev(ctx, base16, (const float*)st.in_dev(ks), (float*)st.out_dev(kd), n));
    return st.finish();
}

int ogrecu_calculate_face_normals_host(ogrecu_ctx* ctx, const float* positions, size_t num_vertices, const void* triangles,
                                       float* face_normals, size_t n)
{
    OGRECU_REQUIRE(ctx, "ctx is NULL");
    if (n == 0) return OGRECU_OK;
    OGRECU_REQUIRE(positions && triangles && face_normals, "NULL pointer");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    OGRECU_CHECK(flush_skin_batch(ctx));          // may read what a queued job writes
    if (num_vertices == 0) {   // the reference signature carries no vertex count
        ogrecu_buf* pb = nullptr;
        size_t poff = 0;
        OGRECU_CHECK(ogrecu_buf_resolve(ctx, positions, &pb, &poff));
        if (pb) num_vertices = (pb->size - poff) / 12;          // resident positions are used in place: nothing to size, nothing to scan
    }
    if (num_vertices == 0) {   // positions in plain host memory have to be staged: their extent is the highest index referenced
        const uint8_t* t = (const uint8_t*)triangles;
        uint32_t mx = 0;
        for (size_t i = 0; i < n; ++i) {
            uint32_t vi[3];
            memcpy(vi, t + i * 32 + 8, 12);
            mx = std::max(mx, std::max(vi[0], std::max(vi[1], vi[2])));
        }
        num_vertices = (size_t)mx + 1;
    }
    Stager st(ctx);
    size_t kp = st.add_in(positions, num_vertices * 12);
    size_t kt = st.add_in(triangles, n * 32);
    size_t kd = st.add_out(face_normals, n * 16, 16, 16, n);
    OGRECU_CHECK(st.commit());
    OGRECU_CHECK(ogrecu_face_normals_dev(ctx, (const float*)st.in_dev(kp), st.in_dev(kt), (float*)st.out_dev(kd), n));
    return st.finish();
}

int ogrecu_calculate_light_facing_host(ogrecu_ctx* ctx, const float* light4, const float* face_normals, char* facings, size_t n)
{
    OGRECU_REQUIRE(ctx && light4, "NULL argument");
    if (n == 0) return OGRECU_OK;
    OGRECU_REQUIRE(face_normals && facings, "NULL pointer");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    OGRECU_CHECK(flush_skin_batch(ctx));          // may read what a queued job writes
    Stager st(ctx);
    size_t kn = st.add_in(face_normals, n * 16);
    size_t kd = st.add_out(facings, n, 1, 1, n);
    OGRECU_CHECK(st.commit());
    OGRECU_CHECK(ogrecu_light_facing_dev(ctx, light4, (const float*)st.in_dev(kn), (char*)st.out_dev(kd), n));
    return st.finish();
}

int ogrecu_extrude_vertices_host(ogrecu_ctx* ctx, const float* light4, float dist, const float* src, float* dst, size_t n)
{
    OGRECU_REQUIRE(ctx && light4, "NULL argument");
    if (n == 0) return OGRECU_OK;
    OGRECU_REQUIRE(src && dst, "NULL pointer");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    OGRECU_CHECK(flush_skin_batch(ctx));          // may read what a queued job writes
    Stager st(ctx);
    size_t ks = st.add_in(src, n * 12);
    size_t kd = st.add_out(dst, n * 12, 12, 12, n);
    OGRECU_CHECK(st.commit());
    OGRECU_CHECK(ogrecu_extrude_dev(ctx, light4, dist, (const float*)st.in_dev(ks), (float*)st.out_dev(kd), n));
    return st.finish();
}

}  // extern "C"
