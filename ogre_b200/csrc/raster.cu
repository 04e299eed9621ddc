// raster.cu — binned, tile-based triangle rasteriser reproducing Tiny's semantics bit for bit.
//
// Reference path replaced (OGRE 14.6 tree):
//   TinyRenderSystem::_render + DefaultShader::vertex/fragment  RenderSystems/Tiny/src/OgreTinyRenderSystem.cpp:299-333,349-414
//   triangle() / barycentric()                                  RenderSystems/Tiny/src/tinyrenderer.h:29-95
//   IShader::sample2D                                           RenderSystems/Tiny/include/OgreTinyRenderSystem.h:34-42
//   clearFrameBuffer, readback                                  OgreTinyRenderSystem.cpp:421-433, OgreTinyWindow.cpp:68-80
//
// Pipeline per flush (all on the context stream):
//   K0 vertex_xf       1 thread / vertex: vertex shader, viewport, 1/w -> float4 (once per flush, not per index)
//   K1 setup_count / setup_scan / setup_write   1 thread / triangle, 1024 per CTA: cull, bbox, tile coverage;
//                      survivors get a compact 64-byte record (varyings are fetched at shade time) and their
//                      (tile, triangle) pairs, both in submission order (spin-free count -> scan -> write)
//      radix sort      stable LSD sort of the pairs by tile key -> per-tile lists in submission order
//   K3 raster_kernel   1 CTA / 16x16 tile, 1 pixel / thread; pixel depth+colour live in registers
//                      while the tile's triangle list is replayed IN ORDER (bit-exact blend,
//                      equal-depth and NaN semantics); colour staged in shared memory and
//                      flushed with 16-byte stores
// Every float op is an explicit __f*_rn intrinsic in the reference's association (SURVEY
// Appendix A): no FMA contraction, IEEE division.
#include "common.cuh"
#include <limits>

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cstring>

using namespace ogrecu;

namespace {

constexpr int TILE = 16;            // tile edge in pixels
constexpr int TILE_SHIFT = 4;
constexpr int RT_THREADS = 256;     // raster CTA = one thread per tile pixel
constexpr int CHUNK = 256;          // triangles staged in shared memory per round

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
// libstdc++ std::min/std::max (NaN falls through to the first argument)
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }
// x86 cvttss2si: NaN / out of range -> INT_MIN ("integer indefinite")
__device__ __forceinline__ int x86_f2i(float f)
{
    if (!(f < 2147483648.0f) || f < -2147483648.0f) return INT_MIN;
    return (int)f;
}

// Per-draw state on the device (one TinyRenderSystem::_render call)
// Ogre::CompareFunction (OgreCommon.h:61-71).  CMPF_LESS_EQUAL keeps Tiny's own form !(new > old) (tinyrenderer.h:79).
__device__ __forceinline__ bool cmp_depth(uint32_t func, float a, float b)
{
    switch (func) {
    case OGRECU_CMPF_ALWAYS_FAIL: return false;
    case OGRECU_CMPF_ALWAYS_PASS: return true;
    case OGRECU_CMPF_LESS: return a < b;
    case OGRECU_CMPF_LESS_EQUAL: return !(a > b);
    case OGRECU_CMPF_EQUAL: return a == b;
    case OGRECU_CMPF_NOT_EQUAL: return a != b;
    case OGRECU_CMPF_GREATER_EQUAL: return a >= b;
    default: return a > b;
    }
}
__device__ __forceinline__ bool cmp_alpha(uint32_t func, uint32_t a, uint32_t b)
{
    switch (func) {
    case OGRECU_CMPF_ALWAYS_FAIL: return false;
    case OGRECU_CMPF_ALWAYS_PASS: return true;
    case OGRECU_CMPF_LESS: return a < b;
    case OGRECU_CMPF_LESS_EQUAL: return a <= b;
    case OGRECU_CMPF_EQUAL: return a == b;
    case OGRECU_CMPF_NOT_EQUAL: return a != b;
    case OGRECU_CMPF_GREATER_EQUAL: return a >= b;
    default: return a > b;
    }
}

struct DrawDev {
    const uint8_t* pos; const uint8_t* uv; const uint8_t* nrm; const void* idx;
    const uint8_t* tex;
    uint32_t pos_stride, uv_stride, nrm_stride, index_type;
    uint32_t is_strip, flags, target, tri_base;
    uint32_t tri_count, tex_w, tex_h, tex_zero_alpha;
    const float4* xf;         // optional: pre-transformed vertices (X, Y, Z, 1/w) of this draw, [vertex_count]
    uint32_t vertex_count, vert_base;
    // "beyond Tiny" state (ogrecu_draw_state), all zero for a plain Tiny draw:
    //   ext = enable bits | depth_func << 8 | cull_mode << 12 | alpha_func << 16 | alpha_ref << 24
    uint32_t ext, col_stride;
    const uint8_t* col;        // VES_DIFFUSE bytes r,g,b,a
    uint32_t sc_x, sc_y;       // scissor: left | (right - 1) << 16, top | (bottom - 1) << 16 (inclusive pixel bounds)
    uint32_t inst_tris, inst_verts;   // instanced list: triangles per instance (0 = plain draw), vertex index step per instance
    uint32_t inst_magic, pad_;        // min(floor(2^32 / inst_tris), 2^32 - 1): triangle -> instance without a hardware divide
    float mvp[16], texm[16], vp[16];
    float mvit[9];
    float light[3];
    float ambient[4];
};

struct TargetDev {
    uint8_t* colour;
    float* depth;
};

// Per-triangle record written by setup (4 x float4 = 64 bytes), COMPACT: only triangles that
// survive cull / clip and touch a tile of this strip get one, in submission order.
struct __align__(16) TriRec {
    float x[3], y[3];   // screen X, Y per vertex (after the divide)
    float iw[3];        // 1/w per vertex
    float z[3];         // screen z per vertex
    uint32_t bx;        // x0 | x1 << 16   (inclusive pixel bbox)
    uint32_t by;        // y0 | y1 << 16
    uint32_t draw;      // index into the DrawDev table
    uint32_t tri;       // triangle index inside its draw (varyings are fetched from the source streams at shade time)
};
static_assert(sizeof(TriRec) == 64, "TriRec must be 4 float4");

struct SegParams {
    const DrawDev* draws;
    uint32_t ndraws;
    uint32_t ntris;
    TriRec* rec;              // [survivors]
    uint64_t* pairs;          // (tile key << 32 | compact triangle index), submission order
    uint64_t pair_capacity;
    uint32_t* block_agg;      // [nblocks][2]: pairs, survivors of each 1024-triangle CTA
    uint8_t* alive;           // [ntris]: tiles the triangle touches per the count pass (0 = dropped: culled, off screen, no covered pixel; 255 = 255 or more)
    unsigned long long* bases;// [nblocks][2]: their exclusive scans
    uint32_t* ctl;            // [2] total pairs  [3] total survivors
    uint32_t width, height;
    uint32_t tiles_x, tiles_y;        // tiles per target
    uint32_t strip_ty0, strip_ty1;    // tile rows rasterised on this context
    uint32_t band_tiles, band_phase, band_period;   // ... of which only bands with (ty / band_tiles) % period == phase (period 1: all)
    uint32_t* tile_flags;     // [ntargets*tiles_x*tiles_y]  bit 0: some triangle needs the in-order path
};

/* ------------------------------------------------------------------- K1: setup ------ */

__device__ __forceinline__ void mat4_mul(const float* __restrict__ m, float x, float y, float z, float w, float out[4])
{
    // Matrix4 * Vector4, OgreMatrix4.h:571-578: ((m0*x + m1*y) + m2*z) + m3*w
#pragma unroll
    for (int r = 0; r < 4; ++r)
        out[r] = fadd(fadd(fadd(fmul(m[4 * r + 0], x), fmul(m[4 * r + 1], y)), fmul(m[4 * r + 2], z)), fmul(m[4 * r + 3], w));
}

// DefaultShader::vertex position part + the head of triangle() for ONE vertex: clip = MVP*(p,1)
// (OgreTinyRenderSystem.cpp:302), q = Viewport*clip (tinyrenderer.h:45), pts /= w as "fInv = 1.0f/w; x*fInv"
// (OgreVector.h:563-570); pts[3] = 1/w (tinyrenderer.h:50) is the same correctly rounded quotient.
// Pure function of the vertex and the draw's matrices: evaluating it once per vertex instead of once per
// index (Tiny has no vertex cache) gives the same bits.
__device__ __forceinline__ float4 transform_vertex(const float* __restrict__ mvp, const float* __restrict__ vp, float px, float py, float pz)
{
    float clip[4], q[4];
    mat4_mul(mvp, px, py, pz, 1.0f, clip);
    mat4_mul(vp, clip[0], clip[1], clip[2], clip[3], q);
    const float fi = fdiv(1.0f, q[3]);
    return make_float4(fmul(q[0], fi), fmul(q[1], fi), fmul(q[2], fi), fi);
}

struct XfParams { const DrawDev* draws; uint32_t ndraws; uint32_t nverts; float4* out; };

// vertex pre-transform: one thread per (draw, vertex) over all draws of the flush that gave a vertex count
__global__ void __launch_bounds__(256) vertex_xf_kernel(XfParams P)
{
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= P.nverts) return;
    uint32_t lo = 0, hi = P.ndraws;                    // last draw with vert_base <= v (draws without xf have vertex_count 0)
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&P.draws[mid].vert_base) <= v) lo = mid; else hi = mid;
    }
    const DrawDev& D = P.draws[lo];
    const uint32_t vi = v - D.vert_base;
    const float* p = (const float*)(D.pos + (size_t)vi * D.pos_stride);
    float4 sv = transform_vertex(D.mvp, D.vp, __ldg(p), __ldg(p + 1), __ldg(p + 2));
    // a vertex with any non-finite screen coordinate drops its triangles (setup_triangle): mark it in X, the one field they test
    if (!(isfinite(sv.y) && isfinite(sv.z) && isfinite(sv.w))) sv.x = __int_as_float(0x7fc00000);
    P.out[v] = sv;
}

// triangle t of an instanced list -> (instance, triangle inside the instance).  q = umulhi(t, floor(2^32 / n)) is the quotient or
// one less (the estimate is low by t * frac(2^32 / n) / 2^32 < 1), so one correction step makes it exact.
__device__ __forceinline__ uint32_t instance_of(uint32_t& t, uint32_t inst_tris, uint32_t inst_magic)
{
    uint32_t q = __umulhi(t, inst_magic), r = t - q * inst_tris;
    if (r >= inst_tris) { ++q; r -= inst_tris; }
    t = r;
    return q;
}

// index of vertex j of triangle t (list: 3t+j, strip: t+j — sliding window without winding flip,
// OgreTinyRenderSystem.cpp:396-403); 16-bit indices are read unsigned (the reference reads int16: UB >= 32768)
__device__ __forceinline__ size_t vertex_index(const DrawDev& D, uint32_t t, int j)
{
    size_t base = 0;
    if (D.inst_tris) base = (size_t)instance_of(t, D.inst_tris, D.inst_magic) * D.inst_verts;   // instanced list: the same indices, shifted per instance
    const size_t k = (D.is_strip ? (size_t)t : (size_t)t * 3) + j;
    if (D.index_type == 2) return base + __ldg((const uint16_t*)D.idx + k);
    if (D.index_type == 4) return base + __ldg((const uint32_t*)D.idx + k);
    return base + k;
}

// The draw fields every triangle's set-up reads, fetched ONCE per triangle through the read-only path (the compiler re-read them
// from the draw table around every branch when they were plain loads: 330 of the count pass's 575 instructions per triangle-warp,
// three hardware divides among them; profiles/r02_notes.md).
struct DrawHot {
    const float4* xf; const void* idx; const uint8_t* pos;
    uint32_t nxf, pos_stride, index_type, is_strip, ext, tri_base, inst_tris, inst_verts, inst_magic;
};
__device__ __forceinline__ DrawHot load_hot(const DrawDev* D)
{
    DrawHot H;
    H.xf = (const float4*)__ldg((const unsigned long long*)&D->xf);
    H.idx = (const void*)__ldg((const unsigned long long*)&D->idx);
    H.pos = (const uint8_t*)__ldg((const unsigned long long*)&D->pos);
    H.nxf = __ldg(&D->vertex_count); H.pos_stride = __ldg(&D->pos_stride); H.index_type = __ldg(&D->index_type);
    H.is_strip = __ldg(&D->is_strip); H.ext = __ldg(&D->ext); H.tri_base = __ldg(&D->tri_base);
    const uint2 in = __ldg((const uint2*)&D->inst_tris);
    H.inst_tris = in.x; H.inst_verts = in.y; H.inst_magic = __ldg(&D->inst_magic);
    return H;
}
// the three vertex indices of triangle t of the draw
__device__ __forceinline__ void triangle_vertices(const DrawHot& H, uint32_t t, size_t vi[3])
{
    size_t base = 0;
    if (H.inst_tris) base = (size_t)instance_of(t, H.inst_tris, H.inst_magic) * H.inst_verts;
    const size_t k = H.is_strip ? (size_t)t : (size_t)t * 3;
    if (H.index_type == 2) {
        const uint16_t* ix = (const uint16_t*)H.idx + k;
        vi[0] = base + __ldg(ix); vi[1] = base + __ldg(ix + 1); vi[2] = base + __ldg(ix + 2);
    } else if (H.index_type == 4) {
        const uint32_t* ix = (const uint32_t*)H.idx + k;
        vi[0] = base + __ldg(ix); vi[1] = base + __ldg(ix + 1); vi[2] = base + __ldg(ix + 2);
    } else { vi[0] = base + k; vi[1] = base + k + 1; vi[2] = base + k + 2; }
}

// CTAs per SM the set-up passes' registers are bounded for.  Both passes wait on gathers (indices -> transformed vertices) more than they
// compute; six CTAs (40 registers, 64 / ~100 bytes spilled) against the four the unbounded 64 registers gave: C4 0.0560 -> 0.0539 ms
// per view, C5 1.132 -> 1.079 ms, the crowd frame's render 2.69 -> 2.56 ms (profiles/r02i_setup_min_ctas.txt; 5 / 5: 0.0548, 1.103, 2.60).
#ifndef SC_MIN_CTAS
#define SC_MIN_CTAS 6
#endif
#ifndef SW_MIN_CTAS
#define SW_MIN_CTAS 6
#endif
constexpr int SB_THREADS = 256;
constexpr int SB_ROUNDS = 4;                         // 256-triangle rounds per CTA
constexpr int SB_BLOCK = SB_THREADS * SB_ROUNDS;     // triangles per CTA = per entry of the block-aggregate table

__device__ __forceinline__ bool owns_row(const SegParams& P, uint32_t ty)
{
    return P.band_period <= 1 || (ty / P.band_tiles) % P.band_period == P.band_phase;
}

#ifndef SETUP_COVER_MAX
#define SETUP_COVER_MAX 6        // triangles whose bounding box holds at most this many pixels are tested for coverage at set-up
#endif
__device__ __forceinline__ void expand_xy(float X0, float X1, float X2, float Y0, float Y1, float Y2, float hf[20]);
__device__ __forceinline__ bool eval_fragment(const float* hf, float Px, float Py, float& bc0, float& bc1, float& bc2, float& fd);
__device__ __forceinline__ bool covers_pixel(const float* hf, float Px, float Py);

struct TriSetup {
    float X[3], Y[3], Z[3], IW[3];
    int x0, x1, y0, y1;
    uint32_t draw, tri, ntiles;
    bool ordered;
};

// Everything Tiny does per triangle before its pixel loop (OgreTinyRenderSystem.cpp:396-410,
// tinyrenderer.h:45-69): fetch / transform the three vertices, cull, bounding box, plus the number of
// 16x16 tiles of this strip the box touches.  Pure function of (g, draw table): both passes below call it.
// a vertex of a draw that gave no vertex count (or an index past it): transformed here, per index, like Tiny.  Out of line and by
// value: the pre-transformed path (every draw OGRE submits) carries none of its address set-up and keeps its registers.
__device__ __noinline__ float4 fetch_untransformed(const DrawDev* D, const uint8_t* pos, uint32_t pos_stride, size_t vi, const float* s_mat)
{
    const float* Mmvp = s_mat ? s_mat : D->mvp;
    const float* Mvp = s_mat ? s_mat + 16 : D->vp;
    const float* p = (const float*)(pos + vi * pos_stride);
    return transform_vertex(Mmvp, Mvp, __ldg(p), __ldg(p + 1), __ldg(p + 2));
}

template <bool COVER>
__device__ __forceinline__ void setup_triangle(const SegParams& P, uint32_t g, uint32_t lo0, uint32_t lo0_end, const float* s_mat, TriSetup& T)
{
    T.ntiles = 0; T.ordered = false; T.x0 = 0; T.x1 = -1; T.y0 = 0; T.y1 = -1;
    // which draw: last d with tri_base <= g (almost always the CTA's first draw, whose matrices sit in shared memory)
    uint32_t lo = lo0;
    const bool same = g < lo0_end;
    if (!same) {
        uint32_t hi = P.ndraws;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&P.draws[mid].tri_base) <= g) lo = mid; else hi = mid;
        }
    }
    const DrawDev& D = P.draws[lo];
    const DrawHot H = load_hot(&D);
    const uint32_t t = g - H.tri_base;
    T.draw = lo; T.tri = t;
    bool ok = true;
    size_t vi[3];
    triangle_vertices(H, t, vi);
    // non-finite screen coordinates: the reference writes out of bounds (undefined) -> drop
    if (H.xf && vi[0] < H.nxf && vi[1] < H.nxf && vi[2] < H.nxf) {
        // transformed once per vertex (vertex_xf_kernel, which folds the four finite tests into X).  The three gathers are issued
        // back to back: behind a branch per vertex they made a chain of three L2 round trips (28 % of the count pass's stall samples).
        const float4 s0 = __ldg(H.xf + vi[0]), s1 = __ldg(H.xf + vi[1]), s2 = __ldg(H.xf + vi[2]);
        T.X[0] = s0.x; T.Y[0] = s0.y; T.Z[0] = s0.z; T.IW[0] = s0.w;
        T.X[1] = s1.x; T.Y[1] = s1.y; T.Z[1] = s1.z; T.IW[1] = s1.w;
        T.X[2] = s2.x; T.Y[2] = s2.y; T.Z[2] = s2.z; T.IW[2] = s2.w;
        ok = isfinite(s0.x) & isfinite(s1.x) & isfinite(s2.x);
    } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float4 sv = (H.xf && vi[j] < H.nxf) ? __ldg(H.xf + vi[j]) : fetch_untransformed(&D, H.pos, H.pos_stride, vi[j], same ? s_mat : nullptr);
            ok = ok && isfinite(sv.x) && isfinite(sv.y) && isfinite(sv.z) && isfinite(sv.w);
            T.X[j] = sv.x; T.Y[j] = sv.y; T.Z[j] = sv.z; T.IW[j] = sv.w;
        }
    }
    const uint32_t ext = H.ext;
    // Tiny: lists are culled by its fixed rule (= CULL_CLOCKWISE), strips never (.cpp:410); beyond Tiny: the draw's culling mode
    const uint32_t cull = (ext & OGRECU_EXT_CULL_MODE) ? ((ext >> 12) & 0xfu) : (H.is_strip ? (uint32_t)OGRECU_CULL_NONE : (uint32_t)OGRECU_CULL_CLOCKWISE);
    if (ok && cull != OGRECU_CULL_NONE) {                                      // cull: tinyrenderer.h:37-39,55
        float ax = fsub(T.X[2], T.X[0]), ay = fsub(T.Y[2], T.Y[0]), bx = fsub(T.X[2], T.X[1]), by = fsub(T.Y[2], T.Y[1]);
        float cr = fsub(fmul(ax, by), fmul(ay, bx));
        if ((ext & OGRECU_EXT_CULL_MODE) && H.is_strip && (t & 1u)) cr = -cr;  // a strip's odd triangles wind the other way
        if (cull == OGRECU_CULL_CLOCKWISE ? cr > 0.0f : cr < 0.0f) ok = false;
    }
    if (ok) {
        // tinyrenderer.h:58-65 clamps as it goes: mn = max(0, min(mn, X[j])), mx = min(c, max(mx, X[j])) from +-FLT_MAX.  For the
        // finite coordinates that reach this point min / max are lattice operations, so that fold equals max(0, min(X0, X1, X2)) and
        // min(c, max(X0, X1, X2)) (it may give -0 where the fold gives +0: both truncate to pixel 0).
        const float cx = (float)(P.width - 1u), cy = (float)(P.height - 1u);
        const float mnx = fmaxf(0.f, fminf(fminf(T.X[0], T.X[1]), T.X[2])), mxx = fminf(cx, fmaxf(fmaxf(T.X[0], T.X[1]), T.X[2]));
        const float mny = fmaxf(0.f, fminf(fminf(T.Y[0], T.Y[1]), T.Y[2])), mxy = fminf(cy, fmaxf(fmaxf(T.Y[0], T.Y[1]), T.Y[2]));
        // beyond +-2^31 the reference's (int) loop bounds are undefined -> drop
        if (!(mnx < 2147483648.0f) || !(mny < 2147483648.0f) || !(mxx > -2147483648.0f) || !(mxy > -2147483648.0f)) ok = false;
        else {
            T.x0 = (int)mnx; T.x1 = (int)mxx; T.y0 = (int)mny; T.y1 = (int)mxy;    // :68-69 truncation toward zero
            if (ext & OGRECU_EXT_SCISSOR) {                                        // beyond Tiny: the bbox is the pixel loop
                T.x0 = max(T.x0, (int)(D.sc_x & 0xffffu)); T.x1 = min(T.x1, (int)(D.sc_x >> 16));
                T.y0 = max(T.y0, (int)(D.sc_y & 0xffffu)); T.y1 = min(T.y1, (int)(D.sc_y >> 16));
            }
            if (T.x0 > T.x1 || T.y0 > T.y1) ok = false;
        }
    }
#if SETUP_COVER_MAX > 0
    // Early coverage cull.  In scenes of sub-pixel triangles (config 5: 0.4 pixels per triangle, the crowd of config 2: 40
    // bounding boxes per pixel) most triangles contain no pixel CENTRE at all: Tiny walks their bounding box and rejects
    // every pixel (tinyrenderer.h:70-76).  For a box of a few pixels that walk is done here, with the very operations the
    // raster kernel would use (expand_xy + eval_fragment): a triangle none of whose pixels passes the inside / depth >= 0
    // tests can produce no fragment whatever the target holds, so it gets no record, no (tile, triangle) pair, no sort key and
    // no visit by the tile's CTA.  A NaN anywhere passes the tests as it does in Tiny (comparisons with NaN are false) and
    // keeps the triangle.
    if (COVER && ok && (T.x1 - T.x0 + 1) * (T.y1 - T.y0 + 1) <= SETUP_COVER_MAX) {
        float hf[20];
        expand_xy(T.X[0], T.X[1], T.X[2], T.Y[0], T.Y[1], T.Y[2], hf);
        hf[9] = T.IW[0]; hf[10] = T.IW[1]; hf[11] = T.IW[2];
        hf[12] = T.Z[0]; hf[13] = T.Z[1]; hf[14] = T.Z[2];
        bool any = false;
        int px = T.x0, py = T.y0;
#pragma unroll 1
        for (int n = (T.x1 - T.x0 + 1) * (T.y1 - T.y0 + 1); n > 0 && !any; --n) {       // one flat loop over the box's few pixels
            any = covers_pixel(hf, (float)px, (float)py);
            if (++px > T.x1) { px = T.x0; ++py; }
        }
        ok = any;
    }
#endif
    if (ok) {
        int tx0 = T.x0 >> TILE_SHIFT, tx1 = T.x1 >> TILE_SHIFT;
        int ty0 = max(T.y0 >> TILE_SHIFT, (int)P.strip_ty0), ty1 = min(T.y1 >> TILE_SHIFT, (int)P.strip_ty1 - 1);
        if (ty0 <= ty1) {
            uint32_t rows = (uint32_t)(ty1 - ty0 + 1);
            if (P.band_period > 1) {                                           // interleaved bands: count the rows owned here
                rows = 0;
                for (int ty = ty0; ty <= ty1; ++ty) rows += owns_row(P, (uint32_t)ty) ? 1u : 0u;
            }
            T.ntiles = (uint32_t)(tx1 - tx0 + 1) * rows;
        }
    }
    if (T.ntiles) {
        // Order-independent resolution (raster_kernel's visibility path) is only equivalent to Tiny's
        // in-order loop for opaque, depth-tested, depth-writing fragments that cannot be discarded.
        const uint32_t fl = D.flags;
        T.ordered = (fl & (OGRECU_DRAW_DEPTH_TEST | OGRECU_DRAW_DEPTH_WRITE | OGRECU_DRAW_BLEND_ADD)) !=
                        (OGRECU_DRAW_DEPTH_TEST | OGRECU_DRAW_DEPTH_WRITE) ||
                    (D.tex != nullptr && D.tex_zero_alpha != 0) ||
                    // beyond Tiny: another depth function or an alpha test make the outcome order dependent
                    ((ext & OGRECU_EXT_DEPTH_FUNC) && ((ext >> 8) & 0xfu) != OGRECU_CMPF_LESS_EQUAL) ||
                    ((ext & OGRECU_EXT_ALPHA_REJECT) && ((ext >> 16) & 0xfu) != OGRECU_CMPF_ALWAYS_PASS);
    }
}

// the draw of the CTA's first triangle: its matrices are cached in shared memory
__device__ __forceinline__ void cta_draw_cache(const SegParams& P, uint32_t g0, uint32_t* s_lo, float* s_mat, uint32_t& lo0, uint32_t& lo0_end)
{
    if (threadIdx.x == 0) {
        uint32_t lo = 0, hi = P.ndraws;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&P.draws[mid].tri_base) <= g0) lo = mid; else hi = mid;
        }
        *s_lo = lo;
    }
    __syncthreads();
    lo0 = *s_lo;
    if (threadIdx.x < 32) s_mat[threadIdx.x] = threadIdx.x < 16 ? P.draws[lo0].mvp[threadIdx.x] : P.draws[lo0].vp[threadIdx.x - 16];
    lo0_end = lo0 + 1 < P.ndraws ? __ldg(&P.draws[lo0 + 1].tri_base) : 0xffffffffu;   // triangles below this belong to draw lo0
    __syncthreads();
}

// K1a: per CTA of 1024 triangles, how many survive and how many (tile, triangle) pairs they make.
// (A spin-free "reduce, scan the block table, then write" compaction: the single-pass chained scan
//  measured 30 % of its time waiting in the look-back, profiles/r01d notes.  Measured and rejected in r01k: issuing
//  the index loads of all four rounds before the first vertex gather — setup_write goes from 64 to 80 registers,
//  one CTA less per SM, C5 2.16 -> 2.25 ms; staging the CTA's 12 KB of indices through shared memory with one coalesced
//  copy — same registers, C5 2.15 -> 2.30 ms, C4 8.8 -> 8.4 k views/s.  The index round trip is not what bounds these kernels.)
__global__ void __launch_bounds__(SB_THREADS, SC_MIN_CTAS) setup_count_kernel(SegParams P)
{
    __shared__ uint32_t s_lo, s_wp[8], s_wt[8];
    __shared__ float s_mat[32];
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t g0 = blockIdx.x * (uint32_t)SB_BLOCK;
    uint32_t lo0, lo0_end;
    cta_draw_cache(P, g0, &s_lo, s_mat, lo0, lo0_end);
    uint32_t np = 0, nt = 0;
#pragma unroll 2
    for (int k = 0; k < SB_ROUNDS; ++k) {
        const uint32_t g = g0 + (uint32_t)k * SB_THREADS + tid;
        if (g < P.ntris) {
            TriSetup T;
            setup_triangle<true>(P, g, lo0, lo0_end, s_mat, T);
            np += T.ntiles; nt += T.ntiles ? 1u : 0u;
            P.alive[g] = (uint8_t)min(T.ntiles, 255u);     // tiles of the triangle (0 = dropped; 255 = that many or more)
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { np += __shfl_xor_sync(0xffffffffu, np, o); nt += __shfl_xor_sync(0xffffffffu, nt, o); }
    if (lane == 0) { s_wp[wid] = np; s_wt[wid] = nt; }
    __syncthreads();
    if (tid == 0) {
        uint32_t a = 0, b = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a += s_wp[w]; b += s_wt[w]; }
        P.block_agg[2 * blockIdx.x] = a; P.block_agg[2 * blockIdx.x + 1] = b;
    }
}

// K1b: exclusive scan of the block table (64-bit running sums), totals -> ctl[2..3].  One CTA per 8192 table entries; each first
// sums every entry before its own (a coalesced read of at most the whole table out of L2: the table has one entry per 1024
// triangles, so this stays far below the count pass even when a thousand CTAs repeat it), then scans its own.  No CTA waits for
// another.
constexpr int SS_ITEMS = 8;
constexpr int SS_CHUNK = 1024 * SS_ITEMS;
__global__ void __launch_bounds__(1024) setup_scan_kernel(const uint32_t* __restrict__ block_agg, uint32_t nblocks, uint32_t* ctl, unsigned long long* bases)
{
    __shared__ unsigned long long s_wp[32], s_wt[32];
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t b0 = blockIdx.x * (uint32_t)SS_CHUNK;
    const uint2* agg = reinterpret_cast<const uint2*>(block_agg);
    // ---- carry: everything before this chunk
    unsigned long long cp = 0, ct = 0;
    for (uint32_t i = tid; i < b0; i += 1024) { const uint2 v = __ldg(agg + i); cp += v.x; ct += v.y; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { cp += __shfl_xor_sync(0xffffffffu, cp, o); ct += __shfl_xor_sync(0xffffffffu, ct, o); }
    if (lane == 0) { s_wp[wid] = cp; s_wt[wid] = ct; }
    __syncthreads();
    cp = 0; ct = 0;
    for (int w = 0; w < 32; ++w) { cp += s_wp[w]; ct += s_wt[w]; }
    __syncthreads();
    // ---- this chunk
    const uint32_t b = b0 + tid * SS_ITEMS;
    uint32_t vp[SS_ITEMS], vt[SS_ITEMS];
    unsigned long long sp = 0, st = 0;
#pragma unroll
    for (int i = 0; i < SS_ITEMS; ++i) {
        const uint2 v = b + i < nblocks ? __ldg(agg + b + i) : make_uint2(0u, 0u);
        vp[i] = v.x; vt[i] = v.y; sp += v.x; st += v.y;
    }
    unsigned long long ip = sp, it = st;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long a = __shfl_up_sync(0xffffffffu, ip, o), c = __shfl_up_sync(0xffffffffu, it, o);
        if (lane >= (unsigned)o) { ip += a; it += c; }
    }
    if (lane == 31) { s_wp[wid] = ip; s_wt[wid] = it; }
    __syncthreads();
    unsigned long long wbp = 0, wbt = 0, tp = 0, tt = 0;
    for (int w = 0; w < 32; ++w) { if (w < (int)wid) { wbp += s_wp[w]; wbt += s_wt[w]; } tp += s_wp[w]; tt += s_wt[w]; }
    unsigned long long rp = cp + wbp + ip - sp, rt = ct + wbt + it - st;
#pragma unroll
    for (int i = 0; i < SS_ITEMS; ++i) {
        if (b + i < nblocks) { bases[2 * (size_t)(b + i)] = rp; bases[2 * (size_t)(b + i) + 1] = rt; }
        rp += vp[i]; rt += vt[i];
    }
    if (tid == 0 && blockIdx.x == gridDim.x - 1) {
        ctl[2] = (uint32_t)min(cp + tp, 0xffffffffull);
        ctl[3] = (uint32_t)min(ct + tt, 0xffffffffull);
    }
}

// K1c: with the CTA's global bases known, compact records and pairs are written in submission order.  The count pass left
// every triangle's tile count, so the prefix sums over the CTA's 1024 triangles need no vertex data (one barrier), and only
// the SURVIVORS are set up again — densely: survivor number i of the CTA is handled by thread i mod 256, whatever lane first
// saw it.  (In a scene of sub-pixel triangles one in sixteen survives: four barrier-separated rounds of a mostly idle CTA
// were latency bound at four CTAs per SM.)
__global__ void __launch_bounds__(SB_THREADS, SW_MIN_CTAS) setup_write_kernel(SegParams P)
{
    __shared__ uint32_t s_lo, s_wp[SB_ROUNDS][8], s_wt[SB_ROUNDS][8];
    __shared__ float s_mat[32];
    __shared__ float4 s_rec[SB_THREADS][4];
    __shared__ uint2 s_list[SB_BLOCK];            // (triangle, exclusive pair offset in the CTA) of the CTA's survivors, in order
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t g0 = blockIdx.x * (uint32_t)SB_BLOCK;
    uint32_t lo0, lo0_end;
    cta_draw_cache(P, g0, &s_lo, s_mat, lo0, lo0_end);
    const unsigned long long base_p = P.bases[2 * blockIdx.x], base_t = P.bases[2 * blockIdx.x + 1];

    // thread t owns the four consecutive triangles g0 + 4t .. 4t+3 (one 32-bit load of their tile-count bytes), so the prefix sums
    // over the CTA's 1024 triangles are ONE warp scan of two values and one pass over the eight warp totals (four scans when every
    // round of 256 was scanned by itself)
    static_assert(SB_ROUNDS == 4, "four tile-count bytes per thread");
    const uint32_t gt = g0 + 4u * tid;
    uint32_t nt[4] = {0u, 0u, 0u, 0u};
    if (gt + 3u < P.ntris) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(P.alive + gt));          // g0 is a multiple of 1024, the plane is 256-byte aligned
        nt[0] = w & 0xffu; nt[1] = (w >> 8) & 0xffu; nt[2] = (w >> 16) & 0xffu; nt[3] = w >> 24;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) nt[k] = gt + (uint32_t)k < P.ntris ? (uint32_t)__ldg(&P.alive[gt + k]) : 0u;
    }
    uint32_t sum_p = 0, sum_t = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (nt[k] == 255u) {                           // a large triangle: its exact tile count
            TriSetup T;
            setup_triangle<false>(P, gt + (uint32_t)k, lo0, lo0_end, s_mat, T);
            nt[k] = T.ntiles;
        }
        sum_p += nt[k]; sum_t += nt[k] ? 1u : 0u;
    }
    uint32_t ip = sum_p, it = sum_t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, ip, o), c = __shfl_up_sync(0xffffffffu, it, o);
        if (lane >= (unsigned)o) { ip += a; it += c; }
    }
    if (lane == 31) { s_wp[0][wid] = ip; s_wt[0][wid] = it; }
    __syncthreads();
    uint32_t run_p = 0, run_t = 0;
    {
        uint32_t wbp = 0, wbt = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { if (w < (int)wid) { wbp += s_wp[0][w]; wbt += s_wt[0][w]; } run_p += s_wp[0][w]; run_t += s_wt[0][w]; }
        uint32_t off_p = wbp + ip - sum_p, off_t = wbt + it - sum_t;          // exclusive prefixes of this thread's first triangle
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (nt[k]) { s_list[off_t] = make_uint2(gt + (uint32_t)k, off_p); off_p += nt[k]; ++off_t; }
    }
    __syncthreads();
    const uint32_t tot_t = run_t, tot_p = run_p;
    const bool room = base_p + tot_p <= P.pair_capacity;
    // ---- survivors in passes of 256: thread i sets up survivor c0 + i
    for (uint32_t c0 = 0; c0 < tot_t; c0 += SB_THREADS) {
        const uint32_t n = min((uint32_t)SB_THREADS, tot_t - c0);
        if (tid < n) {
            const uint2 me = s_list[c0 + tid];
            TriSetup T;
            setup_triangle<false>(P, me.x, lo0, lo0_end, s_mat, T);
            s_rec[tid][0] = make_float4(T.X[0], T.X[1], T.X[2], T.Y[0]);
            s_rec[tid][1] = make_float4(T.Y[1], T.Y[2], T.IW[0], T.IW[1]);
            s_rec[tid][2] = make_float4(T.IW[2], T.Z[0], T.Z[1], T.Z[2]);
            s_rec[tid][3] = make_float4(__uint_as_float((uint32_t)T.x0 | ((uint32_t)T.x1 << 16)), __uint_as_float((uint32_t)T.y0 | ((uint32_t)T.y1 << 16)),
                                        __uint_as_float(T.draw), __uint_as_float(T.tri));
            if (room) {
                const uint32_t target = __ldg(&P.draws[T.draw].target);
                const int tx0 = T.x0 >> TILE_SHIFT, tx1 = T.x1 >> TILE_SHIFT;
                const int ty0 = max(T.y0 >> TILE_SHIFT, (int)P.strip_ty0), ty1 = min(T.y1 >> TILE_SHIFT, (int)P.strip_ty1 - 1);
                const uint32_t tbase = target * P.tiles_x * P.tiles_y;
                const uint32_t ci = (uint32_t)(base_t + c0 + tid);
                uint64_t* out = P.pairs + base_p + me.y;
                for (int ty = ty0; ty <= ty1; ++ty) {
                    if (!owns_row(P, (uint32_t)ty)) continue;
                    for (int tx = tx0; tx <= tx1; ++tx) {
                        const uint32_t key = tbase + (uint32_t)ty * P.tiles_x + (uint32_t)tx;
                        *out++ = ((uint64_t)key << 32) | ci;
                        if (T.ordered) atomicOr(&P.tile_flags[key], 1u);
                    }
                }
            }
        }
        __syncthreads();
        {
            float4* dst = reinterpret_cast<float4*>(P.rec + base_t + c0);
            const float4* src = &s_rec[0][0];
            for (uint32_t q = tid; q < n * 4u; q += SB_THREADS) dst[q] = src[q];
        }
        if (c0 + SB_THREADS < tot_t) __syncthreads();      // s_rec is rewritten by the next pass
    }
}

/* ------------------------------------------------------------ exclusive scan -------- */
// three-phase scan of uint32 (block = 2048 items), used on the radix-sort histograms.

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total)
{
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += n;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= (unsigned)o) s += n;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    uint32_t base = wid ? warp_sums[wid - 1] : 0;
    if (total) *total = warp_sums[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ in, uint64_t* __restrict__ block_sums, size_t n)
{
    size_t base = (size_t)blockIdx.x * SCAN_BLOCK;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        size_t i = base + (size_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of the block sums in place; grand total -> block_sums[nblocks]
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(uint64_t* block_sums, size_t nblocks)
{
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t b0 = 0; b0 < nblocks; b0 += SCAN_THREADS) {
        size_t i = b0 + threadIdx.x;
        uint64_t v = i < nblocks ? block_sums[i] : 0;
        // 64-bit inclusive scan across the block via two 32-bit halves would lose carries; do it in smem
        __shared__ uint64_t tmp[SCAN_THREADS];
        tmp[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < SCAN_THREADS; o <<= 1) {
            uint64_t add = threadIdx.x >= (unsigned)o ? tmp[threadIdx.x - o] : 0;
            __syncthreads();
            tmp[threadIdx.x] += add;
            __syncthreads();
        }
        uint64_t incl = tmp[threadIdx.x];
        if (i < nblocks) block_sums[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[nblocks] = carry;
}

// out may alias in.  out[n] is NOT written here (grand total lives in block_sums[nblocks]).
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t* in, uint32_t* out, const uint64_t* __restrict__ block_sums, size_t n)
{
    size_t base = (size_t)blockIdx.x * SCAN_BLOCK + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = (base + k < n) ? in[base + k] : 0; s += v[k]; }
    uint32_t ex = block_exclusive_scan(s, nullptr) + (uint32_t)block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
}

/* ------------------------------------------------------------- tile boundaries ------ */
// tile_start[k] = first position of key k in the SORTED pair array (== position of the next larger key
// for empty tiles), tile_start[nkeys] = npairs.  Every entry is written exactly once: no memset, no atomics.
// tile_start[k] = first sorted pair whose key is >= k (k = nkeys: the end), by binary search: one thread per tile, no serial fill of
// the runs of empty tiles (a thread per PAIR that fills the gap before its key spent 52 us of a config-5 frame in the one
// thread that owned the empty tail)
__global__ void __launch_bounds__(256) tile_bounds_kernel(const uint64_t* __restrict__ pairs, uint32_t npairs, uint32_t nkeys,
                                                          uint32_t* __restrict__ tile_start)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nkeys) return;
    uint32_t lo = 0, hi = npairs;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((uint32_t)(__ldg(pairs + mid) >> 32) < k) lo = mid + 1; else hi = mid;
    }
    tile_start[k] = lo;
}

/* ------------------------------------------------------ stable LSD radix sort ------- */

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_PER_WARP = 256;                 // items per warp (8 rounds of 32)
constexpr int RS_BLOCK = RS_WARPS * RS_PER_WARP; // 2048 items per block
constexpr int RS_BINS = 256;

// block_hist[bin * nblocks + block]
__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const uint64_t* __restrict__ in, uint32_t* __restrict__ block_hist,
                                                                size_t n, int shift, uint32_t nblocks)
{
    __shared__ uint32_t h[RS_BINS];
    h[threadIdx.x] = 0;
    __syncthreads();
    size_t base = (size_t)blockIdx.x * RS_BLOCK;
    for (int k = 0; k < RS_BLOCK / RS_THREADS; ++k) {
        size_t i = base + (size_t)k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(in[i] >> shift) & (RS_BINS - 1)], 1u);
    }
    __syncthreads();
    block_hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                                                   const uint32_t* __restrict__ block_off, size_t n, int shift,
                                                                   uint32_t nblocks)
{
    __shared__ uint32_t wcount[RS_WARPS][RS_BINS];   // per-warp digit counts -> per-warp bases
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const size_t wbase = (size_t)blockIdx.x * RS_BLOCK + (size_t)wid * RS_PER_WARP;
    uint64_t items[RS_PER_WARP / 32];
    // pass 1: count digits of this warp's contiguous segment
#pragma unroll
    for (int r = 0; r < RS_PER_WARP / 32; ++r) {
        size_t i = wbase + (size_t)r * 32 + lane;
        bool valid = i < n;
        items[r] = valid ? in[i] : 0;
        if (valid) atomicAdd(&wcount[wid][(uint32_t)(items[r] >> shift) & (RS_BINS - 1)], 1u);
    }
    __syncthreads();
    // per bin: exclusive prefix over warps + global base of this (bin, block)
    {
        uint32_t bin = threadIdx.x;
        uint32_t run = block_off[(size_t)bin * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { uint32_t c = wcount[w][bin]; wcount[w][bin] = run; run += c; }
    }
    __syncthreads();
    // pass 2: stable scatter; ranks inside a round come from match_any (lower lanes first)
#pragma unroll
    for (int r = 0; r < RS_PER_WARP / 32; ++r) {
        size_t i = wbase + (size_t)r * 32 + lane;
        bool valid = i < n;
        unsigned active = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            uint32_t d = (uint32_t)(items[r] >> shift) & (RS_BINS - 1);
            unsigned peers = __match_any_sync(active, d);
            unsigned rank = __popc(peers & ((1u << lane) - 1u));
            uint32_t basepos = wcount[wid][d];
            out[basepos + rank] = items[r];
            __syncwarp(active);
            if (rank == 0) wcount[wid][d] = basepos + __popc(peers);
        }
        __syncwarp();
    }
}

/* ------------------------------------------------------------------ K3: raster ------ */

struct RasterParams {
    const DrawDev* draws;
    const TargetDev* targets;
    const TriRec* rec;             // compact per-triangle records (setup_bin_kernel)
    const uint64_t* pairs;         // sorted by tile key
    const uint32_t* tile_start;    // exclusive scan of tile counts
    const uint32_t* tile_flags;    // bit 0: in-order path required
    uint32_t force_ordered;        // config option raster_mode = 1: always take the in-order path (A/B, tests)
    uint32_t* counters;            // [ntargets][2]: tiles resolved by the visibility path / replayed in order
    uint32_t width, height, tiles_x, tiles_y;
    uint32_t strip_ty0, strip_rows;
    uint32_t band_tiles, band_phase, band_period;
};

__device__ __forceinline__ uint32_t shade_pack(float r, float g, float b)
{
    // saturate (Math::Clamp: max(min(v,1),0), OgreMath.h:395,704-708), *255, truncating cast (tinyrenderer.h:88-91)
    float c[3] = {r, g, b};
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = std_max(std_min(c[k], 1.0f), 0.0f);
        v = fmul(v, 255.0f);
        out |= ((uint32_t)x86_f2i(v) & 0xffu) << (8 * k);
    }
    return out;
}

// The fragment-stage part of the "beyond Tiny" state (ogrecu_draw_state, include/ogrecu.h): alpha rejection on the texel's
// alpha byte, then the interpolated vertex colour multiplied into rgb.  Deliberately NOT inlined: the Tiny path keeps its
// register budget (64 per thread, 4 CTAs per SM).
__device__ __noinline__ bool shade_beyond_tiny(const DrawDev* D, uint32_t ext, uint32_t tri, uint32_t alpha, float bc0, float bc1, float bc2,
                                               float& cr, float& cg, float& cb)
{
    if ((ext & OGRECU_EXT_ALPHA_REJECT) && !cmp_alpha((ext >> 16) & 0xfu, alpha, ext >> 24)) return false;
    if (ext & OGRECU_EXT_VERTEX_COLOUR) {
        const float inv255 = 1.0f / 255.0f;
        float vc[3][3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const uchar4 c = __ldg(reinterpret_cast<const uchar4*>(D->col + vertex_index(*D, tri, j) * D->col_stride));
            vc[j][0] = fmul((float)c.x, inv255); vc[j][1] = fmul((float)c.y, inv255); vc[j][2] = fmul((float)c.z, inv255);
        }
        cr = fmul(cr, fadd(fadd(fmul(vc[0][0], bc0), fmul(vc[1][0], bc1)), fmul(vc[2][0], bc2)));
        cg = fmul(cg, fadd(fadd(fmul(vc[0][1], bc0), fmul(vc[1][1], bc1)), fmul(vc[2][1], bc2)));
        cb = fmul(cb, fadd(fadd(fmul(vc[0][2], bc0), fmul(vc[1][2], bc1)), fmul(vc[2][2], bc2)));
    }
    return true;
}

// DefaultShader::fragment + the blend / pack tail of triangle() for one fragment
// (OgreTinyRenderSystem.cpp:310-333, tinyrenderer.h:82-93).  Returns false when the fragment is discarded.
// The varyings DefaultShader::vertex produced (var_uv = (Tex*(u,v,0,1)).xy, var_normal = MVIT.linear()*n,
// .cpp:304-308) are recomputed here from the source streams with the same operations, only for fragments
// that are actually shaded, instead of being stored per triangle.
__device__ __forceinline__ bool shade_fragment(const DrawDev* D, uint32_t flags, uint32_t tri, float bc0, float bc1, float bc2,
                                               uint32_t& rgb)
{
    const float inv255 = 1.0f / 255.0f;
    float cr = 1.0f, cg = 1.0f, cb = 1.0f;   // default ColourValue: white (tinyrenderer.h:82)
    const uint8_t* tex = D->tex;
    const bool lit = (flags & OGRECU_DRAW_LIGHTING) != 0;
    const uint32_t ext = __ldg(&D->ext);
    uint32_t alpha = 255u;
    size_t vi[3];
    if (tex || lit) {
#pragma unroll
        for (int j = 0; j < 3; ++j) vi[j] = vertex_index(*D, tri, j);
    }
    if (tex) {
        float uu[3], vv[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (D->uv) {
                const float* p = (const float*)(D->uv + vi[j] * D->uv_stride);
                const float a = __ldg(p), b = __ldg(p + 1);
                const float* m = D->texm;   // rows 0 and 1 of Matrix4 * (u, v, 0, 1), OgreMatrix4.h:571-578
                uu[j] = fadd(fadd(fadd(fmul(m[0], a), fmul(m[1], b)), fmul(m[2], 0.0f)), fmul(m[3], 1.0f));
                vv[j] = fadd(fadd(fadd(fmul(m[4], a), fmul(m[5], b)), fmul(m[6], 0.0f)), fmul(m[7], 1.0f));
            } else { uu[j] = 0.f; vv[j] = 0.f; }
        }
        const int tw_ = (int)D->tex_w, th_ = (int)D->tex_h;
        float u = fadd(fadd(fmul(uu[0], bc0), fmul(uu[1], bc1)), fmul(uu[2], bc2));
        float v = fadd(fadd(fmul(vv[0], bc0), fmul(vv[1], bc1)), fmul(vv[2], bc2));
        int ix = x86_f2i(fmul(u, (float)(uint32_t)(tw_ - 1)));                 // sample2D, OgreTinyRenderSystem.h:36-41
        int iy = x86_f2i(fmul(v, (float)(uint32_t)(th_ - 1)));
        ix = (tw_ + (ix % tw_)) % tw_;
        iy = (th_ + (iy % th_)) % th_;
        uchar4 t4 = __ldg(reinterpret_cast<const uchar4*>(tex) + (size_t)iy * tw_ + ix);
        if (t4.w < 1) return false;                                            // .cpp:318
        cr = fmul((float)t4.x, inv255); cg = fmul((float)t4.y, inv255); cb = fmul((float)t4.z, inv255);
        alpha = t4.w;
    }
    if (ext & (OGRECU_EXT_ALPHA_REJECT | OGRECU_EXT_VERTEX_COLOUR))            // beyond Tiny, out of line: Tiny draws pay one test
        if (!shade_beyond_tiny(D, ext, tri, alpha, bc0, bc1, bc2, cr, cg, cb)) return false;
    if (lit) {
        float nn[3][3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float* p = (const float*)(D->nrm + vi[j] * D->nrm_stride);    // lighting implies normals (run_segment)
            const float a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
#pragma unroll
            for (int r = 0; r < 3; ++r)                                        // Matrix3 * Vector3, OgreMatrix3.h:321-327
                nn[j][r] = fadd(fadd(fmul(D->mvit[3 * r + 0], a), fmul(D->mvit[3 * r + 1], b)), fmul(D->mvit[3 * r + 2], c));
        }
        float nx = fadd(fadd(fmul(nn[0][0], bc0), fmul(nn[1][0], bc1)), fmul(nn[2][0], bc2));
        float ny = fadd(fadd(fmul(nn[0][1], bc0), fmul(nn[1][1], bc1)), fmul(nn[2][1], bc2));
        float nz = fadd(fadd(fmul(nn[0][2], bc0), fmul(nn[1][2], bc1)), fmul(nn[2][2], bc2));
        float dp = fadd(fadd(fadd(0.0f, fmul(nx, D->light[0])), fmul(ny, D->light[1])), fmul(nz, D->light[2]));
        float diffuse = std_max(0.f, dp);
        cr = fadd(fmul(cr, diffuse), D->ambient[0]);
        cg = fadd(fmul(cg, diffuse), D->ambient[1]);
        cb = fadd(fmul(cb, diffuse), D->ambient[2]);
    }
    if (flags & OGRECU_DRAW_BLEND_ADD) {                                      // tinyrenderer.h:86-87
        cr = fadd(cr, fmul((float)(rgb & 0xff), inv255));
        cg = fadd(cg, fmul((float)((rgb >> 8) & 0xff), inv255));
        cb = fadd(cb, fmul((float)((rgb >> 16) & 0xff), inv255));
    }
    rgb = shade_pack(cr, cg, cb);
    return true;
}

// TriRec -> the 20-word form the pixel loops read: hf[0..8] = Matrix3::Inverse of
// ABC = [[X0,X1,X2],[Y0,Y1,Y2],[1,1,1]] (OgreMatrix3.cpp:160-192, incl. the UNSCALED cofactors when
// abs(det) <= 1e-6, :184-185), hf[9..11] = 1/w, hf[12..14] = z, words 15..18 = bx, by, draw, tri.
// Pixel independent, so it is evaluated once per (triangle, tile) instead of once per pixel as Tiny does.
struct RecRaw { float4 a, b, c, d; };
__device__ __forceinline__ RecRaw load_rec(const TriRec* r)
{
    const float4* rp = reinterpret_cast<const float4*>(r);
    return RecRaw{__ldg(rp), __ldg(rp + 1), __ldg(rp + 2), __ldg(rp + 3)};
}

// hf[0..8] = Matrix3::Inverse of [[X0,X1,X2],[Y0,Y1,Y2],[1,1,1]] (the one definition: set-up's coverage cull and the raster
// kernel's pixel loops must agree bit for bit)
__device__ __forceinline__ void expand_xy(float X0, float X1, float X2, float Y0, float Y1, float Y2, float hf[20])
{
    hf[0] = fsub(Y1, Y2); hf[1] = fsub(X2, X1); hf[2] = fsub(fmul(X1, Y2), fmul(X2, Y1));
    hf[3] = fsub(Y2, Y0); hf[4] = fsub(X0, X2); hf[5] = fsub(fmul(X2, Y0), fmul(X0, Y2));
    hf[6] = fsub(Y0, Y1); hf[7] = fsub(X1, X0); hf[8] = fsub(fmul(X0, Y1), fmul(X1, Y0));
    const float det = fadd(fadd(fmul(X0, hf[0]), fmul(X1, hf[3])), fmul(X2, hf[6]));
    if (!(fabsf(det) <= 1e-06f)) {
        const float gi = fdiv(1.0f, det);
#pragma unroll
        for (int k = 0; k < 9; ++k) hf[k] = fmul(hf[k], gi);
    }
}

__device__ __forceinline__ void expand_raw(const RecRaw& R, float hf[20])
{
    const float4 a = R.a, b = R.b, c = R.c, d = R.d;
    expand_xy(a.x, a.y, a.z, a.w, b.x, b.y, hf);
    hf[9] = b.z; hf[10] = b.w; hf[11] = c.x;
    hf[12] = c.y; hf[13] = c.z; hf[14] = c.w;
    hf[15] = d.x; hf[16] = d.y; hf[17] = d.z; hf[18] = d.w; hf[19] = 0.f;
}

__device__ __forceinline__ void expand_rec(const TriRec* r, float hf[20]) { expand_raw(load_rec(r), hf); }

// barycentric() + perspective correction + depth of one pixel against one hot record
// (tinyrenderer.h:29-35,70-76): same operations, same order, for both raster paths.
__device__ __forceinline__ bool eval_fragment(const float* hf, float Px, float Py, float& bc0, float& bc1, float& bc2, float& fd)
{
    float bs0 = fadd(fadd(fmul(hf[0], Px), fmul(hf[1], Py)), hf[2]);
    float bs1 = fadd(fadd(fmul(hf[3], Px), fmul(hf[4], Py)), hf[5]);
    float bs2 = fadd(fadd(fmul(hf[6], Px), fmul(hf[7], Py)), hf[8]);
    bc0 = fmul(bs0, hf[9]); bc1 = fmul(bs1, hf[10]); bc2 = fmul(bs2, hf[11]);   // :71
    float s = fadd(fadd(bc0, bc1), bc2);
    float f = fdiv(1.0f, s);                                                   // :72, OgreVector.h:659-664
    bc0 = fmul(f, bc0); bc1 = fmul(f, bc1); bc2 = fmul(f, bc2);
    fd = fadd(fadd(fadd(0.0f, fmul(hf[12], bc0)), fmul(hf[13], bc1)), fmul(hf[14], bc2));   // :73
    return !(bs0 < 0.f || bs1 < 0.f || bs2 < 0.f) && !(fd < 0.0f);             // :74,:76
}

// eval_fragment's verdict alone, the three edge functions first: most pixels of a sub-pixel triangle's box fail them and never
// reach the perspective division (same operations in the same order where they are evaluated)
__device__ __forceinline__ bool covers_pixel(const float* hf, float Px, float Py)
{
    const float bs0 = fadd(fadd(fmul(hf[0], Px), fmul(hf[1], Py)), hf[2]);
    const float bs1 = fadd(fadd(fmul(hf[3], Px), fmul(hf[4], Py)), hf[5]);
    const float bs2 = fadd(fadd(fmul(hf[6], Px), fmul(hf[7], Py)), hf[8]);
    if (bs0 < 0.f || bs1 < 0.f || bs2 < 0.f) return false;
    float bc0 = fmul(bs0, hf[9]), bc1 = fmul(bs1, hf[10]), bc2 = fmul(bs2, hf[11]);
    const float f = fdiv(1.0f, fadd(fadd(bc0, bc1), bc2));
    bc0 = fmul(f, bc0); bc1 = fmul(f, bc1); bc2 = fmul(f, bc2);
    const float fd = fadd(fadd(fadd(0.0f, fmul(hf[12], bc0)), fmul(hf[13], bc1)), fmul(hf[14], bc2));
    return !(fd < 0.0f);
}

#ifndef RASTER_PREFETCH
#define RASTER_PREFETCH 1       // 0: none, 1: pairs one chunk ahead (C5 2.21 -> 2.15 ms), 2: also records one chunk ahead (80 registers: 2.18 ms)
#endif
#ifndef RT_MIN_CTAS
#define RT_MIN_CTAS 5           // tile CTAs per SM the register allocation is bounded for (48 registers, 40 bytes spilled: C4 -1.7 %, C5 -1.8 % against 4; 6: 40 registers, 188 bytes, no further gain)
#endif
constexpr int VIS_SMALL_MAX = 12;   // bbox pixels a single thread walks; larger triangles go to a warp

// One CTA per 16x16 tile.  Two ways to resolve the tile's triangle list, same result:
//  * visibility path (tiles whose triangles are all opaque, depth-tested, depth-writing, never
//    discarded): fragments of ALL triangles are generated in parallel (thread per small triangle,
//    warp per large one) and resolved per pixel with a 64-bit shared-memory atomicMin on
//    (depth bits, ~submission order) — Tiny's in-order loop ends with exactly that fragment: the
//    LAST one in submission order among those of minimal depth not above the initial depth
//    ("frag_depth > zbuf" rejects, equal depth passes, tinyrenderer.h:79).  Only the winner of each
//    pixel is shaded.  A NaN depth (degenerate triangle) or an initial depth that is NaN/negative
//    makes the outcome order dependent: the tile then falls back to
//  * the in-order path: one pixel per thread, triangles replayed in submission order.
template <bool VEC>
__global__ void __launch_bounds__(RT_THREADS, RT_MIN_CTAS) raster_kernel(RasterParams P)
{
    __shared__ float4 s_hot[CHUNK][5];
    __shared__ __align__(16) uint8_t s_rgb[TILE][TILE * 3];
    __shared__ uint32_t s_nbig[2];

    // tile of this CTA
    const uint32_t per_target = P.tiles_x * P.strip_rows;
    const uint32_t target = blockIdx.x / per_target;
    const uint32_t rem = blockIdx.x - target * per_target;
    uint32_t ty = P.strip_ty0 + rem / P.tiles_x;
    const uint32_t tx = rem % P.tiles_x;
    if (P.band_period > 1) {                     // row r of this context's bands -> tile row of the frame
        const uint32_t r = rem / P.tiles_x;
        ty = (P.band_phase + (r / P.band_tiles) * P.band_period) * P.band_tiles + r % P.band_tiles;
        if (ty >= P.tiles_y) return;
    }
    const uint32_t key = target * P.tiles_x * P.tiles_y + ty * P.tiles_x + tx;
    const uint32_t start = P.tile_start[key], end = P.tile_start[key + 1];
    if (start == end) return;   // nothing to draw here: the planes already hold the result

    const TargetDev T = P.targets[target];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // warp = 8x4 pixel block of the tile
    const int wx0 = (wid & 1) * 8, wy0 = (wid >> 1) * 4;
    const int lx = wx0 + (lane & 7), ly = wy0 + (lane >> 3);
    const int tile_x0 = tx * TILE, tile_y0 = ty * TILE;
    const int x = tile_x0 + lx, y = tile_y0 + ly;
    const bool inside = x < (int)P.width && y < (int)P.height;
    const int tw = min(TILE, (int)P.width - tile_x0), th = min(TILE, (int)P.height - tile_y0);
    const size_t pix = (size_t)y * P.width + x;

    // ---- load the tile: depth straight to registers, colour through shared memory
    float zbuf = inside ? T.depth[pix] : 0.f;
    if (VEC) {
        // full-width tile rows are 48 bytes = three aligned 16-byte vectors
        if (threadIdx.x < TILE * 3) {
            int r = threadIdx.x / 3, v = threadIdx.x % 3;
            if (r < th) {
                const uint4* src = reinterpret_cast<const uint4*>(T.colour + ((size_t)(tile_y0 + r) * P.width + tile_x0) * 3) + v;
                reinterpret_cast<uint4*>(&s_rgb[r][0])[v] = *src;
            }
        }
    } else {
        for (int i = threadIdx.x; i < th * tw * 3; i += RT_THREADS) {
            int r = i / (tw * 3), b = i - r * (tw * 3);
            s_rgb[r][b] = T.colour[((size_t)(tile_y0 + r) * P.width + tile_x0) * 3 + b];
        }
    }
    __syncthreads();
    uint32_t rgb = 0;
    if (inside) rgb = (uint32_t)s_rgb[ly][lx * 3] | ((uint32_t)s_rgb[ly][lx * 3 + 1] << 8) | ((uint32_t)s_rgb[ly][lx * 3 + 2] << 16);

    const float Px = (float)x, Py = (float)y;
    bool resolved = false;

    // =============================================================== visibility path
    if (!P.force_ordered && !(__ldg(&P.tile_flags[key]) & 1u)) {
        unsigned long long* s_key = reinterpret_cast<unsigned long long*>(&s_hot[0][0]);   // [256]
        uint32_t* s_big = reinterpret_cast<uint32_t*>(s_key + RT_THREADS);                   // [256]
        const uint32_t pidx = (uint32_t)(ly * TILE + lx);
        // the initial depth takes part in the minimum: a fragment wins iff depth <= initial depth
        const bool zbad = inside && !(zbuf >= 0.0f);          // NaN or negative: order dependent
        s_key[pidx] = ((unsigned long long)__float_as_uint(zbuf == 0.0f ? 0.0f : zbuf) << 32) | 0xFFFFFFFFull;
        if (threadIdx.x < 2) s_nbig[threadIdx.x] = 0;
        int fallback = __syncthreads_or(zbad);
        const int tx1 = tile_x0 + tw - 1, ty1 = tile_y0 + th - 1;
        uint32_t par = 0;
        // Software pipeline over the chunks of 256 triangles: the pair (tile key | compact id) is loaded two chunks ahead
        // and the 64-byte record one chunk ahead, so neither round trip of the pair -> record chain is exposed between
        // the barriers of a chunk (ncu r01k: 23 % of the stall samples sat on those two dependent loads).
#if RASTER_PREFETCH >= 1
        uint32_t g_next = start + threadIdx.x < end ? (uint32_t)__ldg(&P.pairs[start + threadIdx.x]) : 0xffffffffu;
#endif
#if RASTER_PREFETCH >= 2
        RecRaw raw_next = {};
        if (g_next != 0xffffffffu) raw_next = load_rec(P.rec + g_next);
        g_next = start + RT_THREADS + threadIdx.x < end ? (uint32_t)__ldg(&P.pairs[start + RT_THREADS + threadIdx.x]) : 0xffffffffu;
#endif
        for (uint32_t c0 = start; c0 < end && !fallback; c0 += RT_THREADS, par ^= 1u) {
            const uint32_t i = c0 + threadIdx.x;
            int nan_seen = 0;
#if RASTER_PREFETCH >= 2
            const RecRaw raw = raw_next;                              // this chunk's record, loaded during the previous chunk
            if (g_next != 0xffffffffu) raw_next = load_rec(P.rec + g_next);
            {
                const uint32_t i2 = i + 2 * RT_THREADS;
                g_next = i2 < end ? (uint32_t)__ldg(&P.pairs[i2]) : 0xffffffffu;
            }
#elif RASTER_PREFETCH == 1
            const uint32_t g_cur = g_next;
            g_next = i + RT_THREADS < end ? (uint32_t)__ldg(&P.pairs[i + RT_THREADS]) : 0xffffffffu;
#endif
            if (i < end) {
                float hf[20];
#if RASTER_PREFETCH >= 2
                expand_raw(raw, hf);
#elif RASTER_PREFETCH == 1
                expand_rec(P.rec + g_cur, hf);
#else
                const uint32_t g = (uint32_t)P.pairs[i];
                expand_rec(P.rec + g, hf);
#endif
                const uint32_t bx = __float_as_uint(hf[15]), by = __float_as_uint(hf[16]);
                const int x0 = max((int)(bx & 0xffff), tile_x0), x1 = min((int)(bx >> 16), tx1);
                const int y0 = max((int)(by & 0xffff), tile_y0), y1 = min((int)(by >> 16), ty1);
                const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
                if (bw > 0 && bh > 0) {
                    if (bw * bh > VIS_SMALL_MAX) {
                        s_big[atomicAdd(&s_nbig[par], 1u)] = i - start;
                    } else {
                        const unsigned long long lo = 0xFFFFFFFFull - (unsigned long long)(i - start + 1);
                        for (int py = y0; py <= y1; ++py)
                            for (int px = x0; px <= x1; ++px) {
                                float b0, b1, b2, fd;
                                if (eval_fragment(hf, (float)px, (float)py, b0, b1, b2, fd)) {
                                    if (fd != fd) nan_seen = 1;
                                    else atomicMin(&s_key[(py - tile_y0) * TILE + (px - tile_x0)],
                                                   ((unsigned long long)__float_as_uint(fd == 0.0f ? 0.0f : fd) << 32) | lo);
                                }
                            }
                    }
                }
            }
            __syncthreads();
            // large triangles of this chunk: one warp each, lanes over the bbox pixels inside the tile
            const uint32_t nbig = s_nbig[par];
            for (uint32_t b = wid; b < nbig; b += RT_THREADS / 32) {
                const uint32_t ord = s_big[b];
                const uint32_t g = (uint32_t)P.pairs[start + ord];
                float hf[20];
                expand_rec(P.rec + g, hf);
                const uint32_t bx = __float_as_uint(hf[15]), by = __float_as_uint(hf[16]);
                const int x0 = max((int)(bx & 0xffff), tile_x0), x1 = min((int)(bx >> 16), tx1);
                const int y0 = max((int)(by & 0xffff), tile_y0), y1 = min((int)(by >> 16), ty1);
                const int bw = x1 - x0 + 1, npx = bw * (y1 - y0 + 1);
                const unsigned long long lo = 0xFFFFFFFFull - (unsigned long long)(ord + 1);
                for (int q = lane; q < npx; q += 32) {
                    const int py = y0 + q / bw, px = x0 + q % bw;
                    float b0, b1, b2, fd;
                    if (eval_fragment(hf, (float)px, (float)py, b0, b1, b2, fd)) {
                        if (fd != fd) nan_seen = 1;
                        else atomicMin(&s_key[(py - tile_y0) * TILE + (px - tile_x0)],
                                       ((unsigned long long)__float_as_uint(fd == 0.0f ? 0.0f : fd) << 32) | lo);
                    }
                }
            }
            if (threadIdx.x == 0) s_nbig[par ^ 1u] = 0;      // the other counter is idle during this chunk
            fallback = __syncthreads_or(nan_seen);
        }
        if (!fallback) {
            // ---- shade the winner of my pixel (deferred: one fragment per covered pixel)
            const unsigned long long k = s_key[pidx];
            const uint32_t field = (uint32_t)k;
            if (inside && field != 0xFFFFFFFFu) {
                const uint32_t ord = 0xFFFFFFFFu - field - 1u;
                const uint32_t g = (uint32_t)P.pairs[start + ord];
                float hf[20];
                expand_rec(P.rec + g, hf);
                float b0, b1, b2, fd;
                eval_fragment(hf, Px, Py, b0, b1, b2, fd);      // same operations as in the resolve pass
                const DrawDev* D = P.draws + __float_as_uint(hf[17]);
                shade_fragment(D, __ldg(&D->flags), __float_as_uint(hf[18]), b0, b1, b2, rgb);   // never discards here
                zbuf = fd;
            }
            resolved = true;
        }
        __syncthreads();   // s_key (aliases s_hot) is dead from here on
    }

    // ================================================================== in-order path
    if (!resolved) {
        // warp's pixel rectangle in screen space (for the per-warp triangle prefilter)
        const int wrx0 = tile_x0 + wx0, wrx1 = wrx0 + 7, wry0 = tile_y0 + wy0, wry1 = wry0 + 3;
        for (uint32_t c0 = start; c0 < end; c0 += CHUNK) {
            const uint32_t nchunk = min((uint32_t)CHUNK, end - c0);
            __syncthreads();   // previous chunk fully consumed
            // ---- stage the chunk: one thread expands one record (inverse computed once per triangle and tile)
            if (threadIdx.x < nchunk) {
                float hf[20];
                expand_rec(P.rec + (uint32_t)P.pairs[c0 + threadIdx.x], hf);
#pragma unroll
                for (int k = 0; k < 5; ++k) s_hot[threadIdx.x][k] = make_float4(hf[4 * k], hf[4 * k + 1], hf[4 * k + 2], hf[4 * k + 3]);
            }
            __syncthreads();

            // ---- replay the chunk in submission order
            for (uint32_t r0 = 0; r0 < nchunk; r0 += 32) {
                // prefilter: lane L tests triangle r0+L against this warp's 8x4 pixels
                bool hit = false;
                if (r0 + lane < nchunk) {
                    const uint32_t* w = reinterpret_cast<const uint32_t*>(&s_hot[r0 + lane][0]);
                    uint32_t bx = w[15], by = w[16];
                    int bx0 = bx & 0xffff, bx1 = bx >> 16, by0 = by & 0xffff, by1 = by >> 16;
                    hit = bx0 <= wrx1 && bx1 >= wrx0 && by0 <= wry1 && by1 >= wry0;
                }
                unsigned mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {
                    const int b = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const uint32_t j = r0 + b;
                    const float* hf = reinterpret_cast<const float*>(&s_hot[j][0]);
                    const uint32_t* hw = reinterpret_cast<const uint32_t*>(hf);
                    const uint32_t bx = hw[15], by = hw[16];
                    const bool in_box = inside && x >= (int)(bx & 0xffff) && x <= (int)(bx >> 16) && y >= (int)(by & 0xffff) && y <= (int)(by >> 16);
                    bool pass = false;
                    float bc0 = 0.f, bc1 = 0.f, bc2 = 0.f, fd = 0.f;
                    const DrawDev* D = P.draws + hw[17];
                    const uint32_t flags = __ldg(&D->flags);
                    if (in_box) {
                        pass = eval_fragment(hf, Px, Py, bc0, bc1, bc2, fd);
                        if (pass && (flags & OGRECU_DRAW_DEPTH_TEST)) {
                            const uint32_t ext = __ldg(&D->ext);
                            if (ext & OGRECU_EXT_DEPTH_FUNC) pass = cmp_depth((ext >> 8) & 0xfu, fd, zbuf);
                            else if (fd > zbuf) pass = false;                                      // :79
                        }
                    }
                    if (!__any_sync(0xffffffffu, pass)) continue;
                    if (pass && shade_fragment(D, flags, hw[18], bc0, bc1, bc2, rgb))
                        if (flags & OGRECU_DRAW_DEPTH_WRITE) zbuf = fd;                           // :92-93
                }
            }
        }
    }

    if (threadIdx.x == 0) atomicAdd(&P.counters[2 * target + (resolved ? 0 : 1)], 1u);

    // ---- flush: depth from registers, colour via shared memory as 16-byte vectors
    __syncthreads();
    if (inside) {
        T.depth[pix] = zbuf;
        s_rgb[ly][lx * 3] = (uint8_t)(rgb & 0xff);
        s_rgb[ly][lx * 3 + 1] = (uint8_t)((rgb >> 8) & 0xff);
        s_rgb[ly][lx * 3 + 2] = (uint8_t)((rgb >> 16) & 0xff);
    }
    __syncthreads();
    if (VEC) {
        if (threadIdx.x < TILE * 3) {
            int r = threadIdx.x / 3, v = threadIdx.x % 3;
            if (r < th) {
                uint4* dst = reinterpret_cast<uint4*>(T.colour + ((size_t)(tile_y0 + r) * P.width + tile_x0) * 3) + v;
                *dst = reinterpret_cast<const uint4*>(&s_rgb[r][0])[v];
            }
        }
    } else {
        for (int i = threadIdx.x; i < th * tw * 3; i += RT_THREADS) {
            int r = i / (tw * 3), b = i - r * (tw * 3);
            T.colour[((size_t)(tile_y0 + r) * P.width + tile_x0) * 3 + b] = s_rgb[r][b];
        }
    }
}

/* ------------------------------------------------------------------ K5: clear ------- */

// colour plane filled with a 3-byte pattern using 16-byte stores; depth with float4 stores
__global__ void clear_colour_kernel(uint8_t* __restrict__ colour, size_t nbytes, uint32_t r, uint32_t g, uint32_t b)
{
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i >= nbytes) return;
    const uint32_t px[3] = {r, g, b};
    if (i + 16 <= nbytes) {
        uint32_t w[4];
        uint32_t ph = (uint32_t)(i % 3);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) { v |= px[ph] << (8 * q); ph = ph == 2 ? 0 : ph + 1; }
            w[k] = v;
        }
        *reinterpret_cast<uint4*>(colour + i) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        for (size_t k = i; k < nbytes; ++k) colour[k] = (uint8_t)px[k % 3];
    }
}

// RGB8 plane -> RGBA8 texture (alpha 255): one thread per 4 pixels = 12 bytes in, 16 bytes out
__global__ void rgb_to_rgba_kernel(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ rgba, size_t npx)
{
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t p0 = q * 4;
    if (p0 >= npx) return;
    if (p0 + 4 <= npx) {
        const uint32_t* s = reinterpret_cast<const uint32_t*>(rgb + p0 * 3);   // plane base is 256-byte aligned, 12*q is 4-aligned
        uint32_t a = s[0], b = s[1], c = s[2];
        uint4 o;
        o.x = (a & 0x00ffffffu) | 0xff000000u;
        o.y = ((a >> 24) | (b << 8)) & 0x00ffffffu | 0xff000000u;
        o.z = ((b >> 16) | (c << 16)) & 0x00ffffffu | 0xff000000u;
        o.w = (c >> 8) | 0xff000000u;
        *reinterpret_cast<uint4*>(rgba + p0 * 4) = o;
    } else {
        for (size_t p = p0; p < npx; ++p) {
            rgba[p * 4] = rgb[p * 3]; rgba[p * 4 + 1] = rgb[p * 3 + 1]; rgba[p * 4 + 2] = rgb[p * 3 + 2]; rgba[p * 4 + 3] = 0xff;
        }
    }
}

// Sort-last composite of one frame (or band) rendered by a LATER rank into the accumulated frame of the earlier ranks:
// Tiny's per-pixel outcome over a sequence of opaque, depth-tested, depth-writing triangles is the fragment with the
// smallest depth, the later one on ties (`depth > zbuf` rejects, tinyrenderer.h:76) — the same minimum over
// (depth, ~submission order) the visibility path of raster_kernel takes.  The minimum over all triangles is the minimum
// of the per-rank minima when rank r holds triangles that come after those of rank r-1.  A source pixel that still
// holds the cleared colour and depth was not touched and never wins.
__global__ void composite_depth_kernel(uint8_t* __restrict__ dst_rgb, float* __restrict__ dst_z, const uint8_t* __restrict__ src_rgb,
                                       const float* __restrict__ src_z, size_t npx, uint32_t nsrc, size_t src_stride_px, uint32_t clear_rgb,
                                       float clear_z)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npx) return;
    float dz = dst_z[i];
    uint32_t dc = 0;
    bool changed = false;
    for (uint32_t s = 0; s < nsrc; ++s) {                             // sources in rank order: a later one wins ties
        const size_t j = (size_t)s * src_stride_px + i;
        const float sz = src_z[j];
        if (!(sz <= dz)) continue;
        const uint32_t c = (uint32_t)src_rgb[3 * j] | ((uint32_t)src_rgb[3 * j + 1] << 8) | ((uint32_t)src_rgb[3 * j + 2] << 16);
        if (sz == clear_z && c == clear_rgb) continue;                // untouched
        dz = sz; dc = c; changed = true;
    }
    if (changed) {
        dst_z[i] = dz;
        dst_rgb[3 * i] = (uint8_t)dc; dst_rgb[3 * i + 1] = (uint8_t)(dc >> 8); dst_rgb[3 * i + 2] = (uint8_t)(dc >> 16);
    }
}


// Sort-last merge of a split target (device group): the root rendered its slice of the triangles into the real planes, peer j
// rendered slice j into private planes that started with depth +INF.  Folding the peers in slice order with Tiny's own test —
// a fragment is rejected when depth > zbuf (tinyrenderer.h:77) — gives what the in-order loop over all slices leaves behind for
// opaque, depth-tested, depth-writing content: the smallest depth, the latest submission among equals.  Every device merges its
// own share of the pixels, reading the peers' planes and writing the root's through peer access.
struct MergeSrc { const uint8_t* rgb[OGRECU_MAX_PEERS]; const float* z[OGRECU_MAX_PEERS]; int n; };
__global__ void __launch_bounds__(256) merge_sort_last_kernel(uint8_t* __restrict__ root_rgb, float* __restrict__ root_z, MergeSrc S, size_t p0, size_t p1)
{
    const size_t p = p0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p1) return;
    float z = root_z[p];
    int best = -1;
    for (int j = 0; j < S.n; ++j) {
        const float zj = S.z[j][p];
        if (__float_as_uint(zj) != 0x7f800000u && !(zj > z)) { z = zj; best = j; }
    }
    if (best >= 0) {
        root_z[p] = z;
        const uint8_t* c = S.rgb[best] + p * 3;
        uint8_t* d = root_rgb + p * 3;
        d[0] = c[0]; d[1] = c[1]; d[2] = c[2];
    }
}

__global__ void clear_depth_kernel(float* __restrict__ depth, size_t n, float v)
{
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    if (i + 4 <= n) *reinterpret_cast<float4*>(depth + i) = make_float4(v, v, v, v);
    else for (size_t k = i; k < n; ++k) depth[k] = v;
}

// clearFrameBuffer of MANY targets of one size in one launch (the batched views of config 4 clear every target before their shared
// pipeline pass: two launches per view were 9 % of a view's time): blockIdx.y = target, the x dimension covers the colour plane's
// 16-byte chunks and then the depth plane's.  Same bytes as clear_colour_kernel / clear_depth_kernel.
constexpr int CLEAR_BATCH = 64;
struct ClearPlanes { uint8_t* colour[CLEAR_BATCH]; float* depth[CLEAR_BATCH]; };
__global__ void __launch_bounds__(256) clear_targets_kernel(const __grid_constant__ ClearPlanes T, size_t npx, uint32_t r, uint32_t g, uint32_t b, float z,
                                                            unsigned buffers)
{
    const size_t colour_chunks = (buffers & OGRECU_CLEAR_COLOUR) ? (npx * 3 + 15) / 16 : 0;
    const size_t depth_chunks = (buffers & OGRECU_CLEAR_DEPTH) ? (npx + 3) / 4 : 0;
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < colour_chunks) {
        uint8_t* colour = T.colour[blockIdx.y];
        const size_t nbytes = npx * 3, i = c * 16;
        const uint32_t px[3] = {r, g, b};
        if (i + 16 <= nbytes) {
            uint32_t w[4];
            uint32_t ph = (uint32_t)(i % 3);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t v = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) { v |= px[ph] << (8 * q); ph = ph == 2 ? 0 : ph + 1; }
                w[k] = v;
            }
            *reinterpret_cast<uint4*>(colour + i) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (size_t k = i; k < nbytes; ++k) colour[k] = (uint8_t)px[k % 3];
        }
        return;
    }
    c -= colour_chunks;
    if (c < depth_chunks) {
        float* depth = T.depth[blockIdx.y];
        const size_t i = c * 4;
        if (i + 4 <= npx) *reinterpret_cast<float4*>(depth + i) = make_float4(z, z, z, z);
        else for (size_t k = i; k < npx; ++k) depth[k] = z;
    }
}

}  // namespace

/* ======================================================================= host side === */

struct Cmd {
    int kind;                  // 0 = draw, 1 = clear
    ogrecu_draw_desc draw;
    ogrecu_draw_state state;   // enable == 0: a plain Tiny draw
    unsigned clear_buffers;
    float clear_colour[4];
    float clear_depth;
};

struct ogrecu_rt {
    ogrecu_ctx* ctx;
    uint32_t w, h;
    uint8_t* colour;
    float* depth;
    uint32_t strip_y0, strip_y1;     // pixel rows rasterised here (multiples of TILE, or h)
    uint32_t band_tiles = 1, band_phase = 0, band_period = 1;   // interleaved bands of tile rows (period 1 = every band)
    // device group: a SPLIT target's planes live on the root; aliases[k - 1] describes the same planes to peer k, which rasterises
    // the bands it owns straight into them over NVLink (no composite pass).  A PLACED target lives wholly on one peer (ctx = that peer).
    bool split = false;
    std::vector<ogrecu_rt*> aliases;
    uint8_t* sl_colour = nullptr;      // aliases: this peer's private planes for the sort-last form of a split segment
    float* sl_depth = nullptr;
    std::vector<Cmd> cmds;
    ogrecu_raster_stats stats;
    int stats_slot;          // index of this target in the last segment's counter table
    bool stats_pending;      // counters of the last segment not fetched yet
};

namespace {

// per-context raster scratch (grown on demand, reused across flushes)
struct RasterScratch {
    Scratch draws, targets, rec, cnt, sums, pairs_a, pairs_b, tile_count, hist, counters, xf, alive;
};
std::map<ogrecu_ctx*, RasterScratch>& scratch_map()
{
    static std::map<ogrecu_ctx*, RasterScratch> m;
    return m;
}
// a context's arenas; the map itself is shared by the host threads of a device group (std::map nodes do not move)
RasterScratch& scratch_of(ogrecu_ctx* ctx)
{
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    return scratch_map()[ctx];
}

int exclusive_scan_u32(ogrecu_ctx* ctx, Scratch& sums, uint32_t* data, size_t n, uint64_t** total_dev)
{
    // in place; afterwards (*total_dev)[0] holds the grand total on the device
    size_t nblocks = div_up(n, SCAN_BLOCK);
    OGRECU_CHECK(scratch_reserve(ctx, sums, (nblocks + 1) * sizeof(uint64_t)));
    uint64_t* bs = (uint64_t*)sums.ptr;
    scan_reduce_kernel<<<(unsigned)nblocks, SCAN_THREADS, 0, ctx->stream>>>(data, bs, n);
    scan_sums_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(bs, nblocks);
    scan_apply_kernel<<<(unsigned)nblocks, SCAN_THREADS, 0, ctx->stream>>>(data, data, bs, n);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx, 3);
    *total_dev = bs + nblocks;
    return OGRECU_OK;
}

uint32_t tri_count_of(const ogrecu_draw_desc& d)
{
    if (d.prim_type == OGRECU_PRIM_TRIANGLE_STRIP) return d.draw_count >= 3 ? d.draw_count - 2 : 0;
    const uint64_t per = d.draw_count / 3;   // the reference reads past the end when drawCount%3 != 0 (UB); we stop
    const uint64_t all = per * (d.instance_count ? d.instance_count : 1);
    return all > 0xffffffffull ? 0xffffffffu : (uint32_t)all;
}

struct Item { const ogrecu_draw_desc* desc; uint32_t target; const float* mvp; const float* mvit; const float* light; const ogrecu_draw_state* state; };

// Run the pipeline for `items` (draws in submission order) over `targets` (same size).
// One pipeline pass over `items` on `ctx`, in two host phases so that the devices of a group overlap: segment_begin enqueues
// everything up to the count pass and the read-back of its two totals, segment_finish waits for them, sizes the record and
// pair arrays and enqueues the rest.
struct SegRun {
    ogrecu_ctx* ctx = nullptr;
    std::vector<ogrecu_rt*> targets;
    SegParams P;
    uint32_t* ctl_host = nullptr;
    uint32_t nblocks = 0, nkeys = 0;
    int tab_slot = 0;           // pinned table block of this pass (its event is recorded once the totals have been copied back)
    bool live = false;          // segment_finish has work to do
};

int segment_begin(ogrecu_ctx* ctx, ogrecu_rt* const* targets, uint32_t ntargets, const std::vector<Item>& items, SegRun& run)
{
    run.live = false;
    run.ctx = ctx;
    run.targets.assign(targets, targets + ntargets);
    if (items.empty()) return OGRECU_OK;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    RasterScratch& S = scratch_of(ctx);
    ogrecu_rt* rt0 = targets[0];
    const uint32_t W = rt0->w, H = rt0->h;
    const uint32_t tiles_x = div_up(W, TILE), tiles_y = div_up(H, TILE);
    const uint32_t strip_ty0 = rt0->strip_y0 / TILE, strip_ty1 = div_up(std::min(rt0->strip_y1, H), TILE);
    if (strip_ty0 >= strip_ty1) return OGRECU_OK;

    // ---- per-draw device table
    std::vector<DrawDev> dd(items.size());
    uint64_t ntris64 = 0, nverts64 = 0;
    for (size_t k = 0; k < items.size(); ++k) {
        const ogrecu_draw_desc& d = *items[k].desc;
        DrawDev& o = dd[k];
        memset(&o, 0, sizeof o);
        o.pos = (const uint8_t*)d.positions; o.uv = (const uint8_t*)d.uvs; o.nrm = (const uint8_t*)d.normals; o.idx = d.indices;
        o.pos_stride = d.pos_stride; o.uv_stride = d.uv_stride; o.nrm_stride = d.normal_stride; o.index_type = d.index_type;
        o.is_strip = d.prim_type == OGRECU_PRIM_TRIANGLE_STRIP;
        o.flags = d.flags;
        if (!d.normals) o.flags &= ~(uint32_t)OGRECU_DRAW_LIGHTING;    // uniform_doLighting &= hasNormals (.cpp:369)
        o.target = items[k].target;
        o.tri_base = (uint32_t)ntris64;
        o.tri_count = tri_count_of(d);
        o.inst_tris = d.instance_count ? d.draw_count / 3 : 0;
        o.inst_verts = d.instance_vertex_count;
        o.inst_magic = o.inst_tris ? (uint32_t)std::min<uint64_t>((1ull << 32) / o.inst_tris, 0xffffffffull) : 0;
        ntris64 += o.tri_count;
        // vertex pre-transform when the caller bounded the vertex range and it is not mostly unused
        o.vert_base = (uint32_t)nverts64;
        if (d.vertex_count && (uint64_t)d.vertex_count <= 4ull * o.tri_count + 64 && ctx->raster_xf != 0) {
            o.vertex_count = d.vertex_count;
            nverts64 += d.vertex_count;
        }
        if (d.texture) {
            OGRECU_CHECK(replica_tex(ctx, const_cast<ogrecu_tex*>(d.texture), &o.tex));                 // a peer of a device group samples its own replica
            o.tex_w = d.texture->w; o.tex_h = d.texture->h; o.tex_zero_alpha = d.texture->has_zero_alpha;
        }
        memcpy(o.mvp, items[k].mvp ? items[k].mvp : d.mvp, 64);
        memcpy(o.texm, d.tex_matrix, 64);
        memcpy(o.vp, d.viewport, 64);
        const float* mv = items[k].mvit ? items[k].mvit : d.mvit;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o.mvit[3 * r + c] = mv[4 * r + c];   // Matrix4::linear()
        memcpy(o.light, items[k].light ? items[k].light : d.light_dir, 12);
        memcpy(o.ambient, d.ambient, 16);
        if (const ogrecu_draw_state* st = items[k].state) if (st->enable) {
            o.ext = (st->enable & 0x1fu) | (st->depth_func & 0xfu) << 8 | (st->cull_mode & 0xfu) << 12 | (st->alpha_func & 0xfu) << 16 |
                    (st->alpha_ref & 0xffu) << 24;
            if (st->enable & OGRECU_EXT_SCISSOR) {
                // an empty rectangle becomes left > right: every bbox is rejected
                const uint32_t l = std::min(st->scissor[0], 0xfffeu), tp = std::min(st->scissor[1], 0xfffeu);
                const uint32_t r = std::min(st->scissor[2], W), b = std::min(st->scissor[3], H);
                if (r <= l || b <= tp) { o.sc_x = 1u; o.sc_y = 1u; }               // left 1, right 0
                else { o.sc_x = l | (r - 1) << 16; o.sc_y = tp | (b - 1) << 16; }
            }
            if (st->enable & OGRECU_EXT_VERTEX_COLOUR) { o.col = (const uint8_t*)st->colours; o.col_stride = st->colour_stride; }
        }
    }
    OGRECU_REQUIRE(ntris64 < 0x7fffff00ull, "too many triangles in one flush");
    OGRECU_REQUIRE(nverts64 < 0x7fffff00ull, "too many vertices in one flush");
    const uint32_t ntris = (uint32_t)ntris64;
    for (uint32_t t = 0; t < ntargets; ++t) { targets[t]->stats.triangles_in += 0; }
    if (ntris == 0) return OGRECU_OK;
    const uint32_t nkeys = ntargets * tiles_x * tiles_y;

    std::vector<TargetDev> td(ntargets);
    for (uint32_t t = 0; t < ntargets; ++t) { td[t].colour = targets[t]->colour; td[t].depth = targets[t]->depth; }

    OGRECU_CHECK(scratch_reserve(ctx, S.draws, dd.size() * sizeof(DrawDev)));
    OGRECU_CHECK(scratch_reserve(ctx, S.targets, td.size() * sizeof(TargetDev)));
    const uint32_t nverts = (uint32_t)nverts64;
    if (nverts) {
        OGRECU_CHECK(scratch_reserve(ctx, S.xf, (size_t)nverts * sizeof(float4)));
        for (DrawDev& o : dd) if (o.vertex_count) o.xf = (const float4*)S.xf.ptr + o.vert_base;
    }
    const uint32_t nblocks = div_up(ntris, SB_BLOCK);
    OGRECU_CHECK(scratch_reserve(ctx, S.alive, (size_t)ntris));
    // control block: block aggregates [nblocks][2] u32 | bases [nblocks][2] u64 | ctl[4] | tile flags [nkeys] (flags zeroed per pass)
    const size_t agg_bytes = ((size_t)nblocks * 8 + 15) & ~(size_t)15, base_bytes = (size_t)nblocks * 16;
    const size_t ctl_bytes = agg_bytes + base_bytes + 16 + (size_t)nkeys * 4;
    OGRECU_CHECK(scratch_reserve(ctx, S.cnt, ctl_bytes));
    OGRECU_CHECK(scratch_reserve(ctx, S.tile_count, ((size_t)nkeys + 1) * sizeof(uint32_t)));
    // tables go through the pinned block (stream-ordered, no pageable-copy stall)
    size_t tab_bytes = dd.size() * sizeof(DrawDev) + td.size() * sizeof(TargetDev) + 16;
    const int slot = ctx->tab_slot;
    ctx->tab_slot ^= 1;
    if (ctx->tab_used[slot]) OGRECU_CUDA(cudaEventSynchronize(ctx->tab_ev[slot]));      // the pass before last has read its tables (long ago)
    if (ctx->tab_cap[slot] < tab_bytes) {
        if (ctx->tab_pin[slot]) cudaFreeHost(ctx->tab_pin[slot]);
        ctx->tab_pin[slot] = nullptr; ctx->tab_cap[slot] = 0;
        const size_t cap = std::max<size_t>(2 * tab_bytes, (size_t)1 << 16);
        OGRECU_CUDA(cudaHostAlloc((void**)&ctx->tab_pin[slot], cap, cudaHostAllocDefault));
        ctx->tab_cap[slot] = cap;
    }
    if (!ctx->tab_ev[slot]) OGRECU_CUDA(cudaEventCreateWithFlags(&ctx->tab_ev[slot], cudaEventDisableTiming));
    uint8_t* tab = ctx->tab_pin[slot];
    memcpy(tab, dd.data(), dd.size() * sizeof(DrawDev));
    memcpy(tab + dd.size() * sizeof(DrawDev), td.data(), td.size() * sizeof(TargetDev));
    OGRECU_CUDA(cudaMemcpyAsync(S.draws.ptr, tab, dd.size() * sizeof(DrawDev), cudaMemcpyHostToDevice, ctx->stream));
    OGRECU_CUDA(cudaMemcpyAsync(S.targets.ptr, tab + dd.size() * sizeof(DrawDev), td.size() * sizeof(TargetDev), cudaMemcpyHostToDevice, ctx->stream));
    uint32_t* ctl_host = (uint32_t*)(tab + dd.size() * sizeof(DrawDev) + td.size() * sizeof(TargetDev));
    run.tab_slot = slot;

    SegParams P;
    P.draws = (const DrawDev*)S.draws.ptr; P.ndraws = (uint32_t)dd.size(); P.ntris = ntris;
    P.rec = nullptr;                 // sized after the count pass (survivors only)
    P.alive = (uint8_t*)S.alive.ptr;
    P.block_agg = (uint32_t*)S.cnt.ptr;
    P.bases = (unsigned long long*)((uint8_t*)S.cnt.ptr + agg_bytes);
    P.ctl = (uint32_t*)((uint8_t*)S.cnt.ptr + agg_bytes + base_bytes);
    P.tile_flags = P.ctl + 4;
    P.pairs = nullptr; P.pair_capacity = 0;
    P.width = W; P.height = H; P.tiles_x = tiles_x; P.tiles_y = tiles_y;
    P.strip_ty0 = strip_ty0; P.strip_ty1 = strip_ty1;
    P.band_tiles = rt0->band_tiles; P.band_phase = rt0->band_phase; P.band_period = rt0->band_period;

    // ---- K0: every vertex of the flush transformed once
    if (nverts) {
        XfParams X;
        X.draws = P.draws; X.ndraws = P.ndraws; X.nverts = nverts; X.out = (float4*)S.xf.ptr;
        vertex_xf_kernel<<<div_up(nverts, 256), 256, 0, ctx->stream>>>(X);
        OGRECU_CUDA(cudaGetLastError());
        count_launch(ctx);
    }
    // ---- K1a/K1b: count survivors and pairs per CTA, scan the CTA table
    OGRECU_CUDA(cudaMemsetAsync(P.ctl, 0, 16 + (size_t)nkeys * 4, ctx->stream));
    setup_count_kernel<<<nblocks, SB_THREADS, 0, ctx->stream>>>(P);
    setup_scan_kernel<<<div_up(nblocks, (uint32_t)SS_CHUNK), 1024, 0, ctx->stream>>>(P.block_agg, nblocks, P.ctl, P.bases);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx, 2);
    OGRECU_CUDA(cudaMemcpyAsync(ctl_host, P.ctl, 16, cudaMemcpyDeviceToHost, ctx->stream));
    OGRECU_CUDA(cudaEventRecord(ctx->tab_ev[run.tab_slot], ctx->stream));
    ctx->tab_used[run.tab_slot] = true;
    run.P = P; run.ctl_host = ctl_host; run.nblocks = nblocks; run.nkeys = nkeys;
    run.live = true;
    return OGRECU_OK;
}

int segment_finish(SegRun& run)
{
    if (!run.live) return OGRECU_OK;
    run.live = false;
    ogrecu_ctx* ctx = run.ctx;
    ogrecu_rt* const* targets = run.targets.data();
    const uint32_t ntargets = (uint32_t)run.targets.size();
    ogrecu_rt* rt0 = targets[0];
    SegParams& P = run.P;
    uint32_t* ctl_host = run.ctl_host;
    const uint32_t nblocks = run.nblocks, nkeys = run.nkeys, ntris = P.ntris;
    const uint32_t W = P.width, H = P.height, tiles_x = P.tiles_x, tiles_y = P.tiles_y, strip_ty0 = P.strip_ty0, strip_ty1 = P.strip_ty1;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    RasterScratch& S = scratch_of(ctx);
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint64_t npairs = ctl_host[2], nsurv = ctl_host[3];
    OGRECU_REQUIRE(npairs < 0xffffffffull, "too many (triangle, tile) pairs in one flush");
    for (uint32_t t = 0; t < ntargets; ++t) {
        targets[t]->stats.triangles_in = ntris / ntargets;
        targets[t]->stats.triangles_binned = nsurv / ntargets;
        targets[t]->stats.tile_pairs = npairs / ntargets;
        targets[t]->stats.tiles_visibility = targets[t]->stats.tiles_ordered = targets[t]->stats.tiles_active = 0;
        targets[t]->stats_pending = false;
    }
    if (npairs == 0) return OGRECU_OK;

    // ---- K1c: compact records + (tile, triangle) pairs in submission order
    OGRECU_CHECK(scratch_reserve(ctx, S.rec, (size_t)(nsurv + SB_THREADS) * sizeof(TriRec)));    // a round's copy-out runs in whole rows
    P.rec = (TriRec*)S.rec.ptr;
    OGRECU_CHECK(scratch_reserve(ctx, S.pairs_a, npairs * sizeof(uint64_t)));
    OGRECU_CHECK(scratch_reserve(ctx, S.pairs_b, npairs * sizeof(uint64_t)));
    P.pairs = (uint64_t*)S.pairs_a.ptr;
    P.pair_capacity = npairs;
    setup_write_kernel<<<nblocks, SB_THREADS, 0, ctx->stream>>>(P);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);

    // ---- stable sort by tile key
    uint64_t* pa = (uint64_t*)S.pairs_a.ptr;
    uint64_t* pb = (uint64_t*)S.pairs_b.ptr;
    int key_bits = 1;
    while ((1ull << key_bits) < nkeys) ++key_bits;
    const uint32_t rs_blocks = div_up(npairs, RS_BLOCK);
    OGRECU_CHECK(scratch_reserve(ctx, S.hist, ((size_t)RS_BINS * rs_blocks + 1) * sizeof(uint32_t)));
    uint32_t* hist = (uint32_t*)S.hist.ptr;
    for (int shift = 0; shift < key_bits; shift += 8) {
        radix_hist_kernel<<<rs_blocks, RS_THREADS, 0, ctx->stream>>>(pa, hist, npairs, 32 + shift, rs_blocks);
        OGRECU_CUDA(cudaGetLastError());
        count_launch(ctx);
        uint64_t* dummy;
        OGRECU_CHECK(exclusive_scan_u32(ctx, S.sums, hist, (size_t)RS_BINS * rs_blocks, &dummy));
        radix_scatter_kernel<<<rs_blocks, RS_THREADS, 0, ctx->stream>>>(pa, pb, hist, npairs, 32 + shift, rs_blocks);
        OGRECU_CUDA(cudaGetLastError());
        count_launch(ctx);
        std::swap(pa, pb);
    }
    uint32_t* tile_start = (uint32_t*)S.tile_count.ptr;
    tile_bounds_kernel<<<div_up((size_t)nkeys + 1, 256), 256, 0, ctx->stream>>>(pa, (uint32_t)npairs, nkeys, tile_start);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);

    // ---- K3
    RasterParams R;
    R.draws = P.draws; R.targets = (const TargetDev*)S.targets.ptr; R.rec = P.rec;
    R.pairs = pa; R.tile_start = tile_start; R.tile_flags = P.tile_flags;
    R.force_ordered = ctx->raster_mode == 1;
    for (ogrecu_rt* other : ctx->targets) other->stats_pending = false;   // the counter table is per context: last flush wins
    OGRECU_CHECK(scratch_reserve(ctx, S.counters, (size_t)ntargets * 2 * sizeof(uint32_t)));
    OGRECU_CUDA(cudaMemsetAsync(S.counters.ptr, 0, (size_t)ntargets * 2 * sizeof(uint32_t), ctx->stream));
    R.counters = (uint32_t*)S.counters.ptr;
    R.width = W; R.height = H; R.tiles_x = tiles_x; R.tiles_y = tiles_y;
    R.strip_ty0 = strip_ty0; R.strip_rows = strip_ty1 - strip_ty0;
    R.band_tiles = rt0->band_tiles; R.band_phase = rt0->band_phase; R.band_period = rt0->band_period;
    if (R.band_period > 1) {                     // rows of the bands owned here (the last band may be cut by the frame: those CTAs exit)
        const uint32_t nbands = div_up(tiles_y, R.band_tiles);
        const uint32_t mine = nbands > R.band_phase ? (nbands - R.band_phase + R.band_period - 1) / R.band_period : 0;
        R.strip_rows = mine * R.band_tiles;
        if (R.strip_rows == 0) return OGRECU_OK;
    }
    const unsigned grid = ntargets * tiles_x * R.strip_rows;
    if (W % 16 == 0) raster_kernel<true><<<grid, RT_THREADS, 0, ctx->stream>>>(R);
    else raster_kernel<false><<<grid, RT_THREADS, 0, ctx->stream>>>(R);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    // per-target path counters: fetched (synchronously) only if ogrecu_rt_last_stats asks
    for (uint32_t t = 0; t < ntargets; ++t) { targets[t]->stats_slot = (int)t; targets[t]->stats_pending = true; }
    return OGRECU_OK;
}

int run_segment(ogrecu_ctx* ctx, ogrecu_rt* const* targets, uint32_t ntargets, const std::vector<Item>& items)
{
    SegRun run;
    OGRECU_CHECK(segment_begin(ctx, targets, ntargets, items, run));
    return segment_finish(run);
}

// The draw descriptions of a segment as device `dev` reads them: pointers into ogrecu_bufs become the peer's replicas (refreshed
// on its stream when stale), textures are replaced inside segment_begin, anything else is read in place over NVLink.
struct PeerSeg {
    std::vector<ogrecu_draw_desc> descs;
    std::vector<ogrecu_draw_state> states;
    std::vector<Item> items;
};
int translate_segment(ogrecu_ctx* dev, const std::vector<Item>& seg, PeerSeg& out)
{
    out.descs.resize(seg.size()); out.states.resize(seg.size()); out.items = seg;
    for (size_t i = 0; i < seg.size(); ++i) {
        ogrecu_draw_desc& d = out.descs[i];
        d = *seg[i].desc;
        OGRECU_CHECK(replica_ptr(dev, d.positions, &d.positions));
        OGRECU_CHECK(replica_ptr(dev, d.uvs, &d.uvs));
        OGRECU_CHECK(replica_ptr(dev, d.normals, &d.normals));
        OGRECU_CHECK(replica_ptr(dev, d.indices, &d.indices));
        out.items[i].desc = &d;
        if (seg[i].state) {
            out.states[i] = *seg[i].state;
            if (out.states[i].enable & OGRECU_EXT_VERTEX_COLOUR) OGRECU_CHECK(replica_ptr(dev, out.states[i].colours, &out.states[i].colours));
            out.items[i].state = &out.states[i];
        }
    }
    return OGRECU_OK;
}


// ---- sort-last form of a split segment (see merge_sort_last_kernel)
bool sort_last_eligible(ogrecu_ctx* root, const std::vector<Item>& seg)
{
    if (root->split_mode != 0) return false;
    uint64_t tris = 0;
    for (const Item& it : seg) {
        const ogrecu_draw_desc& d = *it.desc;
        if (d.prim_type != OGRECU_PRIM_TRIANGLE_LIST || d.instance_count) return false;
        if ((d.flags & (OGRECU_DRAW_DEPTH_TEST | OGRECU_DRAW_DEPTH_WRITE | OGRECU_DRAW_BLEND_ADD)) != (OGRECU_DRAW_DEPTH_TEST | OGRECU_DRAW_DEPTH_WRITE)) return false;
        if (d.texture && d.texture->has_zero_alpha) return false;
        if (const ogrecu_draw_state* st = it.state) {
            if ((st->enable & OGRECU_EXT_DEPTH_FUNC) && st->depth_func != OGRECU_CMPF_LESS_EQUAL) return false;
            if ((st->enable & OGRECU_EXT_ALPHA_REJECT) && st->alpha_func != OGRECU_CMPF_ALWAYS_PASS) return false;
            if ((st->enable & OGRECU_EXT_VERTEX_COLOUR) && d.index_type == OGRECU_INDEX_NONE) return false;     // (its colour stream would have to be sliced too)
        }
        tris += d.draw_count / 3;
    }
    return tris >= 262144;            // below that the private planes' clear and the merge cost more than the split saves
}

int run_segment_sort_last(ogrecu_rt* rt, const std::vector<Item>& seg)
{
    ogrecu_ctx* root = rt->ctx;
    const size_t G = 1 + rt->aliases.size();
    const size_t npx = (size_t)rt->w * rt->h;
    const float inf = std::numeric_limits<float>::infinity();
    std::vector<ogrecu_rt> tmp(G);                   // whole-frame views of the planes each device draws into
    std::vector<std::vector<ogrecu_draw_desc>> descs(G);
    std::vector<std::vector<Item>> items(G);
    std::vector<PeerSeg> tr(G);
    std::vector<SegRun> runs(G);
    int rc = group_parallel(root, G, [&](size_t k) -> int {
        ogrecu_rt* part = k ? rt->aliases[k - 1] : rt;
        ogrecu_ctx* dev = part->ctx;
        OGRECU_CUDA(cudaSetDevice(dev->device));
        tmp[k].ctx = dev; tmp[k].w = rt->w; tmp[k].h = rt->h; tmp[k].strip_y0 = 0; tmp[k].strip_y1 = rt->h;
        tmp[k].colour = rt->colour; tmp[k].depth = rt->depth;
        memset(&tmp[k].stats, 0, sizeof tmp[k].stats); tmp[k].stats_slot = 0; tmp[k].stats_pending = false;
        if (k) {
            if (!part->sl_colour) {
                cudaError_t e = cudaMalloc((void**)&part->sl_colour, npx * 3 + 16);
                if (e == cudaSuccess) e = cudaMalloc((void**)&part->sl_depth, npx * 4 + 16);
                if (e != cudaSuccess) { cudaFree(part->sl_colour); part->sl_colour = nullptr; return cuda_fail(e, "sort-last planes", __FILE__, __LINE__); }
            }
            tmp[k].colour = part->sl_colour; tmp[k].depth = part->sl_depth;
            clear_depth_kernel<<<div_up(div_up(npx, 4), 256), 256, 0, dev->stream>>>(part->sl_depth, npx, inf);
            count_launch(dev);
        }
        // slice k of every draw's triangles
        descs[k].reserve(seg.size());
        for (const Item& it : seg) {
            ogrecu_draw_desc d = *it.desc;
            const uint64_t nt = d.draw_count / 3, lo = nt * k / G, hi = nt * (k + 1) / G;
            if (hi == lo) continue;
            if (d.index_type == OGRECU_INDEX_NONE) {
                d.positions = (const uint8_t*)d.positions + lo * 3 * d.pos_stride;
                if (d.uvs) d.uvs = (const uint8_t*)d.uvs + lo * 3 * d.uv_stride;
                if (d.normals) d.normals = (const uint8_t*)d.normals + lo * 3 * d.normal_stride;
                if (d.vertex_count) d.vertex_count = (uint32_t)((hi - lo) * 3);
            } else d.indices = (const uint8_t*)d.indices + lo * 3 * d.index_type;
            d.draw_count = (uint32_t)((hi - lo) * 3);
            descs[k].push_back(d);
            Item ni = it;
            ni.desc = &descs[k].back();
            items[k].push_back(ni);
        }
        ogrecu_rt* one[1] = {&tmp[k]};
        if (dev->root) {
            OGRECU_CHECK(translate_segment(dev, items[k], tr[k]));
            return segment_begin(dev, one, 1, tr[k].items, runs[k]);
        }
        return segment_begin(dev, one, 1, items[k], runs[k]);
    });
    {
        const int rc_begin = rc;
        const int rc2 = group_parallel(root, G, [&](size_t k) -> int {
            ogrecu_ctx* dev = k ? rt->aliases[k - 1]->ctx : root;
            if (rc_begin == OGRECU_OK) return segment_finish(runs[k]);
            cudaSetDevice(dev->device); cudaStreamSynchronize(dev->stream);
            return OGRECU_OK;
        });
        if (rc == OGRECU_OK) rc = rc2;
    }
    cudaSetDevice(root->device);
    if (rc != OGRECU_OK) return rc;
    rt->stats = tmp[0].stats;
    for (size_t k = 1; k < G; ++k) {
        rt->stats.triangles_in += tmp[k].stats.triangles_in; rt->stats.triangles_binned += tmp[k].stats.triangles_binned;
        rt->stats.tile_pairs += tmp[k].stats.tile_pairs;
    }
    rt->stats_pending = false;
    // every device has to see every slice finished before it merges its share of the pixels
    for (size_t k = 0; k < G; ++k) {
        ogrecu_ctx* dev = k ? rt->aliases[k - 1]->ctx : root;
        OGRECU_CUDA(cudaSetDevice(dev->device));
        OGRECU_CUDA(cudaEventRecord(dev->group_ev, dev->stream));
    }
    MergeSrc S;
    memset(&S, 0, sizeof S);
    S.n = (int)G - 1;
    for (size_t j = 1; j < G; ++j) { S.rgb[j - 1] = rt->aliases[j - 1]->sl_colour; S.z[j - 1] = rt->aliases[j - 1]->sl_depth; }
    for (size_t k = 0; k < G; ++k) {
        ogrecu_ctx* dev = k ? rt->aliases[k - 1]->ctx : root;
        OGRECU_CUDA(cudaSetDevice(dev->device));
        for (size_t j = 0; j < G; ++j) {
            if (j == k) continue;
            ogrecu_ctx* other = j ? rt->aliases[j - 1]->ctx : root;
            OGRECU_CUDA(cudaStreamWaitEvent(dev->stream, other->group_ev, 0));
        }
        const size_t p0 = npx * k / G, p1 = npx * (k + 1) / G;
        if (p1 > p0) {
            merge_sort_last_kernel<<<div_up(p1 - p0, 256), 256, 0, dev->stream>>>(rt->colour, rt->depth, S, p0, p1);
            OGRECU_CUDA(cudaGetLastError());
            count_launch(dev);
        }
    }
    // the next pass on any device (a later segment, the private planes' next clear) comes after every merge
    for (size_t k = 0; k < G; ++k) {
        ogrecu_ctx* dev = k ? rt->aliases[k - 1]->ctx : root;
        OGRECU_CUDA(cudaSetDevice(dev->device));
        OGRECU_CUDA(cudaEventRecord(dev->group_ev, dev->stream));
    }
    for (size_t k = 0; k < G; ++k) {
        ogrecu_ctx* dev = k ? rt->aliases[k - 1]->ctx : root;
        OGRECU_CUDA(cudaSetDevice(dev->device));
        for (size_t j = 0; j < G; ++j) {
            if (j == k) continue;
            ogrecu_ctx* other = j ? rt->aliases[j - 1]->ctx : root;
            OGRECU_CUDA(cudaStreamWaitEvent(dev->stream, other->group_ev, 0));
        }
    }
    OGRECU_CUDA(cudaSetDevice(root->device));
    return OGRECU_OK;
}

// One segment of one target on every device that takes part in it: the target's own context, or — for a split target — the
// root and all peers, each on the bands it owns.  Both host phases run device by device, so the devices' passes overlap.
int run_segment_group(ogrecu_rt* rt, const std::vector<Item>& seg)
{
    if (seg.empty()) return OGRECU_OK;
    ogrecu_ctx* ctx = rt->ctx;
    if (!rt->split && !ctx->root) { ogrecu_rt* one[1] = {rt}; return run_segment(ctx, one, 1, seg); }
    if (rt->split && sort_last_eligible(ctx, seg)) return run_segment_sort_last(rt, seg);
    std::vector<ogrecu_rt*> parts;
    parts.push_back(rt);
    for (ogrecu_rt* a : rt->aliases) parts.push_back(a);
    std::vector<SegRun> runs(parts.size());
    std::vector<PeerSeg> tr(parts.size());
    ogrecu_ctx* root = root_of(ctx);
    // both host phases run on one thread per device (group_parallel), so the devices' passes are submitted side by side
    int rc = group_parallel(root, parts.size(), [&](size_t k) -> int {
        ogrecu_ctx* dev = parts[k]->ctx;
        ogrecu_rt* one[1] = {parts[k]};
        if (dev->root) {
            OGRECU_CHECK(translate_segment(dev, seg, tr[k]));
            return segment_begin(dev, one, 1, tr[k].items, runs[k]);
        }
        return segment_begin(dev, one, 1, seg, runs[k]);
    });
    const int rc_begin = rc;
    const int rc2 = group_parallel(root, parts.size(), [&](size_t k) -> int {      // every begun pass is finished (or at least waited for) on every path
        if (rc_begin == OGRECU_OK) return segment_finish(runs[k]);
        cudaSetDevice(parts[k]->ctx->device); cudaStreamSynchronize(parts[k]->ctx->stream);
        return OGRECU_OK;
    });
    if (rc == OGRECU_OK) rc = rc2;
    cudaSetDevice(ctx->device);
    if (rt->split) {                                            // the statistics of a split target: the sum over its devices
        for (ogrecu_rt* a : rt->aliases) { rt->stats.triangles_binned += a->stats.triangles_binned; rt->stats.tile_pairs += a->stats.tile_pairs; }
    }
    return rc;
}

// `handshake` = false: the caller has already ordered the target's device after the root's stream and will order the root after it
// (ogrecu_draw_views does that once per device for all of its targets instead of once per target)
int rt_flush_cmds(ogrecu_rt* rt, bool handshake = true)
{
    ogrecu_ctx* ctx = rt->ctx;                       // where the planes are cleared / read: the root, or the peer a target was placed on
    ogrecu_ctx* root = root_of(ctx);
    if (rt->cmds.empty() && root->batch_jobs.empty()) return OGRECU_OK;
    ProfileScope range("ogrecu: target flush");
    OGRECU_CUDA(cudaSetDevice(root->device));
    OGRECU_CHECK(flush_skin_batch(root));     // queued OptimisedUtil jobs produce what the draws read
    if (rt->cmds.empty()) return OGRECU_OK;
    // device group: the peers that take part start after everything the root's stream holds so far (uploads, OptimisedUtil results)
    std::vector<ogrecu_ctx*> others;
    if (ctx->root && handshake) others.push_back(ctx);
    for (ogrecu_rt* a : rt->aliases) others.push_back(a->ctx);
    for (ogrecu_ctx* o : others) OGRECU_CHECK(stream_wait(o, root));
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    std::vector<Item> seg;
    // The list is consumed on EVERY exit path: a flush that fails half way (out of memory, too many pairs, a CUDA error)
    // reports once and leaves nothing behind to be replayed — segments that already ran must not run again (additive
    // blending would apply twice) and the failing draw must not fail on every later lock / read-back.
    std::vector<Cmd> cmds;
    cmds.swap(rt->cmds);
    int rc = OGRECU_OK;
    for (size_t i = 0; i < cmds.size() && rc == OGRECU_OK; ++i) {
        Cmd& c = cmds[i];
        if (c.kind == 0) { seg.push_back({&c.draw, 0, nullptr, nullptr, nullptr, &c.state}); continue; }
        rc = run_segment_group(rt, seg);
        if (rc != OGRECU_OK) break;
        const bool had_draws = !seg.empty();
        seg.clear();
        // a clear runs where the planes live; on a split target it is ordered after the peers' earlier passes and before their later ones
        if (rt->split && had_draws) for (ogrecu_rt* a : rt->aliases) if ((rc = stream_wait(ctx, a->ctx)) != OGRECU_OK) break;
        if (rc != OGRECU_OK) break;
        OGRECU_CUDA(cudaSetDevice(ctx->device));
        const size_t npx = (size_t)rt->w * rt->h;
        if (c.clear_buffers & OGRECU_CLEAR_COLOUR) {
            const float* col = c.clear_colour;
            if (col[0] == 0.f && col[1] == 0.f && col[2] == 0.f && col[3] == 0.f) {
                OGRECU_CUDA(cudaMemsetAsync(rt->colour, 0, npx * 3, ctx->stream));     // Image::setTo ZERO shortcut (OgreImage.cpp:124-128)
            } else {
                uint32_t q[3];
                for (int k = 0; k < 3; ++k) {   // Bitwise::floatToFixed(v, 8), OgreBitwise.h:237-242
                    float v = col[k];
                    q[k] = (v <= 0.0f) ? 0u : (v >= 1.0f) ? 255u : (uint32_t)(v * 256.0f);
                    q[k] &= 0xffu;
                }
                clear_colour_kernel<<<div_up(div_up(npx * 3, 16), 256), 256, 0, ctx->stream>>>(rt->colour, npx * 3, q[0], q[1], q[2]);
                OGRECU_CUDA(cudaGetLastError());
                count_launch(ctx);
            }
        }
        if (c.clear_buffers & OGRECU_CLEAR_DEPTH) {
            clear_depth_kernel<<<div_up(div_up(npx, 4), 256), 256, 0, ctx->stream>>>(rt->depth, npx, c.clear_depth);
            OGRECU_CUDA(cudaGetLastError());
            count_launch(ctx);
        }
        if (rt->split) for (ogrecu_rt* a : rt->aliases) if ((rc = stream_wait(a->ctx, ctx)) != OGRECU_OK) break;
    }
    if (rc == OGRECU_OK) rc = run_segment_group(rt, seg);
    // the root goes on (read-back, the next frame's uploads and OptimisedUtil calls into the buffers the peers copied from) only after
    // the peers are through; a placed target's own stream already orders its read-back
    for (ogrecu_ctx* o : others) { const int r = stream_wait(root, o); if (rc == OGRECU_OK) rc = r; }
    cudaSetDevice(root->device);
    return rc;
}

}  // namespace

namespace ogrecu {

int flush_all_targets(ogrecu_ctx* ctx)
{
    int rc = OGRECU_OK;
    ctx = root_of(ctx);
    std::vector<ogrecu_rt*> targets(ctx->targets);
    for (ogrecu_ctx* p : ctx->peers) targets.insert(targets.end(), p->targets.begin(), p->targets.end());     // targets placed on peers
    for (ogrecu_rt* t : targets)
        if (!t->cmds.empty()) { const int r = rt_flush_cmds(t); if (rc == OGRECU_OK) rc = r; }
    ++ctx->draw_epoch;                     // no recorded draw is left: every buffer's draw_ref mark is stale now
    return rc;
}
void free_raster_scratch(ogrecu_ctx* ctx)
{
    auto sm = scratch_map().find(ctx);
    if (sm == scratch_map().end()) return;
    cudaSetDevice(ctx->device);
    RasterScratch& S = sm->second;
    for (Scratch* s : {&S.draws, &S.targets, &S.rec, &S.cnt, &S.sums, &S.pairs_a, &S.pairs_b, &S.tile_count, &S.hist, &S.counters, &S.xf, &S.alive})
        if (s->ptr) cudaFree(s->ptr);
    scratch_map().erase(sm);
}
int rt_release(ogrecu_rt* rt)
{
    if (!rt) return OGRECU_OK;
    ogrecu_ctx* ctx = rt->ctx;
    for (ogrecu_rt* a : rt->aliases) {                 // the peers' views of a split target's planes
        cudaSetDevice(a->ctx->device);
        cudaStreamSynchronize(a->ctx->stream);
        cudaFree(a->sl_colour);
        cudaFree(a->sl_depth);
        delete a;
    }
    rt->aliases.clear();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(rt->colour);
    cudaFree(rt->depth);
    auto it = std::find(ctx->targets.begin(), ctx->targets.end(), rt);
    if (it != ctx->targets.end()) ctx->targets.erase(it);
    if (ctx->targets.empty() && ctx->root == nullptr && ctx->peers.empty()) free_raster_scratch(ctx);      // (a group keeps its arenas until shutdown)
    delete rt;
    return OGRECU_OK;
}
}  // namespace ogrecu

extern "C" {

/* ---- textures ---- */

int ogrecu_tex_create(ogrecu_ctx* ctx, uint32_t w, uint32_t h, ogrecu_tex** out)
{
    OGRECU_REQUIRE(ctx && out, "NULL argument");
    OGRECU_REQUIRE(w >= 1 && h >= 1 && w <= 32768 && h <= 32768, "texture size out of range");
    *out = nullptr;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    ogrecu_tex* t = new ogrecu_tex();
    memset(t, 0, sizeof *t);
    t->ctx = ctx; t->w = w; t->h = h; t->has_zero_alpha = 0; t->dev = nullptr;
    cudaError_t e = cudaMalloc((void**)&t->dev, (size_t)w * h * 4);
    if (e != cudaSuccess) { delete t; return cuda_fail(e, "cudaMalloc(texture)", __FILE__, __LINE__); }
    e = cudaMemsetAsync(t->dev, 0xff, (size_t)w * h * 4, ctx->stream);
    if (e != cudaSuccess) { cudaFree(t->dev); delete t; return cuda_fail(e, "cudaMemset(texture)", __FILE__, __LINE__); }
    for (size_t k = 0; k < ctx->peers.size(); ++k) {              // device group: a replica per peer, refreshed when a draw there samples it
        cudaSetDevice(ctx->peers[k]->device);
        e = cudaMalloc((void**)&t->peer_dev[k], (size_t)w * h * 4);
        if (e != cudaSuccess) {
            for (size_t q = 0; q < k; ++q) { cudaSetDevice(ctx->peers[q]->device); cudaFree(t->peer_dev[q]); }
            cudaSetDevice(ctx->device); cudaFree(t->dev); delete t;
            return cuda_fail(e, "cudaMalloc(texture replica)", __FILE__, __LINE__);
        }
        t->peer_stale[k] = true;
    }
    if (!ctx->peers.empty()) cudaSetDevice(ctx->device);
    *out = t;
    return OGRECU_OK;
}

int ogrecu_tex_destroy(ogrecu_tex* t)
{
    if (!t) return OGRECU_OK;
    OGRECU_CUDA(cudaSetDevice(t->ctx->device));
    OGRECU_CUDA(cudaStreamSynchronize(t->ctx->stream));
    for (size_t k = 0; k < t->ctx->peers.size(); ++k)
        if (t->peer_dev[k]) { cudaSetDevice(t->ctx->peers[k]->device); cudaStreamSynchronize(t->ctx->peers[k]->stream); cudaFree(t->peer_dev[k]); }
    cudaSetDevice(t->ctx->device);
    cudaFree(t->dev);
    delete t;
    return OGRECU_OK;
}

int ogrecu_tex_upload(ogrecu_tex* t, const uint8_t* rgba, size_t pitch)
{
    OGRECU_REQUIRE(t && rgba, "NULL argument");
    OGRECU_REQUIRE(pitch >= (size_t)t->w * 4, "row pitch smaller than a row");
    ogrecu_ctx* ctx = t->ctx;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    const size_t row = (size_t)t->w * 4, bytes = row * t->h;
    OGRECU_CHECK(pinned_reserve(ctx, bytes));
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    int zero_alpha = 0;
    for (uint32_t y = 0; y < t->h; ++y) {
        const uint8_t* s = rgba + y * pitch;
        memcpy((uint8_t*)ctx->pinned + y * row, s, row);
        if (!zero_alpha) for (uint32_t x = 0; x < t->w; ++x) if (s[4 * x + 3] == 0) { zero_alpha = 1; break; }
    }
    t->has_zero_alpha = zero_alpha;
    OGRECU_CUDA(cudaMemcpyAsync(t->dev, ctx->pinned, bytes, cudaMemcpyHostToDevice, ctx->stream));
    for (size_t k = 0; k < ctx->peers.size(); ++k) t->peer_stale[k] = true;
    return OGRECU_OK;
}

int ogrecu_tex_download(ogrecu_tex* t, uint8_t* rgba, size_t pitch)
{
    OGRECU_REQUIRE(t && rgba, "NULL argument");
    OGRECU_REQUIRE(pitch >= (size_t)t->w * 4, "row pitch smaller than a row");
    ogrecu_ctx* ctx = t->ctx;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    const size_t row = (size_t)t->w * 4, bytes = row * t->h;
    OGRECU_CHECK(pinned_reserve(ctx, bytes));
    OGRECU_CUDA(cudaMemcpyAsync(ctx->pinned, t->dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    for (uint32_t y = 0; y < t->h; ++y) memcpy(rgba + y * pitch, (uint8_t*)ctx->pinned + y * row, row);
    return OGRECU_OK;
}

int ogrecu_tex_copy_from_rt(ogrecu_tex* t, ogrecu_rt* rt)
{
    OGRECU_REQUIRE(t && rt, "NULL argument");
    OGRECU_REQUIRE(t->ctx == root_of(rt->ctx), "texture and target belong to different contexts");
    OGRECU_REQUIRE(t->w == rt->w && t->h == rt->h, "texture and target sizes differ");
    OGRECU_CHECK(ogrecu_flush(rt));
    ogrecu_ctx* ctx = rt->ctx;                         // where the planes live (a peer for a placed target: it writes the root's texels over NVLink)
    if (ctx != t->ctx) OGRECU_CHECK(stream_wait(ctx, t->ctx));
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)t->w * t->h;
    rgb_to_rgba_kernel<<<div_up(div_up(npx, 4), 256), 256, 0, ctx->stream>>>(rt->colour, t->dev, npx);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    if (ctx != t->ctx) { OGRECU_CHECK(stream_wait(t->ctx, ctx)); OGRECU_CUDA(cudaSetDevice(t->ctx->device)); }
    for (size_t k = 0; k < t->ctx->peers.size(); ++k) t->peer_stale[k] = true;
    t->has_zero_alpha = 0;
    return OGRECU_OK;
}

int ogrecu_tex_device_ptr(ogrecu_tex* t, void** dev_ptr)
{
    OGRECU_REQUIRE(t && dev_ptr, "NULL argument");
    *dev_ptr = t->dev;
    return OGRECU_OK;
}

/* ---- render targets ---- */

int ogrecu_rt_create_placed(ogrecu_ctx* ctx, uint32_t w, uint32_t h, int placement, ogrecu_rt** out)
{
    OGRECU_REQUIRE(ctx && out, "NULL argument");
    OGRECU_REQUIRE(w >= 1 && h >= 1 && w <= 32768 && h <= 32768, "render target size out of range");
    *out = nullptr;
    ogrecu_ctx* root = root_of(ctx);
    const int ndev = 1 + (int)root->peers.size();
    OGRECU_REQUIRE(placement == OGRECU_PLACE_SPLIT || (placement >= 0 && placement < ndev), "placement: a device index of the group or OGRECU_PLACE_SPLIT");
    ogrecu_ctx* home = placement > 0 ? root->peers[placement - 1] : root;       // where the planes live
    OGRECU_CUDA(cudaSetDevice(home->device));
    ogrecu_rt* rt = new ogrecu_rt();
    rt->ctx = home; rt->w = w; rt->h = h; rt->colour = nullptr; rt->depth = nullptr;
    rt->strip_y0 = 0; rt->strip_y1 = h;
    memset(&rt->stats, 0, sizeof rt->stats);
    rt->stats_slot = 0; rt->stats_pending = false;
    const size_t npx = (size_t)w * h;
    cudaError_t e = cudaMalloc((void**)&rt->colour, npx * 3 + 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&rt->depth, npx * 4 + 16);
    if (e == cudaSuccess) e = cudaMemsetAsync(rt->colour, 0, npx * 3 + 16, home->stream);    // Image::create zero-fills
    if (e == cudaSuccess) e = cudaMemsetAsync(rt->depth, 0, npx * 4 + 16, home->stream);
    if (e != cudaSuccess) { cudaFree(rt->colour); cudaFree(rt->depth); delete rt; cudaSetDevice(root->device); return cuda_fail(e, "render target allocation", __FILE__, __LINE__); }
    if (placement == OGRECU_PLACE_SPLIT && ndev > 1) {
        // interleaved bands of tile rows, band b -> device b mod ndev; the peers see the root's planes through alias targets
        rt->split = true;
        rt->band_tiles = (uint32_t)root->split_band_tiles; rt->band_phase = 0; rt->band_period = (uint32_t)ndev;
        for (int k = 1; k < ndev; ++k) {
            ogrecu_rt* a = new ogrecu_rt();
            a->ctx = root->peers[k - 1]; a->w = w; a->h = h; a->colour = rt->colour; a->depth = rt->depth;
            a->strip_y0 = 0; a->strip_y1 = h;
            a->band_tiles = rt->band_tiles; a->band_phase = (uint32_t)k; a->band_period = (uint32_t)ndev;
            memset(&a->stats, 0, sizeof a->stats);
            a->stats_slot = 0; a->stats_pending = false;
            rt->aliases.push_back(a);
        }
    }
    home->targets.push_back(rt);
    if (home != root) OGRECU_CUDA(cudaSetDevice(root->device));
    *out = rt;
    return OGRECU_OK;
}

int ogrecu_rt_create(ogrecu_ctx* ctx, uint32_t w, uint32_t h, ogrecu_rt** out)
{
    OGRECU_REQUIRE(ctx, "NULL argument");
    return ogrecu_rt_create_placed(ctx, w, h, ctx->dev_index, out);
}

int ogrecu_rt_destroy(ogrecu_rt* rt) { return rt_release(rt); }

int ogrecu_rt_size(const ogrecu_rt* rt, uint32_t* w, uint32_t* h)
{
    OGRECU_REQUIRE(rt && w && h, "NULL argument");
    *w = rt->w; *h = rt->h;
    return OGRECU_OK;
}

int ogrecu_rt_clear(ogrecu_rt* rt, unsigned buffers, const float colour_rgba[4], float depth)
{
    OGRECU_REQUIRE(rt, "rt is NULL");
    OGRECU_REQUIRE(!(buffers & OGRECU_CLEAR_COLOUR) || colour_rgba, "colour is NULL");
    Cmd c;
    memset(&c, 0, sizeof c);
    c.kind = 1;
    c.clear_buffers = buffers;
    if (colour_rgba) memcpy(c.clear_colour, colour_rgba, 16);
    c.clear_depth = depth;
    // a clear of both planes makes everything recorded before it dead
    if ((buffers & (OGRECU_CLEAR_COLOUR | OGRECU_CLEAR_DEPTH)) == (OGRECU_CLEAR_COLOUR | OGRECU_CLEAR_DEPTH)) rt->cmds.clear();
    rt->cmds.push_back(c);
    return OGRECU_OK;
}

int ogrecu_draw(ogrecu_rt* rt, const ogrecu_draw_desc* d) { return ogrecu_draw_with_state(rt, d, nullptr); }

static int record_draw(ogrecu_rt* rt, const ogrecu_draw_desc* d, const ogrecu_draw_state* st, const ogrecu_buf* const* refs, uint32_t num_refs);

int ogrecu_draw_with_state(ogrecu_rt* rt, const ogrecu_draw_desc* d, const ogrecu_draw_state* st) { return record_draw(rt, d, st, nullptr, 0); }

int ogrecu_draw_refs(ogrecu_rt* rt, const ogrecu_draw_desc* d, const ogrecu_draw_state* st, const ogrecu_buf* const* refs, uint32_t num_refs)
{
    OGRECU_REQUIRE(refs || num_refs == 0, "refs is NULL");
    return record_draw(rt, d, st, refs ? refs : (const ogrecu_buf* const*)&refs, num_refs);
}

// refs == NULL: the buffers the draw reads are looked up from its device pointers; otherwise the caller names them
static int record_draw(ogrecu_rt* rt, const ogrecu_draw_desc* d, const ogrecu_draw_state* st, const ogrecu_buf* const* refs, uint32_t num_refs)
{
    OGRECU_REQUIRE(rt && d, "NULL argument");
    if (st && st->enable) {
        OGRECU_REQUIRE((st->enable & ~0x1fu) == 0, "unknown OGRECU_EXT_* bit");
        OGRECU_REQUIRE(!(st->enable & OGRECU_EXT_DEPTH_FUNC) || st->depth_func <= OGRECU_CMPF_GREATER, "bad depth function");
        OGRECU_REQUIRE(!(st->enable & OGRECU_EXT_ALPHA_REJECT) || (st->alpha_func <= OGRECU_CMPF_GREATER && st->alpha_ref <= 255), "bad alpha rejection");
        OGRECU_REQUIRE(!(st->enable & OGRECU_EXT_CULL_MODE) || (st->cull_mode >= OGRECU_CULL_NONE && st->cull_mode <= OGRECU_CULL_ANTICLOCKWISE),
                       "bad culling mode");
        OGRECU_REQUIRE(!(st->enable & OGRECU_EXT_VERTEX_COLOUR) || (st->colours && ((st->colour_stride | (uintptr_t)st->colours) & 3) == 0),
                       "vertex colours: NULL or not 4-byte aligned");
    }
    if (d->prim_type != OGRECU_PRIM_TRIANGLE_LIST && d->prim_type != OGRECU_PRIM_TRIANGLE_STRIP) {
        // TinyRenderSystem::_render silently ignores other operation types (.cpp:354-357)
        return OGRECU_OK;
    }
    OGRECU_REQUIRE(d->positions, "VES_POSITION required");   // OgreAssert in the reference (.cpp:361)
    OGRECU_REQUIRE(d->index_type == OGRECU_INDEX_NONE || d->index_type == OGRECU_INDEX_U16 || d->index_type == OGRECU_INDEX_U32, "bad index type");
    OGRECU_REQUIRE(d->index_type == OGRECU_INDEX_NONE || d->indices, "indices are NULL");
    OGRECU_REQUIRE(((d->pos_stride | d->uv_stride | d->normal_stride | (uintptr_t)d->positions | (uintptr_t)d->uvs | (uintptr_t)d->normals) & 3) == 0,
                   "vertex streams must be 4-byte aligned");
    OGRECU_REQUIRE(!d->texture || d->texture->ctx == root_of(rt->ctx), "texture belongs to another context");
    OGRECU_REQUIRE(!d->instance_count || d->prim_type == OGRECU_PRIM_TRIANGLE_LIST, "instanced draws are triangle lists");
    OGRECU_REQUIRE((uint64_t)(d->draw_count / 3) * (d->instance_count ? d->instance_count : 1) < 0xffffffffull, "too many triangles in one draw");
    if (tri_count_of(*d) == 0) return OGRECU_OK;
    Cmd c;
    memset(&c, 0, sizeof c);
    c.kind = 0;
    c.draw = *d;
    if (st) c.state = *st;
    rt->cmds.push_back(c);
    // the buffers the draw will read when it is flushed: a host write to one of them (or a deferred OptimisedUtil job that
    // overwrites one) must run the recorded draws first (ogrecu_buf_map, ogrecu_software_vertex_skinning_host)
    ogrecu_ctx* root = root_of(rt->ctx);
    if (refs) {
        for (uint32_t k = 0; k < num_refs; ++k)
            if (refs[k]) { OGRECU_REQUIRE(refs[k]->ctx == root, "a referenced buffer belongs to another context"); const_cast<ogrecu_buf*>(refs[k])->draw_ref = root->draw_epoch; }
        return OGRECU_OK;
    }
    const void* ptrs[5] = {d->positions, d->uvs, d->normals, d->indices, (st && (st->enable & OGRECU_EXT_VERTEX_COLOUR)) ? st->colours : nullptr};
    for (const void* r : ptrs)
        if (ogrecu_buf* b = buf_of_device_ptr(root, r)) b->draw_ref = root->draw_epoch;
    return OGRECU_OK;
}

int ogrecu_flush(ogrecu_rt* rt)
{
    OGRECU_REQUIRE(rt, "rt is NULL");
    if (rt->cmds.empty()) return OGRECU_OK;
    return rt_flush_cmds(rt);
}

int ogrecu_draw_views(ogrecu_rt* const* rts, uint32_t n_views, const ogrecu_draw_desc* desc, const float* per_view_mvp,
                      const float* per_view_mvit, const float* per_view_light_dir)
{
    OGRECU_REQUIRE(rts && desc && per_view_mvp, "NULL argument");
    if (n_views == 0) return OGRECU_OK;
    OGRECU_REQUIRE(rts[0], "NULL target");
    ogrecu_ctx* root = root_of(rts[0]->ctx);
    for (uint32_t v = 0; v < n_views; ++v) {
        OGRECU_REQUIRE(rts[v] && root_of(rts[v]->ctx) == root, "targets must share one context (or one device group)");
        OGRECU_REQUIRE(!rts[v]->split, "batched views go to whole targets: place them on the group's devices instead of splitting them");
        OGRECU_REQUIRE(rts[v]->w == rts[0]->w && rts[v]->h == rts[0]->h, "targets must have identical size");
        OGRECU_REQUIRE(rts[v]->strip_y0 == rts[0]->strip_y0 && rts[v]->strip_y1 == rts[0]->strip_y1 && rts[v]->band_period == rts[0]->band_period &&
                       rts[v]->band_phase == rts[0]->band_phase && rts[v]->band_tiles == rts[0]->band_tiles, "targets must share the strip");
    }
    OGRECU_CUDA(cudaSetDevice(root->device));
    OGRECU_CHECK(flush_skin_batch(root));      // queued OptimisedUtil jobs may produce what this draw reads
    // One pipeline pass per device over the views whose targets live there (a single context: one pass over all of them).  The
    // geometry is replicated, so the passes share nothing; both host phases go device by device and the passes overlap.
    std::vector<ogrecu_ctx*> owners;
    for (uint32_t v = 0; v < n_views; ++v)
        if (std::find(owners.begin(), owners.end(), rts[v]->ctx) == owners.end()) owners.push_back(rts[v]->ctx);
    const size_t G = owners.size();
    // the peers start after everything the root's stream holds so far — ordered once per device, not once per target
    for (ogrecu_ctx* dev : owners) if (dev->root) OGRECU_CHECK(stream_wait(dev, root));
    // earlier clears / draws on each target come first.  The usual case — every target holds exactly one pending clear, all with the
    // same values — runs as one launch per device and 64 targets instead of two launches per target
    {
        auto only_clear = [](const ogrecu_rt* t) { return t->cmds.size() == 1 && t->cmds[0].kind == 1; };
        auto same_clear = [](const Cmd& a, const Cmd& b) {
            return a.clear_buffers == b.clear_buffers && memcmp(a.clear_colour, b.clear_colour, 16) == 0 && memcmp(&a.clear_depth, &b.clear_depth, 4) == 0;
        };
        bool batch = n_views >= 2;
        for (uint32_t v = 0; v < n_views && batch; ++v) batch = only_clear(rts[v]) && same_clear(rts[v]->cmds[0], rts[0]->cmds[0]);
        if (batch) {
            const Cmd c = rts[0]->cmds[0];
            uint32_t q[3];
            for (int k = 0; k < 3; ++k) {   // Bitwise::floatToFixed(v, 8), OgreBitwise.h:237-242 (zero colour: the same bytes as Image::setTo's shortcut)
                const float f = c.clear_colour[k];
                q[k] = ((f <= 0.0f) ? 0u : (f >= 1.0f) ? 255u : (uint32_t)(f * 256.0f)) & 0xffu;
            }
            const size_t npx = (size_t)rts[0]->w * rts[0]->h;
            const size_t chunks = ((c.clear_buffers & OGRECU_CLEAR_COLOUR) ? (npx * 3 + 15) / 16 : 0) + ((c.clear_buffers & OGRECU_CLEAR_DEPTH) ? (npx + 3) / 4 : 0);
            for (ogrecu_ctx* dev : owners) {
                OGRECU_CUDA(cudaSetDevice(dev->device));
                ClearPlanes T;
                uint32_t n = 0;
                auto launch = [&]() -> int {
                    if (n && chunks) {
                        clear_targets_kernel<<<dim3(div_up(chunks, 256), n), 256, 0, dev->stream>>>(T, npx, q[0], q[1], q[2], c.clear_depth, c.clear_buffers);
                        OGRECU_CUDA(cudaGetLastError());
                        count_launch(dev);
                    }
                    n = 0;
                    return OGRECU_OK;
                };
                for (uint32_t v = 0; v < n_views; ++v) {
                    if (rts[v]->ctx != dev) continue;
                    T.colour[n] = rts[v]->colour; T.depth[n] = rts[v]->depth;
                    rts[v]->cmds.clear();
                    if (++n == CLEAR_BATCH) OGRECU_CHECK(launch());
                }
                OGRECU_CHECK(launch());
            }
            OGRECU_CUDA(cudaSetDevice(root->device));
        } else {
            for (uint32_t v = 0; v < n_views; ++v)
                if (!rts[v]->cmds.empty()) OGRECU_CHECK(rt_flush_cmds(rts[v], false));
        }
    }
    if (desc->prim_type != OGRECU_PRIM_TRIANGLE_LIST && desc->prim_type != OGRECU_PRIM_TRIANGLE_STRIP) {
        for (ogrecu_ctx* dev : owners) if (dev->root) OGRECU_CHECK(stream_wait(root, dev));
        return OGRECU_OK;
    }
    OGRECU_REQUIRE(desc->positions, "VES_POSITION required");
    std::vector<std::vector<ogrecu_rt*>> targets(G);
    std::vector<std::vector<Item>> items(G);
    std::vector<PeerSeg> tr(G);
    std::vector<SegRun> runs(G);
    int rc = group_parallel(root, G, [&](size_t g) -> int {
        ogrecu_ctx* dev = owners[g];
        const ogrecu_draw_desc* d = desc;
        if (dev->root) {
            std::vector<Item> one(1, Item{desc, 0, nullptr, nullptr, nullptr, nullptr});
            OGRECU_CHECK(translate_segment(dev, one, tr[g]));
            d = &tr[g].descs[0];
        }
        for (uint32_t v = 0; v < n_views; ++v) {
            if (rts[v]->ctx != dev) continue;
            items[g].push_back({d, (uint32_t)targets[g].size(), per_view_mvp + 16 * v, per_view_mvit ? per_view_mvit + 16 * v : nullptr,
                                per_view_light_dir ? per_view_light_dir + 3 * v : nullptr, nullptr});
            targets[g].push_back(rts[v]);
        }
        return segment_begin(dev, targets[g].data(), (uint32_t)targets[g].size(), items[g], runs[g]);
    });
    {
        const int rc_begin = rc;
        const int rc2 = group_parallel(root, G, [&](size_t g) -> int {
            if (rc_begin == OGRECU_OK) return segment_finish(runs[g]);
            cudaSetDevice(owners[g]->device); cudaStreamSynchronize(owners[g]->stream);
            return OGRECU_OK;
        });
        if (rc == OGRECU_OK) rc = rc2;
    }
    for (size_t g = 0; g < G; ++g)
        if (owners[g]->root) { const int r = stream_wait(root, owners[g]); if (rc == OGRECU_OK) rc = r; }
    cudaSetDevice(root->device);
    return rc;
}

int ogrecu_rt_set_strip(ogrecu_rt* rt, uint32_t y0, uint32_t y1)
{
    OGRECU_REQUIRE(rt, "rt is NULL");
    OGRECU_REQUIRE(y0 % TILE == 0 && (y1 % TILE == 0 || y1 >= rt->h) && y0 <= y1, "strip rows must be multiples of the tile height");
    OGRECU_REQUIRE(!rt->split, "a split target's rows are assigned by the device group");
    OGRECU_CHECK(ogrecu_flush(rt));
    rt->strip_y0 = std::min(y0, rt->h);
    rt->strip_y1 = std::min(y1, rt->h);
    rt->band_tiles = 1; rt->band_phase = 0; rt->band_period = 1;
    return OGRECU_OK;
}

int ogrecu_composite_depth_dev(ogrecu_ctx* ctx, uint8_t* dst_rgb, float* dst_depth, const uint8_t* src_rgb, const float* src_depth,
                               size_t num_pixels, uint32_t num_sources, size_t src_stride_pixels, const float clear_rgba[4], float clear_depth)
{
    OGRECU_REQUIRE(ctx && clear_rgba, "NULL argument");
    if (!num_pixels || !num_sources) return OGRECU_OK;
    OGRECU_REQUIRE(dst_rgb && dst_depth && src_rgb && src_depth, "NULL plane");
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    uint32_t q = 0;
    for (int k = 0; k < 3; ++k) {                                   // the bytes clearFrameBuffer writes (Bitwise::floatToFixed(v, 8))
        const float v = clear_rgba[k];
        const uint32_t b = (v <= 0.0f) ? 0u : (v >= 1.0f) ? 255u : ((uint32_t)(v * 256.0f) & 0xffu);
        q |= b << (8 * k);
    }
    composite_depth_kernel<<<div_up(num_pixels, 256), 256, 0, ctx->stream>>>(dst_rgb, dst_depth, src_rgb, src_depth, num_pixels, num_sources,
                                                                             src_stride_pixels, q, clear_depth);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

int ogrecu_rt_set_bands(ogrecu_rt* rt, uint32_t band_rows, uint32_t phase, uint32_t period)
{
    OGRECU_REQUIRE(rt, "rt is NULL");
    OGRECU_REQUIRE(band_rows >= TILE && band_rows % TILE == 0, "band height must be a positive multiple of the tile height");
    OGRECU_REQUIRE(period >= 1 && phase < period, "phase must be below period");
    OGRECU_REQUIRE(!rt->split, "a split target's rows are assigned by the device group");
    OGRECU_CHECK(ogrecu_flush(rt));
    rt->strip_y0 = 0; rt->strip_y1 = rt->h;
    rt->band_tiles = band_rows / TILE; rt->band_phase = phase; rt->band_period = period;
    return OGRECU_OK;
}

static int read_plane(ogrecu_rt* rt, const uint8_t* plane, size_t px_bytes, uint32_t l, uint32_t t, uint32_t r, uint32_t b,
                      uint8_t* dst, size_t pitch)
{
    // validation mirrors TinyWindow::copyContentsToMemory (OgreTinyWindow.cpp:73-77)
    OGRECU_REQUIRE(r <= rt->w && b <= rt->h && l <= r && t <= b, "Invalid box");
    ogrecu_ctx* ctx = rt->ctx;
    OGRECU_CHECK(ogrecu_flush(rt));
    DeviceScope on_device(rt->ctx->device);
    const size_t row = (size_t)(r - l) * px_bytes, rows = b - t;
    if (!row || !rows) return OGRECU_OK;
    OGRECU_REQUIRE(dst && pitch >= row, "bad destination");
    // a destination in pinned host memory (ogrecu_host_alloc, a buffer shadow, cudaHostRegister'ed application
    // memory) takes the DMA directly; anything else is staged through the context's pinned block
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, dst) == cudaSuccess && pa.type == cudaMemoryTypeHost) {
        OGRECU_CUDA(cudaMemcpy2DAsync(dst, pitch, plane + ((size_t)t * rt->w + l) * px_bytes, (size_t)rt->w * px_bytes, row, rows,
                                      cudaMemcpyDeviceToHost, ctx->stream));
        OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
        return OGRECU_OK;
    }
    (void)cudaGetLastError();
    OGRECU_CHECK(pinned_reserve(ctx, row * rows));
    OGRECU_CUDA(cudaMemcpy2DAsync(ctx->pinned, row, plane + ((size_t)t * rt->w + l) * px_bytes, (size_t)rt->w * px_bytes, row, rows,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    if (pitch == row) memcpy(dst, ctx->pinned, row * rows);
    else for (size_t y = 0; y < rows; ++y) memcpy(dst + y * pitch, (uint8_t*)ctx->pinned + y * row, row);
    return OGRECU_OK;
}

int ogrecu_rt_read_colour(ogrecu_rt* rt, uint32_t l, uint32_t t, uint32_t r, uint32_t b, uint8_t* dst, size_t pitch)
{
    OGRECU_REQUIRE(rt, "rt is NULL");
    return read_plane(rt, rt->colour, 3, l, t, r, b, dst, pitch);
}

// rows of `row` bytes, tightly packed in ctx->stage_out, -> host rows `pitch` bytes apart; waits for the copy.  A destination
// in pinned host memory takes the DMA directly, anything else is staged through the context's pinned block.
static int stage_out_to_host(ogrecu_ctx* ctx, void* dst, size_t pitch, size_t row, size_t rows)
{
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, dst) == cudaSuccess && pa.type == cudaMemoryTypeHost) {
        OGRECU_CUDA(cudaMemcpy2DAsync(dst, pitch, ctx->stage_out.ptr, row, row, rows, cudaMemcpyDeviceToHost, ctx->stream));
        OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
        return OGRECU_OK;
    }
    (void)cudaGetLastError();
    OGRECU_CHECK(pinned_reserve(ctx, row * rows));
    OGRECU_CUDA(cudaMemcpyAsync(ctx->pinned, ctx->stage_out.ptr, row * rows, cudaMemcpyDeviceToHost, ctx->stream));
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    if (pitch == row) memcpy(dst, ctx->pinned, row * rows);
    else for (size_t y = 0; y < rows; ++y) memcpy((uint8_t*)dst + y * pitch, (uint8_t*)ctx->pinned + y * row, row);
    return OGRECU_OK;
}

// RGB8 sub-box -> one packed 32-bit word per pixel: r << r_shift | g << g_shift | b << b_shift | fill.  What
// PixelUtil::bulkPixelConversion does on the CPU in TinyWindow::copyContentsToMemory (OgreTinyWindow.cpp:79) for the
// 32-bit 8-bits-per-channel destination formats (OgrePixelConversions.h:333-351 and the generic unpack/pack path, which
// is the identity on 8-bit channels), done before the copy so that the host receives the final bytes.
__global__ void pack32_kernel(const uint8_t* __restrict__ rgb, uint32_t plane_w, uint32_t l, uint32_t t, uint32_t bw, uint32_t bh,
                              uint32_t* __restrict__ out, uint32_t rs, uint32_t gs, uint32_t bs, uint32_t fill)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= bw || y >= bh) return;
    const uint8_t* p = rgb + ((size_t)(t + y) * plane_w + l + x) * 3;
    out[(size_t)y * bw + x] = ((uint32_t)p[0] << rs) | ((uint32_t)p[1] << gs) | ((uint32_t)p[2] << bs) | fill;
}

int ogrecu_rt_read_colour_packed32(ogrecu_rt* rt, uint32_t l, uint32_t t, uint32_t r, uint32_t b, void* dst, size_t pitch, uint32_t r_shift,
                                   uint32_t g_shift, uint32_t b_shift, uint32_t fill)
{
    OGRECU_REQUIRE(rt, "rt is NULL");
    OGRECU_REQUIRE(r <= rt->w && b <= rt->h && l <= r && t <= b, "Invalid box");
    OGRECU_REQUIRE(r_shift <= 24 && g_shift <= 24 && b_shift <= 24, "channel shifts are 0..24");
    ogrecu_ctx* ctx = rt->ctx;
    OGRECU_CHECK(ogrecu_flush(rt));
    DeviceScope on_device(rt->ctx->device);
    const uint32_t bw = r - l, bh = b - t;
    if (!bw || !bh) return OGRECU_OK;
    const size_t row = (size_t)bw * 4;
    OGRECU_REQUIRE(dst && pitch >= row, "bad destination");
    OGRECU_CHECK(scratch_reserve(ctx, ctx->stage_out, row * bh));
    pack32_kernel<<<dim3(div_up(bw, 256), bh), 256, 0, ctx->stream>>>(rt->colour, rt->w, l, t, bw, bh, (uint32_t*)ctx->stage_out.ptr, r_shift, g_shift,
                                                                       b_shift, fill);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return stage_out_to_host(ctx, dst, pitch, row, bh);
}

// Image::scale(FILTER_BILINEAR) of an RGB8 sub-box = LinearResampler_Byte<3>::scale (OgreMain/src/OgreImage.cpp:389-458,
// OgreMain/src/OgreImageResampler.h:283-339), one thread per destination pixel: the reference walks the source position
// incrementally in 16.48 fixed point (sx_48 += stepx), which is x * stepx modulo 2^64 — evaluated directly here.  The blend is
// the reference's unsigned 8.24 arithmetic, so the bytes are identical.  out_px = 3: RGB8 rows; 4: the scaled pixel packed as
// pack32_kernel does (the temp-buffer + bulkPixelConversion tail of Image::scale when the formats differ, OgreImage.cpp:453-457).
__device__ __forceinline__ void scale_axis(uint64_t pos_48, uint32_t extent, uint32_t& s1, uint32_t& s2, uint32_t& frac)
{
    uint32_t temp = (uint32_t)(pos_48 >> 36);
    temp = (temp > 0x800u) ? temp - 0x800u : 0u;
    frac = temp & 0xFFFu;
    s1 = temp >> 12;
    s2 = min(s1 + 1u, extent - 1u);
}

__global__ void __launch_bounds__(256) scale_rgb8_kernel(const uint8_t* __restrict__ rgb, uint32_t plane_w, uint32_t l, uint32_t t, uint32_t sw,
                                                         uint32_t sh, uint8_t* __restrict__ out, uint32_t dw, uint32_t dh, uint64_t stepx,
                                                         uint64_t stepy, uint32_t out_px, uint32_t rs, uint32_t gs, uint32_t bs, uint32_t fill)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    uint32_t sx1, sx2, sxf, sy1, sy2, syf;
    scale_axis((stepx >> 1) - 1 + (uint64_t)x * stepx, sw, sx1, sx2, sxf);
    scale_axis((stepy >> 1) - 1 + (uint64_t)y * stepy, sh, sy1, sy2, syf);
    const uint8_t* r1 = rgb + ((size_t)(t + sy1) * plane_w + l) * 3;
    const uint8_t* r2 = rgb + ((size_t)(t + sy2) * plane_w + l) * 3;
    const uint32_t sxfsyf = sxf * syf;
    const uint32_t w11 = 0x1000000u - (sxf << 12) - (syf << 12) + sxfsyf, w21 = (sxf << 12) - sxfsyf, w12 = (syf << 12) - sxfsyf;
    uint32_t c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t accum = r1[sx1 * 3 + k] * w11 + r1[sx2 * 3 + k] * w21 + r2[sx1 * 3 + k] * w12 + r2[sx2 * 3 + k] * sxfsyf;
        c[k] = (accum + 0x800000u) >> 24;
    }
    if (out_px == 4) reinterpret_cast<uint32_t*>(out)[(size_t)y * dw + x] = (c[0] << rs) | (c[1] << gs) | (c[2] << bs) | fill;
    else {
        uint8_t* o = out + ((size_t)y * dw + x) * 3;
        o[0] = (uint8_t)c[0]; o[1] = (uint8_t)c[1]; o[2] = (uint8_t)c[2];
    }
}

int ogrecu_rt_read_colour_scaled(ogrecu_rt* rt, uint32_t l, uint32_t t, uint32_t r, uint32_t b, uint32_t dst_w, uint32_t dst_h, void* dst,
                                 size_t pitch, uint32_t dst_px_bytes, uint32_t r_shift, uint32_t g_shift, uint32_t b_shift, uint32_t fill)
{
    OGRECU_REQUIRE(rt, "rt is NULL");
    OGRECU_REQUIRE(r <= rt->w && b <= rt->h && l <= r && t <= b, "Invalid box");
    OGRECU_REQUIRE(dst_px_bytes == 3 || dst_px_bytes == 4, "destination pixels are 3 (RGB8) or 4 (packed) bytes");
    OGRECU_REQUIRE(dst_px_bytes == 3 || (r_shift <= 24 && g_shift <= 24 && b_shift <= 24), "channel shifts are 0..24");
    ogrecu_ctx* ctx = rt->ctx;
    OGRECU_CHECK(ogrecu_flush(rt));
    DeviceScope on_device(rt->ctx->device);
    const uint32_t sw = r - l, sh = b - t;
    if (!dst_w || !dst_h) return OGRECU_OK;
    OGRECU_REQUIRE(sw && sh, "empty source box");
    OGRECU_REQUIRE(sw < 65536u && sh < 65536u && dst_h < 65536u, "boxes are limited to 65535 pixels per side");
    const size_t row = (size_t)dst_w * dst_px_bytes;
    OGRECU_REQUIRE(dst && pitch >= row, "bad destination");
    OGRECU_CHECK(scratch_reserve(ctx, ctx->stage_out, row * dst_h));
    const uint64_t stepx = ((uint64_t)sw << 48) / dst_w, stepy = ((uint64_t)sh << 48) / dst_h;      // OgreImageResampler.h:299-300
    scale_rgb8_kernel<<<dim3(div_up(dst_w, 256), dst_h), 256, 0, ctx->stream>>>(rt->colour, rt->w, l, t, sw, sh, (uint8_t*)ctx->stage_out.ptr, dst_w,
                                                                                 dst_h, stepx, stepy, dst_px_bytes, r_shift, g_shift, b_shift, fill);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return stage_out_to_host(ctx, dst, pitch, row, dst_h);
}

int ogrecu_rt_read_depth(ogrecu_rt* rt, uint32_t l, uint32_t t, uint32_t r, uint32_t b, float* dst, size_t pitch)
{
    OGRECU_REQUIRE(rt, "rt is NULL");
    return read_plane(rt, (const uint8_t*)rt->depth, 4, l, t, r, b, (uint8_t*)dst, pitch);
}

static int write_plane(ogrecu_rt* rt, uint8_t* plane, size_t px_bytes, const uint8_t* src, size_t pitch)
{
    ogrecu_ctx* ctx = rt->ctx;
    OGRECU_CHECK(ogrecu_flush(rt));
    DeviceScope on_device(rt->ctx->device);
    const size_t row = (size_t)rt->w * px_bytes;
    OGRECU_REQUIRE(src && pitch >= row, "bad source");
    OGRECU_CHECK(pinned_reserve(ctx, row * rt->h));
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t y = 0; y < rt->h; ++y) memcpy((uint8_t*)ctx->pinned + y * row, src + y * pitch, row);
    OGRECU_CUDA(cudaMemcpyAsync(plane, ctx->pinned, row * rt->h, cudaMemcpyHostToDevice, ctx->stream));
    return OGRECU_OK;
}

int ogrecu_rt_write_colour(ogrecu_rt* rt, const uint8_t* src, size_t pitch)
{
    OGRECU_REQUIRE(rt, "rt is NULL");
    return write_plane(rt, rt->colour, 3, src, pitch);
}

int ogrecu_rt_write_depth(ogrecu_rt* rt, const float* src, size_t pitch)
{
    OGRECU_REQUIRE(rt, "rt is NULL");
    return write_plane(rt, (uint8_t*)rt->depth, 4, (const uint8_t*)src, pitch);
}

int ogrecu_rt_colour_device_ptr(ogrecu_rt* rt, void** p)
{
    OGRECU_REQUIRE(rt && p, "NULL argument");
    *p = rt->colour;
    return OGRECU_OK;
}

int ogrecu_rt_depth_device_ptr(ogrecu_rt* rt, void** p)
{
    OGRECU_REQUIRE(rt && p, "NULL argument");
    *p = rt->depth;
    return OGRECU_OK;
}

int ogrecu_rt_last_stats(ogrecu_rt* rt, ogrecu_raster_stats* out)
{
    OGRECU_REQUIRE(rt && out, "NULL argument");
    if (rt->stats_pending) {
        auto sm = scratch_map().find(rt->ctx);
        if (sm != scratch_map().end() && sm->second.counters.ptr) {
            uint32_t c[2] = {0, 0};
            DeviceScope on_device(rt->ctx->device);
            OGRECU_CUDA(cudaMemcpyAsync(c, (uint32_t*)sm->second.counters.ptr + 2 * rt->stats_slot, sizeof c, cudaMemcpyDeviceToHost, rt->ctx->stream));
            OGRECU_CUDA(cudaStreamSynchronize(rt->ctx->stream));
            rt->stats.tiles_visibility = c[0];
            rt->stats.tiles_ordered = c[1];
            rt->stats.tiles_active = (uint64_t)c[0] + c[1];
        }
        rt->stats_pending = false;
    }
    *out = rt->stats;
    return OGRECU_OK;
}

}  // extern "C"
