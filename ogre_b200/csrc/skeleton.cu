// skeleton.cu — bone-palette production for many instances of one skeleton, on the device.
//
// Reference work replaced (per entity, per frame, on the CPU render thread):
//   Entity::cacheBoneMatrices                     OgreMain/src/OgreEntity.cpp:1307-1330
//     SkeletonInstance::setAnimationState         OgreMain/src/OgreSkeleton.cpp:162-217
//       Animation::apply / _getTimeIndex          OgreMain/src/OgreAnimation.cpp:326-342,594-618
//       NodeAnimationTrack::applyToNode           OgreMain/src/OgreAnimationTrack.cpp:57-126,373-512
//     Skeleton::_getBoneMatrices                  OgreMain/src/OgreSkeleton.cpp:402-424
//       Node::updateFromParentImpl                OgreMain/src/OgreNode.cpp:202-256
//       Bone::_getOffsetTransform                 OgreMain/src/OgreBone.cpp:88-105
//       TransformBaseReal::makeTransform          OgreMain/src/OgreMatrix4.cpp:205-222
//
// Why on the device: at 10 000 instances the palettes are 44 MB per frame; uploading them costs more than skinning
// the 55 M vertices.  The animation STATES (index, time, weight) are 12 bytes per slot.
//
// Layout: the key frames of every animation live in HBM once (shared by all instances, L2-resident: Sinbad's 13
// animations are 440 KB).  One thread per (instance, bone):
//   1. local transform = binding pose, then each enabled state's interpolated key frame applied in slot order;
//   2. derived transforms level by level through shared memory (parents are one level up, __syncthreads between);
//      threads take the bones in level order, so a warp runs the derivation step for one or two levels only;
//   3. offset transform -> Affine3, written as four 16-byte stores (consecutive bones are consecutive in memory).
// Every float operation is written in the reference's association; the library is compiled with -fmad=false, so
// each operator below is one IEEE binary32 rounding on the device exactly as it is on the host (x86-64, no FMA).
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

using namespace ogrecu;

namespace ogrecu {

struct Quat { float w, x, y, z; };
struct Vec3 { float x, y, z; };
struct Xform { Vec3 p; Quat q; Vec3 s; };

#define HD __host__ __device__ __forceinline__

HD Quat q_mul(const Quat& a, const Quat& b)
{
    return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}

HD Quat q_unit(const Quat& q)       // Quaternion::normalise: (1.0f / len) * q
{
    const float len = sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    const float s = 1.0f / len;
    return Quat{s * q.w, s * q.x, s * q.y, s * q.z};
}

HD Quat q_nlerp(float t, const Quat& p, const Quat& q, bool shortest)
{
    const float c = p.w * q.w + p.x * q.x + p.y * q.y + p.z * q.z;
    Quat r;
    if (c < 0.0f && shortest) r = Quat{p.w + t * ((-q.w) - p.w), p.x + t * ((-q.x) - p.x), p.y + t * ((-q.y) - p.y), p.z + t * ((-q.z) - p.z)};
    else r = Quat{p.w + t * (q.w - p.w), p.x + t * (q.x - p.x), p.y + t * (q.y - p.y), p.z + t * (q.z - p.z)};
    return q_unit(r);
}

// Quaternion::Slerp, OgreQuaternion.cpp:373-413 (msEpsilon = 1e-03f).  sqrtf is correctly rounded on both sides; atan2f and sinf
// are libm's on the host and CUDA's (<= 2 ulp) on the device: RIM_SPHERICAL / IM_SPLINE rotations carry a TOLERANCE, not the
// bit-identity of the linear modes (tests/test_skeleton_gpu.py states it).
HD Quat q_slerp(float t, const Quat& p, const Quat& q, bool shortest)
{
    float c = p.w * q.w + p.x * q.x + p.y * q.y + p.z * q.z;
    Quat r = q;
    if (c < 0.0f && shortest) { c = -c; r = Quat{-q.w, -q.x, -q.y, -q.z}; }
    if (fabsf(c) < 1 - 1e-03f) {
        const float sn = sqrtf(1 - c * c);
        const float angle = atan2f(sn, c);
        const float inv = 1.0f / sn;
        const float c0 = sinf((1.0f - t) * angle) * inv;
        const float c1 = sinf(t * angle) * inv;
        return Quat{c0 * p.w + c1 * r.w, c0 * p.x + c1 * r.x, c0 * p.y + c1 * r.y, c0 * p.z + c1 * r.z};
    }
    const float u = 1.0f - t;
    return q_unit(Quat{u * p.w + t * r.w, u * p.x + t * r.x, u * p.y + t * r.y, u * p.z + t * r.z});
}

// Quaternion::Squad, OgreQuaternion.cpp:449-457
HD Quat q_squad(float t, const Quat& p, const Quat& a, const Quat& b, const Quat& q, bool shortest)
{
    const float st = 2.0f * t * (1.0f - t);
    return q_slerp(st, q_slerp(t, p, q, shortest), q_slerp(t, a, b, false), false);
}

// SimpleSpline::interpolate(fromIndex, t), OgreSimpleSpline.cpp:87-154: powers * hCoeffs * pt with Vector4 * Matrix4
// (OgreMatrix4.h:588-596), every product and sum of the reference kept, zeros included — pure arithmetic, bit-exact
HD Vec3 hermite(float t, const Vec3& p1, const Vec3& p2, const Vec3& a, const Vec3& b)
{
    const float t2 = t * t, t3 = t2 * t;
    const float v0 = t3 * 2.0f + t2 * -3.0f + t * 0.0f + 1.0f * 1.0f;
    const float v1 = t3 * -2.0f + t2 * 3.0f + t * 0.0f + 1.0f * 0.0f;
    const float v2 = t3 * 1.0f + t2 * -2.0f + t * 1.0f + 1.0f * 0.0f;
    const float v3 = t3 * 1.0f + t2 * -1.0f + t * 0.0f + 1.0f * 0.0f;
    return Vec3{v0 * p1.x + v1 * p2.x + v2 * a.x + v3 * b.x, v0 * p1.y + v1 * p2.y + v2 * a.y + v3 * b.y, v0 * p1.z + v1 * p2.z + v2 * a.z + v3 * b.z};
}

HD Vec3 q_rotate(const Quat& q, const Vec3& v)      // Quaternion * Vector3 ("nVidia SDK implementation")
{
    Vec3 uv{q.y * v.z - q.z * v.y, q.z * v.x - q.x * v.z, q.x * v.y - q.y * v.x};
    Vec3 uuv{q.y * uv.z - q.z * uv.y, q.z * uv.x - q.x * uv.z, q.x * uv.y - q.y * uv.x};
    const float tw = 2.0f * q.w;
    uv.x *= tw; uv.y *= tw; uv.z *= tw;
    uuv.x *= 2.0f; uuv.y *= 2.0f; uuv.z *= 2.0f;
    return Vec3{(v.x + uv.x) + uuv.x, (v.y + uv.y) + uuv.y, (v.z + uv.z) + uuv.z};
}

// Node::updateFromParentImpl with a parent
HD Xform derive(const Xform& P, const Xform& L, bool inherit_orientation, bool inherit_scale)
{
    Xform d;
    d.q = inherit_orientation ? q_mul(P.q, L.q) : L.q;
    d.s = inherit_scale ? Vec3{P.s.x * L.s.x, P.s.y * L.s.y, P.s.z * L.s.z} : L.s;
    const Vec3 r = q_rotate(P.q, Vec3{P.s.x * L.p.x, P.s.y * L.p.y, P.s.z * L.p.z});
    d.p = Vec3{r.x + P.p.x, r.y + P.p.y, r.z + P.p.z};
    return d;
}

HD Xform load_xform(const float* f) { return Xform{Vec3{f[0], f[1], f[2]}, Quat{f[3], f[4], f[5], f[6]}, Vec3{f[7], f[8], f[9]}}; }

// per-animation device record
struct AnimDev {
    float length;
    uint32_t num_times;             // entries of the global key-time list (Animation::buildKeyFrameTimeList)
    const float* times;             // [num_times] sorted, unique
    const int32_t* bone_track;      // [num_bones] track of the bone, -1 = none
    const uint32_t* key_offset;     // [num_tracks + 1]
    const uint8_t* shortest;        // [num_tracks]
    const uint16_t* index_map;      // [num_tracks][num_times] AnimationTrack::_buildKeyFrameIndexMap
    const float4* keys;             // [num_keys][3]: (time, tx, ty, tz) (rw, rx, ry, rz) (sx, sy, sz, -)
    const uint4* bone_rec;          // [num_bones] (first key, key count, track, shortest-path flag) of the bone's track; count 0 = none
    uint32_t mode;                  // bit 0: IM_SPLINE, bit 1: RIM_SPHERICAL (OgreAnimation.h:70-90)
    const float4* tangents;         // IM_SPLINE: [num_keys][3] spline tangents of (translate, -) (rotation wxyz) (scale, -), made on the host
};

struct SkelDev {
    uint32_t num_bones, blend_mode, num_levels, num_anims;
    const int32_t* parent;          // [num_bones]
    const uint8_t* level;           // [num_bones] depth below the root
    const uint16_t* order;          // [num_bones] bone handled by thread x: bones sorted by level, so that a warp holds one or
                                    // two adjacent levels and runs the derivation step once or twice, not once per level
    const uint8_t* inherit;         // [num_bones][2]
    const float* bind;              // [num_bones][10]
    const float* bind_inverse;      // [num_bones][10]
    const AnimDev* anims;           // [num_anims]
    const int16_t* ppos;            // [num_bones] position (in `order`) of the parent of the bone at position p, -1 = root
    const uint16_t* level_start;    // [num_levels + 1] first position of each tree level in `order`
};

// blockDim = (bones rounded up to a warp, instances per CTA)
__global__ void __launch_bounds__(512) skel_palette_kernel(const SkelDev S, uint32_t num_instances, uint32_t num_states,
                                                           const int32_t* __restrict__ anim, const float* __restrict__ time,
                                                           const float* __restrict__ weight, float* __restrict__ out,
                                                           size_t stride_floats)
{
    extern __shared__ float s_derived[];                      // [instances per CTA][num_bones][10]
    const uint32_t inst = blockIdx.x * blockDim.y + threadIdx.y;
    const bool live = threadIdx.x < S.num_bones && inst < num_instances;
    const uint32_t b = live ? S.order[threadIdx.x] : 0u;
    float* mine = s_derived + ((size_t)threadIdx.y * S.num_bones + (live ? b : 0)) * 10;

    Xform L{};
    uint32_t lvl = 0xffffffffu;
    int32_t par = -1;
    if (live) {
        L = load_xform(S.bind + 10 * b);                      // Skeleton::reset
        lvl = S.level[b];
        par = S.parent[b];
        const int32_t* ia = anim + (size_t)inst * num_states;
        const float* it = time + (size_t)inst * num_states;
        const float* iw = weight + (size_t)inst * num_states;
        float factor = 1.0f;
        if (S.blend_mode == 0) {                              // ANIMBLEND_AVERAGE: rebalance when the weights sum above 1
            float total = 0.0f;
            for (uint32_t s = 0; s < num_states; ++s)
                if (ia[s] >= 0) total += iw[s];
            if (total > 1.0f) factor = 1.0f / total;
        }
        for (uint32_t s = 0; s < num_states; ++s) {
            const int32_t a = ia[s];
            if (a < 0 || (uint32_t)a >= S.num_anims) continue;
            const AnimDev A = S.anims[a];
            const float w = iw[s] * factor;
            const int32_t trk = A.bone_track[b];
            if (trk < 0 || !w) continue;
            const uint32_t k0 = A.key_offset[trk], nk = A.key_offset[trk + 1] - k0;
            if (nk == 0) continue;
            // Animation::_getTimeIndex: wrap, then lower_bound over [begin, end - 1)
            float tp = it[s];
            if (tp > A.length && A.length > 0.0f) tp = fmodf(tp, A.length);
            uint32_t lo = 0, hi = A.num_times - 1;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (A.times[mid] < tp) lo = mid + 1; else hi = mid;
            }
            // AnimationTrack::getKeyFramesAtTime
            uint32_t i = A.index_map[(size_t)trk * A.num_times + lo];
            const float4* K = A.keys + 3 * (size_t)k0;
            const uint32_t i2 = i;                            // k2: the key the index map names (k1 may step back from it)
            const float4 k2a = K[3 * i2];
            const float t2 = k2a.x;
            if (i != 0 && tp < t2) --i;
            const float4 k1a = K[3 * i], k1b = K[3 * i + 1], k1c = K[3 * i + 2];
            const float t1 = k1a.x;
            const float f = (t1 == t2) ? 0.0f : (tp - t1) / (t2 - t1);
            const bool shortest = A.shortest[trk] != 0;
            // NodeAnimationTrack::getInterpolatedKeyFrame (IM_LINEAR, RIM_LINEAR)
            Vec3 tr{k1a.y, k1a.z, k1a.w}, sc{k1c.x, k1c.y, k1c.z};
            Quat ro{k1b.x, k1b.y, k1b.z, k1b.w};
            if (f != 0.0f) {
                const float4 k2b = K[3 * i2 + 1], k2c = K[3 * i2 + 2];
                ro = q_nlerp(f, ro, Quat{k2b.x, k2b.y, k2b.z, k2b.w}, shortest);
                tr = Vec3{k1a.y + f * (k2a.y - k1a.y), k1a.z + f * (k2a.z - k1a.z), k1a.w + f * (k2a.w - k1a.w)};
                sc = Vec3{k1c.x + f * (k2c.x - k1c.x), k1c.y + f * (k2c.y - k1c.y), k1c.z + f * (k2c.z - k1c.z)};
            }
            // NodeAnimationTrack::applyToNode (scale argument 1.0: no linked skeleton)
            const float scl = 1.0f;
            L.p.x += tr.x * w * scl; L.p.y += tr.y * w * scl; L.p.z += tr.z * w * scl;
            const Quat rot = q_nlerp(w, Quat{1.0f, 0.0f, 0.0f, 0.0f}, ro, shortest);
            L.q = q_unit(q_mul(L.q, rot));                    // Node::rotate, TS_LOCAL, then normalise
            if ((sc.x != 1.0f || sc.y != 1.0f || sc.z != 1.0f) && w != 1.0f)
                sc = Vec3{1.0f + w * (sc.x - 1.0f), 1.0f + w * (sc.y - 1.0f), 1.0f + w * (sc.z - 1.0f)};
            L.s.x *= sc.x; L.s.y *= sc.y; L.s.z *= sc.z;
        }
    }

    // derived transforms, one tree level per round
    Xform D = L;
    for (uint32_t l = 0; l < S.num_levels; ++l) {
        if (live && lvl == l) {
            if (par >= 0) {
                const Xform P = load_xform(s_derived + ((size_t)threadIdx.y * S.num_bones + par) * 10);
                D = derive(P, L, S.inherit[2 * b] != 0, S.inherit[2 * b + 1] != 0);
            }
            mine[0] = D.p.x; mine[1] = D.p.y; mine[2] = D.p.z; mine[3] = D.q.w; mine[4] = D.q.x; mine[5] = D.q.y; mine[6] = D.q.z;
            mine[7] = D.s.x; mine[8] = D.s.y; mine[9] = D.s.z;
        }
        __syncthreads();
    }
    if (!live) return;

    // Bone::_getOffsetTransform + makeTransform (Quaternion::ToRotationMatrix, OgreQuaternion.cpp:84-108)
    const Xform I = load_xform(S.bind_inverse + 10 * b);
    const Vec3 ls{D.s.x * I.s.x, D.s.y * I.s.y, D.s.z * I.s.z};
    const Quat lr = q_mul(D.q, I.q);
    const Vec3 rp = q_rotate(lr, Vec3{ls.x * I.p.x, ls.y * I.p.y, ls.z * I.p.z});
    const Vec3 lt{D.p.x + rp.x, D.p.y + rp.y, D.p.z + rp.z};
    const float tx = lr.x + lr.x, ty = lr.y + lr.y, tz = lr.z + lr.z;
    const float twx = tx * lr.w, twy = ty * lr.w, twz = tz * lr.w;
    const float txx = tx * lr.x, txy = ty * lr.x, txz = tz * lr.x;
    const float tyy = ty * lr.y, tyz = tz * lr.y, tzz = tz * lr.z;
    float4* o = reinterpret_cast<float4*>(out + (size_t)inst * stride_floats + 16 * (size_t)b);
    o[0] = make_float4(ls.x * (1.0f - (tyy + tzz)), ls.y * (txy - twz), ls.z * (txz + twy), lt.x);
    o[1] = make_float4(ls.x * (txy + twz), ls.y * (1.0f - (txx + tzz)), ls.z * (tyz - twx), lt.y);
    o[2] = make_float4(ls.x * (txz - twy), ls.y * (tyz + twx), ls.z * (1.0f - (txx + tyy)), lt.z);
    o[3] = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
}


// ---- the kernel in use: 16 instances x 16 bone slots per CTA.  A half-warp holds 16 INSTANCES of one bone, so the track
// lookups are uniform per half-warp, the tree walk needs no per-lane level test, and one bone's transforms of the 16
// instances sit side by side in shared memory.  Bones are taken in tree-level order (`order`); slot s handles the
// positions s, s + 16, ...  (First version above — a lane per bone — ran every level's derivation step with mostly idle
// lanes: 55 us for 10 000 x 69; kept for A/B through "skel_mode".)
constexpr int SK_INST = 16;                 // instances per CTA
constexpr int SK_SLOTS = 16;                // bone slots per CTA (32: 52.9 us instead of 49.8)
constexpr int SK_XF = 10 * SK_INST;         // floats per bone in shared memory: [component][instance], the component index of
                                            // odd tree positions XOR 1 (sk_at) so that a warp's two half-warps — positions p and
                                            // p + 1 — hit different bank halves.  (A 16-float pad per bone did the same with 48.6 KB
                                            // per CTA for Sinbad's 69 bones: four CTAs per SM = 592 slots for the 625 CTAs of
                                            // 10 000 instances, i.e. a second wave of 33 CTAs doubled the 50 us launch.  44.4 KB
                                            // admit five: one wave.)
__device__ __forceinline__ uint32_t sk_at(uint32_t p, uint32_t c) { return p * SK_XF + ((c * SK_INST) ^ ((p & 1u) * SK_INST)); }

// getInterpolatedKeyFrame for IM_SPLINE (OgreAnimationTrack.cpp:438-450: the three splines' interpolate(fromIndex, t); a track's
// tangents are precomputed at ogrecu_skel_add_animation) and for IM_LINEAR with RIM_SPHERICAL (:417-432).  `key` = absolute index of
// the first key of the pair, whose successor is the second (f != 0 implies the step back of getKeyFramesAtTime).  Out of line:
// the linear path keeps its registers.
__device__ __noinline__ void interpolate_nonlinear(const AnimDev& A, uint32_t key, float f, bool shortest, Vec3& tr, Quat& ro, Vec3& sc)
{
    const float4* K = A.keys + 3 * (size_t)key;
    const float4 a0 = K[0], a1 = K[1], a2 = K[2], b0 = K[3], b1 = K[4], b2 = K[5];
    const Quat qa{a1.x, a1.y, a1.z, a1.w}, qb{b1.x, b1.y, b1.z, b1.w};
    if (A.mode & 1u) {
        // (interpolate() returns the second point itself at t == 1; f < 1 here: tp < t2)
        const float4* T = A.tangents + 3 * (size_t)key;
        const float4 ta0 = T[0], ta1 = T[1], ta2 = T[2], tb0 = T[3], tb1 = T[4], tb2 = T[5];
        tr = hermite(f, Vec3{a0.y, a0.z, a0.w}, Vec3{b0.y, b0.z, b0.w}, Vec3{ta0.x, ta0.y, ta0.z}, Vec3{tb0.x, tb0.y, tb0.z});
        sc = hermite(f, Vec3{a2.x, a2.y, a2.z}, Vec3{b2.x, b2.y, b2.z}, Vec3{ta2.x, ta2.y, ta2.z}, Vec3{tb2.x, tb2.y, tb2.z});
        ro = q_squad(f, qa, Quat{ta1.x, ta1.y, ta1.z, ta1.w}, Quat{tb1.x, tb1.y, tb1.z, tb1.w}, qb, shortest);
    } else {
        ro = q_slerp(f, qa, qb, shortest);
        tr = Vec3{a0.y + f * (b0.y - a0.y), a0.z + f * (b0.z - a0.z), a0.w + f * (b0.w - a0.w)};
        sc = Vec3{a2.x + f * (b2.x - a2.x), a2.y + f * (b2.y - a2.y), a2.z + f * (b2.z - a2.z)};
    }
}

__device__ __noinline__ Quat slerp_out_of_line(float t, const Quat& p, const Quat& q, bool shortest) { return q_slerp(t, p, q, shortest); }

struct StatePre { int32_t anim; float tp; uint32_t gi; float w; };

// NONLINEAR = some animation of the skeleton is IM_SPLINE / RIM_SPHERICAL; the all-linear instantiation has no trace of that path
// (with the out-of-line call in it the linear launch went from 64 to 81 us: stack frame, one CTA less in flight)
// EXTRA = the call carries AnimationState blend masks and / or manually controlled bones (ogrecu_palette_options)
struct PalExtra {
    const float* mask;              // [instance][state][bone] (instance stride mask_stride floats, 0 = shared), NULL = none
    size_t mask_stride;
    const int16_t* manual_slot;     // [bone] row of the instance's manual transforms, -1 = animated from the binding pose; NULL = none
    const float* manual;            // [instance][num_manual][10]
    uint32_t num_manual;
};
template <bool NONLINEAR, bool EXTRA>
__global__ void __launch_bounds__(SK_INST * SK_SLOTS, 5) skel_palette_kernel2(const SkelDev S, uint32_t num_instances, uint32_t num_states,
                                                                           const int32_t* __restrict__ anim, const float* __restrict__ time,
                                                                           const float* __restrict__ weight, float* __restrict__ out,
                                                                           size_t stride_floats, const PalExtra X)
{
    extern __shared__ float s_xf[];                                  // [num_bones][SK_XF]
    StatePre* s_st = reinterpret_cast<StatePre*>(s_xf + (size_t)S.num_bones * SK_XF);     // [SK_INST][num_states]
    const uint32_t ix = threadIdx.x, slot = threadIdx.y;
    const uint32_t inst = blockIdx.x * SK_INST + ix;
    const bool live = inst < num_instances;

    auto store_s = [&](uint32_t p, const Xform& X) {
        float* o = s_xf + ix;
        o[sk_at(p, 0)] = X.p.x; o[sk_at(p, 1)] = X.p.y; o[sk_at(p, 2)] = X.p.z; o[sk_at(p, 3)] = X.q.w; o[sk_at(p, 4)] = X.q.x;
        o[sk_at(p, 5)] = X.q.y; o[sk_at(p, 6)] = X.q.z; o[sk_at(p, 7)] = X.s.x; o[sk_at(p, 8)] = X.s.y; o[sk_at(p, 9)] = X.s.z;
    };
    auto load_s = [&](uint32_t p) {
        const float* o = s_xf + ix;
        return Xform{Vec3{o[sk_at(p, 0)], o[sk_at(p, 1)], o[sk_at(p, 2)]},
                     Quat{o[sk_at(p, 3)], o[sk_at(p, 4)], o[sk_at(p, 5)], o[sk_at(p, 6)]},
                     Vec3{o[sk_at(p, 7)], o[sk_at(p, 8)], o[sk_at(p, 9)]}};
    };

    // ---- phase 0: per (instance, state): wrapped time, global key index, effective weight — once, not once per bone
    for (uint32_t s = slot; s < num_states; s += SK_SLOTS) {
        StatePre P{-1, 0.f, 0u, 0.f};
        if (live) {
            const int32_t* ia = anim + (size_t)inst * num_states;
            const float* iw = weight + (size_t)inst * num_states;
            const int32_t a = ia[s];
            if (a >= 0 && (uint32_t)a < S.num_anims) {
                float factor = 1.0f;
                if (S.blend_mode == 0) {                              // ANIMBLEND_AVERAGE: rebalance when the weights sum above 1
                    float total = 0.0f;
                    for (uint32_t q = 0; q < num_states; ++q)
                        if (ia[q] >= 0) total += iw[q];
                    if (total > 1.0f) factor = 1.0f / total;
                }
                const AnimDev A = S.anims[a];
                float tp = time[(size_t)inst * num_states + s];
                if (tp > A.length && A.length > 0.0f) tp = fmodf(tp, A.length);
                uint32_t lo = 0, hi = A.num_times - 1;                // Animation::_getTimeIndex: lower_bound over [begin, end - 1)
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (A.times[mid] < tp) lo = mid + 1; else hi = mid;
                }
                P = StatePre{a, tp, lo, iw[s] * factor};
            }
        }
        s_st[ix * num_states + s] = P;
    }
    __syncthreads();

    // ---- phase 1: local transforms (binding pose, then every enabled state in slot order)
    for (uint32_t p = slot; p < S.num_bones; p += SK_SLOTS) {
        const uint32_t b = S.order[p];
        Xform L = load_xform(S.bind + 10 * b);
        if (EXTRA && live && X.manual_slot) {                  // Skeleton::reset leaves a manually controlled bone with what it holds
            const int ms = X.manual_slot[b];
            if (ms >= 0) L = load_xform(X.manual + ((size_t)inst * X.num_manual + (uint32_t)ms) * 10);
        }
        if (live) {
            for (uint32_t s = 0; s < num_states; ++s) {
                const StatePre P = s_st[ix * num_states + s];
                if (P.anim < 0) continue;
                const AnimDev A = S.anims[P.anim];
                float w = P.w;
                if (EXTRA && X.mask) w = X.mask[(size_t)inst * X.mask_stride + (size_t)s * S.num_bones + b] * w;     // (*blendMask)[handle] * weight
                if (!w) continue;
                const uint4 R = __ldg(A.bone_rec + b);                             // one load: key range, track, shortest-path flag
                const uint32_t k0 = R.x, nk = R.y, trk = R.z;
                if (nk == 0) continue;
                const float tp = P.tp;
                const uint32_t i2 = A.index_map[(size_t)trk * A.num_times + P.gi]; // AnimationTrack::getKeyFramesAtTime
                const float4* K = A.keys + 3 * (size_t)k0;
                // both candidate keys at once (k1 is k2 or the key before it): no load depends on the time comparison
                const uint32_t ip = i2 ? i2 - 1 : 0;
                const float4 k2a = K[3 * i2], k2b = K[3 * i2 + 1], k2c = K[3 * i2 + 2];
                const float4 kpa = K[3 * ip], kpb = K[3 * ip + 1], kpc = K[3 * ip + 2];
                const float t2 = k2a.x;
                const bool back = i2 != 0 && tp < t2;
                const float4 k1a = back ? kpa : k2a, k1b = back ? kpb : k2b, k1c = back ? kpc : k2c;
                const float t1 = k1a.x;
                const float f = (t1 == t2) ? 0.0f : (tp - t1) / (t2 - t1);
                const bool shortest = R.w != 0;
                Vec3 tr{k1a.y, k1a.z, k1a.w}, sc{k1c.x, k1c.y, k1c.z};
                Quat ro{k1b.x, k1b.y, k1b.z, k1b.w};
                if (f != 0.0f) {                                                   // getInterpolatedKeyFrame
                    if (!NONLINEAR || A.mode == 0) {                               // IM_LINEAR, RIM_LINEAR
                        ro = q_nlerp(f, ro, Quat{k2b.x, k2b.y, k2b.z, k2b.w}, shortest);
                        tr = Vec3{k1a.y + f * (k2a.y - k1a.y), k1a.z + f * (k2a.z - k1a.z), k1a.w + f * (k2a.w - k1a.w)};
                        sc = Vec3{k1c.x + f * (k2c.x - k1c.x), k1c.y + f * (k2c.y - k1c.y), k1c.z + f * (k2c.z - k1c.z)};
                    } else interpolate_nonlinear(A, k0 + (back ? ip : i2), f, shortest, tr, ro, sc);
                }
                const float scl = 1.0f;                                            // applyToNode, no linked skeleton
                L.p.x += tr.x * w * scl; L.p.y += tr.y * w * scl; L.p.z += tr.z * w * scl;
                const Quat ident{1.0f, 0.0f, 0.0f, 0.0f};
                const Quat rot = (NONLINEAR && (A.mode & 2u)) ? slerp_out_of_line(w, ident, ro, shortest) : q_nlerp(w, ident, ro, shortest);
                L.q = q_unit(q_mul(L.q, rot));
                if ((sc.x != 1.0f || sc.y != 1.0f || sc.z != 1.0f) && w != 1.0f)
                    sc = Vec3{1.0f + w * (sc.x - 1.0f), 1.0f + w * (sc.y - 1.0f), 1.0f + w * (sc.z - 1.0f)};
                L.s.x *= sc.x; L.s.y *= sc.y; L.s.z *= sc.z;
            }
        }
        store_s(p, L);
    }
    __syncthreads();

    // ---- phase 2: derived transforms in place, one tree level per round (level 0: derived = local)
    for (uint32_t l = 1; l < S.num_levels; ++l) {
        const uint32_t a = S.level_start[l], e = S.level_start[l + 1];
        for (uint32_t p = a + slot; p < e; p += SK_SLOTS) {
            const uint32_t b = S.order[p];
            const Xform D = derive(load_s((uint32_t)S.ppos[p]), load_s(p), S.inherit[2 * b] != 0, S.inherit[2 * b + 1] != 0);
            store_s(p, D);
        }
        __syncthreads();
    }
    if (!live) return;
    // ---- phase 3: Bone::_getOffsetTransform + makeTransform, 64 contiguous bytes per (instance, bone)
    for (uint32_t p = slot; p < S.num_bones; p += SK_SLOTS) {
        const uint32_t b = S.order[p];
        const Xform D = load_s(p);
        const Xform I = load_xform(S.bind_inverse + 10 * b);
        const Vec3 ls{D.s.x * I.s.x, D.s.y * I.s.y, D.s.z * I.s.z};
        const Quat lr = q_mul(D.q, I.q);
        const Vec3 rp = q_rotate(lr, Vec3{ls.x * I.p.x, ls.y * I.p.y, ls.z * I.p.z});
        const Vec3 lt{D.p.x + rp.x, D.p.y + rp.y, D.p.z + rp.z};
        const float tx = lr.x + lr.x, ty = lr.y + lr.y, tz = lr.z + lr.z;
        const float twx = tx * lr.w, twy = ty * lr.w, twz = tz * lr.w;
        const float txx = tx * lr.x, txy = ty * lr.x, txz = tz * lr.x;
        const float tyy = ty * lr.y, tyz = tz * lr.y, tzz = tz * lr.z;
        float4* o = reinterpret_cast<float4*>(out + (size_t)inst * stride_floats + 16 * (size_t)b);
        o[0] = make_float4(ls.x * (1.0f - (tyy + tzz)), ls.y * (txy - twz), ls.z * (txz + twy), lt.x);
        o[1] = make_float4(ls.x * (txy + twz), ls.y * (1.0f - (txx + tzz)), ls.z * (tyz - twx), lt.y);
        o[2] = make_float4(ls.x * (txz - twy), ls.y * (tyz + twx), ls.z * (1.0f - (txx + tyy)), lt.z);
        o[3] = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    }
}

}  // namespace ogrecu

// ---------------------------------------------------------------------------------------------- host side

struct ogrecu_skel {
    ogrecu_ctx* ctx;
    uint32_t num_bones, blend_mode, num_levels;
    std::vector<int32_t> parent;
    std::vector<float> bind_inverse;
    // device copies of the bone table
    int32_t* d_parent; uint8_t* d_level; uint16_t* d_order; uint8_t* d_inherit; float* d_bind; float* d_bind_inverse;
    int16_t* d_ppos; uint16_t* d_level_start;
    // animations: host mirror of the device records + the device table
    std::vector<AnimDev> anims;
    std::vector<void*> owned;               // device blocks of the animations
    AnimDev* d_anims; size_t d_anims_cap;
    int16_t* d_manual_slot = nullptr;       // [num_bones] of the last call with manually controlled bones
    std::vector<int16_t> manual_slot_host;
};

namespace {

template <typename T>
int upload(ogrecu_ctx* ctx, const T* host, size_t n, T** dev)
{
    *dev = nullptr;
    OGRECU_CUDA(cudaMalloc((void**)dev, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) OGRECU_CUDA(cudaMemcpyAsync(*dev, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));      // the host source may be a temporary
    return OGRECU_OK;
}

}  // namespace

extern "C" {

int ogrecu_skel_create(ogrecu_ctx* ctx, const ogrecu_skeleton_desc* d, ogrecu_skel** out)
{
    OGRECU_REQUIRE(ctx && d && out, "NULL argument");
    *out = nullptr;
    OGRECU_REQUIRE(d->num_bones >= 1 && d->num_bones <= 256, "num_bones must be 1..256");
    OGRECU_REQUIRE(d->parent && d->bind, "parent and bind tables are required");
    OGRECU_REQUIRE(d->blend_mode <= 1, "blend_mode is 0 (average) or 1 (cumulative)");
    const uint32_t n = d->num_bones;
    // tree levels (also rejects cycles and dangling parents)
    std::vector<uint8_t> level(n, 0xff);
    uint32_t done = 0, levels = 0;
    for (uint32_t b = 0; b < n; ++b) OGRECU_REQUIRE(d->parent[b] < (int32_t)n && d->parent[b] != (int32_t)b, "bad parent handle");
    for (uint32_t l = 0; done < n; ++l) {
        OGRECU_REQUIRE(l < 255, "bone hierarchy deeper than 254 levels or cyclic");
        uint32_t before = done;
        std::vector<uint32_t> now;
        for (uint32_t b = 0; b < n; ++b)
            if (level[b] == 0xff && (d->parent[b] < 0 ? l == 0 : (level[d->parent[b]] != 0xff && level[d->parent[b]] < l))) now.push_back(b);
        for (uint32_t b : now) { level[b] = (uint8_t)l; ++done; }
        OGRECU_REQUIRE(done > before, "bone hierarchy is cyclic");
        levels = l + 1;
    }
    std::vector<uint8_t> inherit(2 * n, 1);
    if (d->inherit) memcpy(inherit.data(), d->inherit, 2 * n);
    // binding-pose inverses (Bone::setBindingPose, OgreBone.cpp:66-74) unless the engine's own are given
    std::vector<float> inv(10 * (size_t)n);
    if (d->bind_inverse) {
        memcpy(inv.data(), d->bind_inverse, inv.size() * sizeof(float));
    } else {
        std::vector<Xform> der(n);
        for (uint32_t l = 0; l < levels; ++l)
            for (uint32_t b = 0; b < n; ++b) {
                if (level[b] != l) continue;
                const Xform L = load_xform(d->bind + 10 * b);
                der[b] = d->parent[b] < 0 ? L : derive(der[d->parent[b]], L, inherit[2 * b] != 0, inherit[2 * b + 1] != 0);
            }
        for (uint32_t b = 0; b < n; ++b) {
            float* o = inv.data() + 10 * (size_t)b;
            const Xform& D = der[b];
            o[0] = -D.p.x; o[1] = -D.p.y; o[2] = -D.p.z;
            const float norm = D.q.w * D.q.w + D.q.x * D.q.x + D.q.y * D.q.y + D.q.z * D.q.z;     // Quaternion::Inverse
            if (norm > 0.0f) { const float r = 1.0f / norm; o[3] = D.q.w * r; o[4] = -D.q.x * r; o[5] = -D.q.y * r; o[6] = -D.q.z * r; }
            else o[3] = o[4] = o[5] = o[6] = 0.0f;
            o[7] = 1.0f / D.s.x; o[8] = 1.0f / D.s.y; o[9] = 1.0f / D.s.z;
        }
    }
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    ogrecu_skel* s = new ogrecu_skel();
    s->ctx = ctx; s->num_bones = n; s->blend_mode = d->blend_mode; s->num_levels = levels;
    s->parent.assign(d->parent, d->parent + n);
    s->bind_inverse = inv;
    s->d_anims = nullptr; s->d_anims_cap = 0;
    int rc = upload(ctx, d->parent, n, &s->d_parent);
    if (!rc) rc = upload(ctx, level.data(), n, &s->d_level);
    std::vector<uint16_t> order(n);
    for (uint32_t b = 0; b < n; ++b) order[b] = (uint16_t)b;
    std::stable_sort(order.begin(), order.end(), [&](uint16_t x, uint16_t y) { return level[x] < level[y]; });
    if (!rc) rc = upload(ctx, order.data(), n, &s->d_order);
    std::vector<int16_t> ppos(n, -1);
    std::vector<uint16_t> pos_of(n), level_start(levels + 1, (uint16_t)n);
    for (uint32_t p = 0; p < n; ++p) pos_of[order[p]] = (uint16_t)p;
    for (uint32_t p = 0; p < n; ++p) {
        if (d->parent[order[p]] >= 0) ppos[p] = (int16_t)pos_of[d->parent[order[p]]];
        const uint8_t l = level[order[p]];
        if (p < level_start[l]) level_start[l] = (uint16_t)p;
    }
    if (!rc) rc = upload(ctx, ppos.data(), n, &s->d_ppos);
    if (!rc) rc = upload(ctx, level_start.data(), level_start.size(), &s->d_level_start);
    if (!rc) rc = upload(ctx, inherit.data(), 2 * (size_t)n, &s->d_inherit);
    if (!rc) rc = upload(ctx, d->bind, 10 * (size_t)n, &s->d_bind);
    if (!rc) rc = upload(ctx, inv.data(), inv.size(), &s->d_bind_inverse);
    if (rc) { ogrecu_skel_destroy(s); return rc; }
    *out = s;
    return OGRECU_OK;
}

int ogrecu_skel_destroy(ogrecu_skel* s)
{
    if (!s) return OGRECU_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    for (void* p : {(void*)s->d_parent, (void*)s->d_level, (void*)s->d_order, (void*)s->d_ppos, (void*)s->d_level_start, (void*)s->d_inherit, (void*)s->d_bind, (void*)s->d_bind_inverse, (void*)s->d_anims})
        if (p) cudaFree(p);
    for (void* p : s->owned) cudaFree(p);
    delete s;
    return OGRECU_OK;
}

int ogrecu_skel_set_blend_mode(ogrecu_skel* s, uint32_t mode)
{
    OGRECU_REQUIRE(s && mode <= 1, "blend_mode is 0 (average) or 1 (cumulative)");
    s->blend_mode = mode;
    return OGRECU_OK;
}

int ogrecu_skel_get_bind_inverse(ogrecu_skel* s, float* out)
{
    OGRECU_REQUIRE(s && out, "NULL argument");
    memcpy(out, s->bind_inverse.data(), s->bind_inverse.size() * sizeof(float));
    return OGRECU_OK;
}

int ogrecu_skel_add_animation(ogrecu_skel* s, const ogrecu_animation_desc* d, uint32_t* out_index)
{
    OGRECU_REQUIRE(s && d, "NULL argument");
    OGRECU_REQUIRE(d->interpolation <= 1 && d->rotation_interpolation <= 1, "interpolation is IM_LINEAR (0) or IM_SPLINE (1), rotation RIM_LINEAR (0) or RIM_SPHERICAL (1)");
    OGRECU_REQUIRE(d->num_tracks == 0 || (d->track_bone && d->key_offset && d->keys), "track tables are required");
    const uint32_t nt = d->num_tracks, nb = s->num_bones;
    const uint32_t nkeys = nt ? d->key_offset[nt] : 0;
    std::vector<int32_t> bone_track(nb, -1);
    for (uint32_t t = 0; t < nt; ++t) {
        OGRECU_REQUIRE(d->track_bone[t] >= 0 && (uint32_t)d->track_bone[t] < nb, "track names a bone outside the skeleton");
        OGRECU_REQUIRE(bone_track[d->track_bone[t]] < 0, "two tracks for one bone");
        OGRECU_REQUIRE(d->key_offset[t] <= d->key_offset[t + 1] && d->key_offset[t + 1] - d->key_offset[t] <= 65535, "bad key range");
        bone_track[d->track_bone[t]] = (int32_t)t;
        for (uint32_t k = d->key_offset[t] + 1; k < d->key_offset[t + 1]; ++k)
            OGRECU_REQUIRE(d->keys[11 * (size_t)k] >= d->keys[11 * (size_t)(k - 1)], "key frames of a track must be sorted by time");
    }
    // Animation::buildKeyFrameTimeList (OgreAnimation.cpp:620-660): the sorted set of all key times ...
    std::vector<float> times(nkeys);
    for (uint32_t k = 0; k < nkeys; ++k) times[k] = d->keys[11 * (size_t)k];
    std::sort(times.begin(), times.end());
    times.erase(std::unique(times.begin(), times.end()), times.end());
    if (times.empty()) times.push_back(0.0f);
    const uint32_t ntimes = (uint32_t)times.size();
    // ... and every track's global -> local key index map (AnimationTrack::_buildKeyFrameIndexMap, OgreAnimationTrack.cpp:190-206)
    std::vector<uint16_t> index_map((size_t)std::max(nt, 1u) * ntimes, 0);
    for (uint32_t t = 0; t < nt; ++t) {
        const uint32_t k0 = d->key_offset[t], nk = d->key_offset[t + 1] - k0;
        uint32_t i = 0;
        for (uint32_t j = 0; j < ntimes; ++j) {
            index_map[(size_t)t * ntimes + j] = (uint16_t)i;
            while (nk && i + 1 < nk && d->keys[11 * (size_t)(k0 + i)] <= times[j]) ++i;
        }
    }
    std::vector<float4> keys(3 * (size_t)std::max(nkeys, 1u));
    for (uint32_t k = 0; k < nkeys; ++k) {
        const float* f = d->keys + 11 * (size_t)k;
        keys[3 * k] = make_float4(f[0], f[1], f[2], f[3]);
        keys[3 * k + 1] = make_float4(f[4], f[5], f[6], f[7]);
        keys[3 * k + 2] = make_float4(f[8], f[9], f[10], 0.0f);
    }
    std::vector<uint8_t> shortest(std::max(nt, 1u), 1);
    if (d->track_shortest) memcpy(shortest.data(), d->track_shortest, nt);
    std::vector<uint32_t> key_offset(nt + 1, 0);
    if (nt) memcpy(key_offset.data(), d->key_offset, (nt + 1) * sizeof(uint32_t));

    ogrecu_ctx* ctx = s->ctx;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    AnimDev A{};
    A.length = d->length; A.num_times = ntimes;
    std::vector<uint4> bone_rec(nb, make_uint4(0, 0, 0, 0));
    for (uint32_t t = 0; t < nt; ++t)
        bone_rec[d->track_bone[t]] = make_uint4(key_offset[t], key_offset[t + 1] - key_offset[t], t, shortest[t] ? 1u : 0u);
    uint4* d_br;
    OGRECU_CHECK(upload(ctx, bone_rec.data(), bone_rec.size(), &d_br)); s->owned.push_back(d_br);
    float* d_times; int32_t* d_bt; uint32_t* d_ko; uint8_t* d_sh; uint16_t* d_im; float4* d_keys;
    OGRECU_CHECK(upload(ctx, times.data(), times.size(), &d_times)); s->owned.push_back(d_times);
    OGRECU_CHECK(upload(ctx, bone_track.data(), bone_track.size(), &d_bt)); s->owned.push_back(d_bt);
    OGRECU_CHECK(upload(ctx, key_offset.data(), key_offset.size(), &d_ko)); s->owned.push_back(d_ko);
    OGRECU_CHECK(upload(ctx, shortest.data(), shortest.size(), &d_sh)); s->owned.push_back(d_sh);
    OGRECU_CHECK(upload(ctx, index_map.data(), index_map.size(), &d_im)); s->owned.push_back(d_im);
    OGRECU_CHECK(upload(ctx, keys.data(), keys.size(), &d_keys)); s->owned.push_back(d_keys);
    A.mode = (d->interpolation ? 1u : 0u) | (d->rotation_interpolation ? 2u : 0u);
    A.tangents = nullptr;
    if (d->interpolation) {
        // NodeAnimationTrack::buildInterpolationSplines (OgreAnimationTrack.cpp:515-549): SimpleSpline::recalcTangents
        // (OgreSimpleSpline.cpp:156-207) for translate and scale, RotationalSpline::recalcTangents (OgreRotationSpline.cpp) with
        // Quaternion::Log / Exp / Inverse (OgreQuaternion.cpp:274-357) for the rotations — once per track, on the host, in the
        // reference's operation order (the float overloads of libm, as OgreMath.h uses them)
        std::vector<float4> tan(3 * (size_t)std::max(nkeys, 1u), make_float4(0.f, 0.f, 0.f, 0.f));
        auto KQ = [&](uint32_t k) { const float* f = d->keys + 11 * (size_t)k; return Quat{f[4], f[5], f[6], f[7]}; };
        auto q_log = [](const Quat& q) {
            Quat r{0.0f, q.x, q.y, q.z};
            if (fabsf(q.w) < 1.0) {
                const float nv = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
                const float angle = atan2f(nv, q.w);
                const float sn = sinf(angle);
                if (fabsf(sn) >= 1e-03f) { const float c = angle / sn; r.x = c * q.x; r.y = c * q.y; r.z = c * q.z; }
            }
            return r;
        };
        auto q_exp = [](const Quat& q) {
            const float angle = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
            const float sn = sinf(angle);
            const float ew = expf(q.w);
            Quat r;
            r.w = ew * cosf(angle);
            if (fabsf(angle) >= 1e-03f) { const float c = ew * (sn / angle); r.x = c * q.x; r.y = c * q.y; r.z = c * q.z; }
            else { r.x = ew * q.x; r.y = ew * q.y; r.z = ew * q.z; }
            return r;
        };
        auto q_inverse = [](const Quat& q) {
            const float n = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
            Quat r{0.0f, 0.0f, 0.0f, 0.0f};
            if (n > 0.0) { const float inv = 1.0f / n; r = Quat{q.w * inv, -q.x * inv, -q.y * inv, -q.z * inv}; }
            return r;
        };
        for (uint32_t t = 0; t < nt; ++t) {
            const uint32_t k0 = d->key_offset[t], n = d->key_offset[t + 1] - k0;
            if (n < 2) continue;
            for (int which = 0; which < 2; ++which) {                  // translate (floats 1..3) and scale (8..10)
                const int off = which ? 8 : 1;
                auto P = [&](uint32_t i, int c) { return d->keys[11 * (size_t)(k0 + i) + off + c]; };
                const bool closed = P(0, 0) == P(n - 1, 0) && P(0, 1) == P(n - 1, 1) && P(0, 2) == P(n - 1, 2);
                for (uint32_t i = 0; i < n; ++i) {
                    uint32_t a, b;
                    float4& out = tan[3 * (size_t)(k0 + i) + (which ? 2 : 0)];
                    if (i == 0) { a = 1; b = closed ? n - 2 : 0; }
                    else if (i == n - 1) { if (closed) { out = tan[3 * (size_t)k0 + (which ? 2 : 0)]; continue; } a = i; b = i - 1; }
                    else { a = i + 1; b = i - 1; }
                    out = make_float4(0.5f * (P(a, 0) - P(b, 0)), 0.5f * (P(a, 1) - P(b, 1)), 0.5f * (P(a, 2) - P(b, 2)), 0.f);
                }
            }
            const Quat first = KQ(k0), last = KQ(k0 + n - 1);
            const bool closed = first.x == last.x && first.y == last.y && first.z == last.z && first.w == last.w;
            for (uint32_t i = 0; i < n; ++i) {
                const Quat pq = KQ(k0 + i), invp = q_inverse(pq);
                Quat part1, part2;
                if (i == 0) {
                    part1 = q_log(q_mul(invp, KQ(k0 + 1)));
                    part2 = closed ? q_log(q_mul(invp, KQ(k0 + n - 2))) : q_log(q_mul(invp, pq));
                } else if (i == n - 1) {
                    part1 = closed ? q_log(q_mul(invp, KQ(k0 + 1))) : q_log(q_mul(invp, pq));
                    part2 = q_log(q_mul(invp, KQ(k0 + i - 1)));
                } else {
                    part1 = q_log(q_mul(invp, KQ(k0 + i + 1)));
                    part2 = q_log(q_mul(invp, KQ(k0 + i - 1)));
                }
                const float m = -0.25f;
                const Quat pre{m * (part1.w + part2.w), m * (part1.x + part2.x), m * (part1.y + part2.y), m * (part1.z + part2.z)};
                const Quat tq = q_mul(pq, q_exp(pre));
                tan[3 * (size_t)(k0 + i) + 1] = make_float4(tq.w, tq.x, tq.y, tq.z);
            }
        }
        float4* d_tan;
        OGRECU_CHECK(upload(ctx, tan.data(), tan.size(), &d_tan)); s->owned.push_back(d_tan);
        A.tangents = d_tan;
    }
    A.bone_rec = d_br;
    A.times = d_times; A.bone_track = d_bt; A.key_offset = d_ko; A.shortest = d_sh; A.index_map = d_im; A.keys = d_keys;
    s->anims.push_back(A);
    // the device table of records: regrown geometrically, rewritten whole (animations are added at load time)
    if (s->anims.size() > s->d_anims_cap) {
        OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
        if (s->d_anims) OGRECU_CUDA(cudaFree(s->d_anims));
        s->d_anims = nullptr;
        s->d_anims_cap = std::max<size_t>(16, 2 * s->anims.size());
        OGRECU_CUDA(cudaMalloc((void**)&s->d_anims, s->d_anims_cap * sizeof(AnimDev)));
    }
    OGRECU_CUDA(cudaMemcpyAsync(s->d_anims, s->anims.data(), s->anims.size() * sizeof(AnimDev), cudaMemcpyHostToDevice, ctx->stream));
    OGRECU_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out_index) *out_index = (uint32_t)s->anims.size() - 1;
    return OGRECU_OK;
}

int ogrecu_skel_palettes_dev(ogrecu_skel* s, uint32_t num_instances, uint32_t num_states, const int32_t* anim_dev,
                             const float* time_dev, const float* weight_dev, float* palettes_dev, size_t stride_floats)
{
    return ogrecu_skel_palettes_ex_dev(s, num_instances, num_states, anim_dev, time_dev, weight_dev, nullptr, palettes_dev, stride_floats);
}

int ogrecu_skel_palettes_ex_dev(ogrecu_skel* s, uint32_t num_instances, uint32_t num_states, const int32_t* anim_dev, const float* time_dev,
                                const float* weight_dev, const ogrecu_palette_options* opt, float* palettes_dev, size_t stride_floats)
{
    OGRECU_REQUIRE(s && palettes_dev, "NULL argument");
    if (!num_instances) return OGRECU_OK;
    PalExtra X{};
    if (opt) {
        if (opt->mask_dev) {
            OGRECU_REQUIRE(opt->mask_instance_stride_floats == 0 || opt->mask_instance_stride_floats >= (size_t)num_states * s->num_bones,
                           "blend-mask instance stride smaller than states x bones");
            X.mask = opt->mask_dev; X.mask_stride = opt->mask_instance_stride_floats;
        }
        if (opt->num_manual) {
            OGRECU_REQUIRE(opt->manual_bones && opt->manual_dev, "manual bones need their handles and their local transforms");
            OGRECU_REQUIRE(s->ctx->skel_mode == 0, "manual bones need the default palette kernel (skel_mode 0)");
            std::vector<int16_t> slot(s->num_bones, -1);
            for (uint32_t k = 0; k < opt->num_manual; ++k) {
                OGRECU_REQUIRE(opt->manual_bones[k] < s->num_bones, "manual bone handle outside the skeleton");
                OGRECU_REQUIRE(slot[opt->manual_bones[k]] < 0, "a manual bone is listed twice");
                slot[opt->manual_bones[k]] = (int16_t)k;
            }
            OGRECU_CUDA(cudaSetDevice(s->ctx->device));
            if (slot != s->manual_slot_host) {                          // the list rarely changes: uploaded when it does
                if (!s->d_manual_slot) { OGRECU_CUDA(cudaMalloc((void**)&s->d_manual_slot, (size_t)s->num_bones * 2)); s->owned.push_back(s->d_manual_slot); }
                OGRECU_CUDA(cudaStreamSynchronize(s->ctx->stream));
                s->manual_slot_host = slot;
                OGRECU_CUDA(cudaMemcpyAsync(s->d_manual_slot, s->manual_slot_host.data(), (size_t)s->num_bones * 2, cudaMemcpyHostToDevice, s->ctx->stream));
            }
            X.manual_slot = s->d_manual_slot; X.manual = opt->manual_dev; X.num_manual = opt->num_manual;
        }
        OGRECU_REQUIRE(!X.mask || s->ctx->skel_mode == 0, "blend masks need the default palette kernel (skel_mode 0)");
    }
    OGRECU_REQUIRE(num_states == 0 || (anim_dev && time_dev && weight_dev), "state arrays are required");
    OGRECU_REQUIRE(num_states <= 64, "at most 64 animation-state slots per instance");
    OGRECU_REQUIRE(stride_floats >= 16 * (size_t)s->num_bones && (stride_floats & 3) == 0, "palette stride too small or not a multiple of 4 floats");
    OGRECU_REQUIRE(((uintptr_t)palettes_dev & 15) == 0, "palettes_dev must be 16-byte aligned");
    ogrecu_ctx* ctx = s->ctx;
    OGRECU_CUDA(cudaSetDevice(ctx->device));
    SkelDev S{};
    S.num_bones = s->num_bones; S.blend_mode = s->blend_mode; S.num_levels = s->num_levels; S.num_anims = (uint32_t)s->anims.size();
    S.parent = s->d_parent; S.level = s->d_level; S.order = s->d_order; S.inherit = s->d_inherit; S.bind = s->d_bind; S.bind_inverse = s->d_bind_inverse;
    S.anims = s->d_anims; S.ppos = s->d_ppos; S.level_start = s->d_level_start;
    for (const AnimDev& a : s->anims)
        OGRECU_REQUIRE(a.mode == 0 || ctx->skel_mode == 0, "IM_SPLINE / RIM_SPHERICAL animations need the default palette kernel (skel_mode 0)");
    if (ctx->skel_mode == 0) {
        const size_t smem2 = (size_t)s->num_bones * SK_XF * sizeof(float) + (size_t)SK_INST * std::max(num_states, 1u) * sizeof(StatePre);
        bool nonlinear = false;
        for (const AnimDev& a : s->anims) nonlinear |= a.mode != 0;
        const bool extra = X.mask != nullptr || X.manual_slot != nullptr;
        const void* fn = nonlinear ? (extra ? (const void*)skel_palette_kernel2<true, true> : (const void*)skel_palette_kernel2<true, false>)
                                   : (extra ? (const void*)skel_palette_kernel2<false, true> : (const void*)skel_palette_kernel2<false, false>);
        static size_t granted[4][64] = {};
        const int which = (nonlinear ? 2 : 0) + (extra ? 1 : 0);
        if (smem2 > 48 * 1024 && smem2 > granted[which][ctx->device & 63]) {
            OGRECU_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            granted[which][ctx->device & 63] = smem2;
        }
        const dim3 grid2(div_up(num_instances, SK_INST)), block2(SK_INST, SK_SLOTS);
        void* args[] = {(void*)&S, (void*)&num_instances, (void*)&num_states, (void*)&anim_dev, (void*)&time_dev, (void*)&weight_dev,
                        (void*)&palettes_dev, (void*)&stride_floats, (void*)&X};
        OGRECU_CUDA(cudaLaunchKernel(fn, grid2, block2, args, smem2, ctx->stream));
        OGRECU_CUDA(cudaGetLastError());
        count_launch(ctx);
        return OGRECU_OK;
    }
    const uint32_t bx = (s->num_bones + 31) & ~31u;
    const uint32_t by = std::max(1u, 256u / bx);
    dim3 block(bx, by), grid(div_up(num_instances, by));
    const size_t smem = (size_t)by * s->num_bones * 10 * sizeof(float);
    skel_palette_kernel<<<grid, block, smem, ctx->stream>>>(S, num_instances, num_states, anim_dev, time_dev, weight_dev, palettes_dev,
                                                           stride_floats);
    OGRECU_CUDA(cudaGetLastError());
    count_launch(ctx);
    return OGRECU_OK;
}

}  // extern "C"
