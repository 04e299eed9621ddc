#include "OgreCudaSkeletonPalettes.h"
#include "OgreCudaRenderSystem.h"
#include "OgreAnimation.h"
#include "OgreAnimationTrack.h"
#include "OgreBone.h"
#include "OgreKeyFrame.h"

#include <cstring>

namespace Ogre
{
static_assert(sizeof(Affine3) == 64, "Affine3 must be 16 floats");

static void put10(float* o, const Vector3& p, const Quaternion& q, const Vector3& s)
{
    o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = q.w; o[4] = q.x; o[5] = q.y; o[6] = q.z; o[7] = s.x; o[8] = s.y; o[9] = s.z;
}

CudaSkeletonPalettes::CudaSkeletonPalettes(const Skeleton* sk)
    : mCtx(0), mSkel(0), mNumBones(sk->getNumBones()), mStates(0), mStatesCap(0), mPalettes(0), mPalettesCap(0)
{
    CudaRenderSystem* rs = CudaRenderSystem::getInstance();
    OgreAssert(rs, "CUDA render system is not installed");
    mCtx = rs->getContext();
    const size_t n = mNumBones;
    std::vector<int32> parent(n);
    std::vector<float> bind(10 * n), inv(10 * n);
    std::vector<uint8> inherit(2 * n);
    for (unsigned short h = 0; h < mNumBones; ++h)
    {
        Bone* b = sk->getBone(h);
        Node* p = b->getParent();
        parent[h] = p ? (int32) static_cast<Bone*>(p)->getHandle() : -1;
        put10(&bind[10 * h], b->getInitialPosition(), b->getInitialOrientation(), b->getInitialScale());
        inherit[2 * h] = b->getInheritOrientation();
        inherit[2 * h + 1] = b->getInheritScale();
        // the engine's own Bone::setBindingPose products (OgreBone.cpp:66-74); the library would compute the same
        put10(&inv[10 * h], b->_getBindingPoseInversePosition(), b->_getBindingPoseInverseOrientation(), b->_getBindingPoseInverseScale());
    }
    ogrecu_skeleton_desc d;
    d.num_bones = mNumBones;
    d.blend_mode = sk->getBlendMode() == ANIMBLEND_CUMULATIVE ? 1u : 0u;
    d.parent = parent.data(); d.bind = bind.data(); d.inherit = inherit.data(); d.bind_inverse = inv.data();
    OGRECU_THROW_IF(ogrecu_skel_create(mCtx, &d, &mSkel));

    for (unsigned short a = 0; a < sk->getNumAnimations(); ++a)
    {
        Animation* an = sk->getAnimation(a);
        std::vector<int32> trackBone;
        std::vector<uint8> shortest;
        std::vector<uint32> keyOffset(1, 0);
        std::vector<float> keys;
        for (auto& kv : an->_getNodeTrackList())
        {
            NodeAnimationTrack* tr = kv.second;
            trackBone.push_back(kv.first);
            shortest.push_back(tr->getUseShortestRotationPath());
            for (unsigned short i = 0; i < tr->getNumKeyFrames(); ++i)
            {
                TransformKeyFrame* kf = tr->getNodeKeyFrame(i);
                const size_t at = keys.size();
                keys.resize(at + 11);
                keys[at] = kf->getTime();
                const Vector3& t = kf->getTranslate(); const Quaternion& q = kf->getRotation(); const Vector3& s = kf->getScale();
                float* o = &keys[at + 1];
                put10(o, t, q, s);
            }
            keyOffset.push_back((uint32)(keys.size() / 11));
        }
        ogrecu_animation_desc ad;
        ad.length = an->getLength();
        ad.interpolation = an->getInterpolationMode() == Animation::IM_LINEAR ? 0u : 1u;
        ad.rotation_interpolation = an->getRotationInterpolationMode() == Animation::RIM_LINEAR ? 0u : 1u;
        ad.num_tracks = (uint32)trackBone.size();
        ad.track_bone = trackBone.data(); ad.track_shortest = shortest.data(); ad.key_offset = keyOffset.data(); ad.keys = keys.data();
        uint32 index = 0;
        OGRECU_THROW_IF(ogrecu_skel_add_animation(mSkel, &ad, &index));
        mAnimIndex[an->getName()] = (int32)index;
    }
}

CudaSkeletonPalettes::~CudaSkeletonPalettes()
{
    if (mStates) ogrecu_free(mCtx, mStates);
    if (mMasks) ogrecu_free(mCtx, mMasks);
    if (mPalettes) ogrecu_free(mCtx, mPalettes);
    ogrecu_skel_destroy(mSkel);
}

void CudaSkeletonPalettes::run(const AnimationStateSet* const* sets, size_t count)
{
    // widest enabled list decides the slot count; unused slots carry animation -1
    size_t slots = 1;
    for (size_t i = 0; i < count; ++i) slots = std::max(slots, sets[i]->getEnabledAnimationStates().size());
    const size_t n = count * slots;
    mHostStates.assign(n * 12, 0);
    int32* anim = reinterpret_cast<int32*>(mHostStates.data());
    float* time = reinterpret_cast<float*>(mHostStates.data() + n * 4);
    float* weight = reinterpret_cast<float*>(mHostStates.data() + n * 8);
    // AnimationState blend masks (Animation::apply with a mask, OgreAnimation.cpp:344-357): [count][slots][bones], ones where a
    // state has none; only built when some state carries one
    bool anyMask = false;
    for (size_t i = 0; i < count && !anyMask; ++i)
        for (const AnimationState* st : sets[i]->getEnabledAnimationStates()) anyMask |= st->hasBlendMask();
    if (anyMask) mHostMasks.assign(n * mNumBones, 1.0f);
    for (size_t i = 0; i < count; ++i)
    {
        size_t s = 0;
        for (const AnimationState* st : sets[i]->getEnabledAnimationStates())
        {
            if (anyMask && st->hasBlendMask())
            {
                const AnimationState::BoneBlendMask* m = st->getBlendMask();
                OgreAssert(m->size() >= mNumBones, "blend mask smaller than the skeleton");
                memcpy(&mHostMasks[(i * slots + s) * mNumBones], m->data(), mNumBones * sizeof(float));
            }
            auto it = mAnimIndex.find(st->getAnimationName());
            // "tolerate state entries for animations we're not aware of" (OgreSkeleton.cpp:203)
            anim[i * slots + s] = it == mAnimIndex.end() ? -1 : it->second;
            time[i * slots + s] = st->getTimePosition();
            weight[i * slots + s] = st->getWeight();
            ++s;
        }
        for (; s < slots; ++s) anim[i * slots + s] = -1;
    }
    if (mHostStates.size() > mStatesCap)
    {
        if (mStates) OGRECU_THROW_IF(ogrecu_free(mCtx, mStates));
        mStates = 0;
        mStatesCap = mHostStates.size() * 2;
        OGRECU_THROW_IF(ogrecu_malloc(mCtx, mStatesCap, &mStates));
    }
    const size_t palBytes = count * mNumBones * sizeof(Affine3);
    if (palBytes > mPalettesCap)
    {
        if (mPalettes) OGRECU_THROW_IF(ogrecu_free(mCtx, mPalettes));
        mPalettes = 0;
        mPalettesCap = palBytes * 2;
        OGRECU_THROW_IF(ogrecu_malloc(mCtx, mPalettesCap, &mPalettes));
    }
    OGRECU_THROW_IF(ogrecu_memcpy_h2d(mCtx, mStates, mHostStates.data(), mHostStates.size()));
    uint8* base = static_cast<uint8*>(mStates);
    ogrecu_palette_options opt;
    memset(&opt, 0, sizeof opt);
    if (anyMask)
    {
        const size_t maskBytes = mHostMasks.size() * sizeof(float);
        if (maskBytes > mMasksCap)
        {
            if (mMasks) OGRECU_THROW_IF(ogrecu_free(mCtx, mMasks));
            mMasks = 0;
            mMasksCap = maskBytes * 2;
            OGRECU_THROW_IF(ogrecu_malloc(mCtx, mMasksCap, &mMasks));
        }
        OGRECU_THROW_IF(ogrecu_memcpy_h2d(mCtx, mMasks, mHostMasks.data(), maskBytes));
        opt.mask_dev = static_cast<const float*>(mMasks);
        opt.mask_instance_stride_floats = slots * mNumBones;
    }
    OGRECU_THROW_IF(ogrecu_skel_palettes_ex_dev(mSkel, (uint32)count, (uint32)slots, reinterpret_cast<const int32_t*>(base),
                                                reinterpret_cast<const float*>(base + n * 4), reinterpret_cast<const float*>(base + n * 8),
                                                anyMask ? &opt : NULL, static_cast<float*>(mPalettes), (size_t)mNumBones * 16));
}

void CudaSkeletonPalettes::getBoneMatrices(const AnimationStateSet* const* sets, size_t count, Affine3* dst)
{
    if (!count) return;
    run(sets, count);
    OGRECU_THROW_IF(ogrecu_memcpy_d2h(mCtx, dst, mPalettes, count * mNumBones * sizeof(Affine3)));   // waits for the copy
}

const float* CudaSkeletonPalettes::getBoneMatricesDevice(const AnimationStateSet* const* sets, size_t count)
{
    if (!count) return 0;
    run(sets, count);
    OGRECU_THROW_IF(ogrecu_sync(mCtx));      // mHostStates may be rewritten by the next call
    return static_cast<const float*>(mPalettes);
}
} // namespace Ogre
