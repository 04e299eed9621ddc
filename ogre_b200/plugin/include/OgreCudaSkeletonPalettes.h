// Bone-palette production for many instances of one skeleton on the device: the plugin-side mirror of
// SkeletonInstance::setAnimationState + Skeleton::_getBoneMatrices (OgreMain/src/OgreSkeleton.cpp:162-217,402-424,
// as Entity::cacheBoneMatrices calls them, OgreEntity.cpp:1307-1330).  Built once from an Ogre::Skeleton (bone
// table, binding pose, every animation's node tracks); evaluates any number of AnimationStateSets in one launch
// through ogrecu_skel_palettes_dev.  Results are bit-identical to the engine's.
#ifndef __OgreCudaSkeletonPalettes_H__
#define __OgreCudaSkeletonPalettes_H__

#include "OgreCudaPrerequisites.h"
#include "OgreAnimationState.h"
#include "OgreMatrix4.h"
#include "OgreSkeleton.h"

#include <map>
#include <vector>

namespace Ogre
{
class _OgreCudaExport CudaSkeletonPalettes
{
public:
    /// uploads the skeleton with all its animations (IM_LINEAR / RIM_LINEAR: bit-identical palettes; IM_SPLINE / RIM_SPHERICAL:
    /// within 2e-5 of a matrix's largest element, the device's sinf / atan2f stand in for libm's); throws ERR_RENDERINGAPI_ERROR
    /// for what the device path does not cover — there is no CPU path to fall back to
    explicit CudaSkeletonPalettes(const Skeleton* skeleton);
    ~CudaSkeletonPalettes();

    unsigned short getNumBones() const { return mNumBones; }

    /// `count` state sets (one per entity instance) -> count * getNumBones() matrices in `dst` (host memory):
    /// for each set what skeletonInstance->setAnimationState(*set); skeletonInstance->_getBoneMatrices(dst) gives.
    /// Enabled states are applied in the set's enabled order, with their blend masks.  Manually controlled bones are a
    /// property of a SkeletonInstance, which this call does not see: ogrecu_skel_palettes_ex_dev takes them.
    void getBoneMatrices(const AnimationStateSet* const* sets, size_t count, Affine3* dst);

    /// same, leaving the palettes in device memory for the skinning kernels (ogrecu_skin_instanced_dev /
    /// ogrecu_anim_instanced_dev read exactly this layout); the pointer stays valid until the next call
    const float* getBoneMatricesDevice(const AnimationStateSet* const* sets, size_t count);

private:
    void run(const AnimationStateSet* const* sets, size_t count);

    ogrecu_ctx* mCtx;
    ogrecu_skel* mSkel;
    unsigned short mNumBones;
    std::map<String, int32> mAnimIndex;
    // device scratch, grown on demand
    void* mStates; size_t mStatesCap;
    void* mPalettes; size_t mPalettesCap;
    std::vector<uint8> mHostStates;
    std::vector<float> mHostMasks;           // blend masks of the last call, [count][slots][bones]
    void* mMasks = 0; size_t mMasksCap = 0;
};
} // namespace Ogre
#endif
