// Pose animation on the device: the plugin-side mirror of the VAT_POSE path of Entity::_updateAnimation
// (OgreMain/src/OgreEntity.cpp:995-1070: initialisePoseVertexData :1164-1204, VertexAnimationTrack::applyToVertexData's pose
// branch OgreAnimationTrack.cpp:840-907 -> applyPoseToVertexData :909-942 -> Mesh::softwareVertexPoseBlend OgreMesh.cpp:1948-2007,
// finalisePoseNormals OgreEntity.cpp:1207-1261) over the C ABI (ogrecu_pose_*).  Those reference functions are static /
// non-virtual, so OGRE cannot be redirected to this class without a source change (INTEGRATION.md shows the three call sites);
// an application that drives its own vertex animation calls it directly.  Built once per Mesh: every Pose's offset and
// normal maps are flattened and uploaded.  The destination VertexData's buffers must come from the CUDA render system's
// buffer manager; they stay on the device (marked device-dirty) for _render or the next OptimisedUtil call.
#ifndef __OgreCudaPoseAnimation_H__
#define __OgreCudaPoseAnimation_H__

#include "OgreCudaPrerequisites.h"
#include "OgreIteratorWrapper.h"
#include "OgreKeyFrame.h"
#include "OgreMesh.h"
#include "OgreVertexIndexData.h"

namespace Ogre
{
class _OgreCudaExport CudaPoseAnimation
{
public:
    explicit CudaPoseAnimation(const Mesh* mesh);
    ~CudaPoseAnimation();

    /// Entity::initialisePoseVertexData: dest's position buffer = src's, normals zeroed when animateNormals
    void initialise(const VertexData* src, VertexData* dest, bool animateNormals);
    /// VertexAnimationTrack::applyPoseToVertexData (software): dest += pose * influence, positions and (when the pose carries
    /// normals and dest has them) normals; nothing at all when influence == 0
    void applyPose(ushort poseIndex, VertexData* dest, float influence);
    /// the pose branch of VertexAnimationTrack::applyToVertexData: the references of kf1 with their influence interpolated
    /// towards kf2's (lerp by t), then kf2's references that kf1 lacks (t * influence), each scaled by `weight`
    void applyKeyFrames(const VertexPoseKeyFrame* kf1, const VertexPoseKeyFrame* kf2, float t, float weight, VertexData* dest);
    /// Entity::finalisePoseNormals
    void finalise(const VertexData* src, VertexData* dest);

    size_t getNumPoses() const { return mPoses.size(); }

private:
    struct DevicePose { uint32 numOffsets, numNormals; void* offsetIndex; void* offsets; void* normalIndex; void* normals; };
    ogrecu_ctx* mCtx;
    std::vector<DevicePose> mPoses;
};
} // namespace Ogre
#endif
