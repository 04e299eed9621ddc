// Shadow-volume index generation on the device: the plugin-side mirror of ShadowCaster::updateEdgeListLightFacing +
// ShadowCaster::generateShadowVolume (OgreMain/src/OgreShadowCaster.cpp:128-132,191-515) over the C ABI
// (ogrecu_light_facing_dev, ogrecu_edges_*, ogrecu_shadow_volume_dev).  Built once per EdgeData; every call takes the
// arguments generateShadowVolume takes and leaves the same indices in the index buffer and the same ranges in the
// ShadowRenderables.
#ifndef __OgreCudaShadowVolume_H__
#define __OgreCudaShadowVolume_H__

#include "OgreCudaPrerequisites.h"
#include "OgreAxisAlignedBox.h"
#include "OgreEdgeListBuilder.h"
#include "OgreHardwareIndexBuffer.h"
#include "OgreLight.h"
#include "OgreShadowCaster.h"

namespace Ogre
{
class _OgreCudaExport CudaShadowVolume
{
public:
    /// uploads the edge groups, edges, triangles and face normals of `edgeData`
    explicit CudaShadowVolume(const EdgeData* edgeData);
    ~CudaShadowVolume();

    /// after EdgeData::updateFaceNormals (vertex animation): refresh the device copy of the face normals
    void updateFaceNormals(const EdgeData* edgeData);

    /// EdgeData::updateTriangleLightFacing + generateShadowVolume.  `indexBuffer` must come from the CUDA render system's
    /// buffer manager (16- or 32-bit); `lightCapBounds` = the caster's getLightCapBounds().
    void generate(const HardwareIndexBufferSharedPtr& indexBuffer, size_t& indexBufferUsedSize, const Light* light,
                  const AxisAlignedBox& lightCapBounds, ShadowRenderableList& shadowRenderables, unsigned long flags);

private:
    ogrecu_ctx* mCtx;
    ogrecu_edges* mEdges;
    uint32 mNumGroups, mNumTris;
    void* mFaceNormals;     // device, numTris x Vector4
    void* mFacings;         // device, numTris x char
};
} // namespace Ogre
#endif
