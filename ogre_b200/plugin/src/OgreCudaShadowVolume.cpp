#include "OgreCudaShadowVolume.h"
#include "OgreCudaHardwareBufferManager.h"
#include "OgreCudaRenderSystem.h"

#include <vector>

namespace Ogre
{
static_assert(sizeof(Vector4) == 16, "Vector4 must be 4 floats");

CudaShadowVolume::CudaShadowVolume(const EdgeData* ed) : mCtx(0), mEdges(0), mNumGroups(0), mNumTris(0), mFaceNormals(0), mFacings(0)
{
    CudaRenderSystem* rs = CudaRenderSystem::getInstance();
    OgreAssert(rs, "CUDA render system is not installed");
    mCtx = rs->getContext();
    mNumGroups = (uint32)ed->edgeGroups.size();
    mNumTris = (uint32)ed->triangles.size();
    std::vector<uint32> groups, edges, tris(3 * (size_t)mNumTris);
    for (const EdgeData::EdgeGroup& eg : ed->edgeGroups)
    {
        const uint32 start = (uint32)(edges.size() / 5);
        const uint32 g[5] = {start, (uint32)eg.edges.size(), eg.triStart, eg.triCount, (uint32)eg.vertexData->vertexCount};
        groups.insert(groups.end(), g, g + 5);
        for (const EdgeData::Edge& e : eg.edges)
        {
            const uint32 q[5] = {e.triIndex[0], e.triIndex[1], e.vertIndex[0], e.vertIndex[1], e.degenerate ? 1u : 0u};
            edges.insert(edges.end(), q, q + 5);
        }
    }
    for (uint32 t = 0; t < mNumTris; ++t)
        for (int k = 0; k < 3; ++k) tris[3 * (size_t)t + k] = ed->triangles[t].vertIndex[k];
    OGRECU_THROW_IF(ogrecu_edges_create(mCtx, mNumGroups, groups.data(), edges.data(), tris.data(), mNumTris, &mEdges));
    OGRECU_THROW_IF(ogrecu_malloc(mCtx, (size_t)mNumTris * sizeof(Vector4), &mFaceNormals));
    OGRECU_THROW_IF(ogrecu_malloc(mCtx, mNumTris, &mFacings));
    updateFaceNormals(ed);
}

CudaShadowVolume::~CudaShadowVolume()
{
    ogrecu_edges_destroy(mEdges);
    if (mFaceNormals) ogrecu_free(mCtx, mFaceNormals);
    if (mFacings) ogrecu_free(mCtx, mFacings);
}

void CudaShadowVolume::updateFaceNormals(const EdgeData* ed)
{
    OgreAssert(ed->triangleFaceNormals.size() == mNumTris, "edge data changed shape");
    if (!mNumTris) return;
    OGRECU_THROW_IF(ogrecu_memcpy_h2d(mCtx, mFaceNormals, ed->triangleFaceNormals.data(), (size_t)mNumTris * sizeof(Vector4)));
    OGRECU_THROW_IF(ogrecu_sync(mCtx));
}

void CudaShadowVolume::generate(const HardwareIndexBufferSharedPtr& indexBuffer, size_t& indexBufferUsedSize, const Light* light,
                                const AxisAlignedBox& bounds, ShadowRenderableList& shadowRenderables, unsigned long flags)
{
    OgreAssert(shadowRenderables.size() == mNumGroups, "edge groups should be 1:1 with shadow renderables");
    CudaHardwareBuffer* ib = dynamic_cast<CudaHardwareBuffer*>(indexBuffer->_getImpl<HardwareBuffer>());
    if (!ib) OGRE_EXCEPT(Exception::ERR_INVALIDPARAMS, "the index buffer does not belong to the CUDA render system");
    CudaRenderSystem::flushPendingDraws();                       // recorded draws may still read the shared index buffer
    const Vector4 l4 = light->getAs4DVector();
    const float light4[4] = {l4.x, l4.y, l4.z, l4.w};
    // EdgeData::updateTriangleLightFacing -> OptimisedUtil::calculateLightFacing, here on the resident face normals
    OGRECU_THROW_IF(ogrecu_light_facing_dev(mCtx, light4, static_cast<const float*>(mFaceNormals), static_cast<char*>(mFacings), mNumTris));
    const bool directional = light->getType() == Light::LT_DIRECTIONAL;
    const Vector3& mn = bounds.getMinimum(); const Vector3& mx = bounds.getMaximum();
    const Vector3 lp = light->getDerivedPosition();
    const float b6[6] = {mn.x, mn.y, mn.z, mx.x, mx.y, mx.z};
    const float lpos4[4] = {lp.x, lp.y, lp.z, directional ? 0.0f : 1.0f};     // the McGuire test looks at the derived position
    const int mcguire = ogrecu_shadow_use_mcguire(mNumGroups, b6, lpos4);
    bool separate = false;
    for (ShadowRenderable* sr : shadowRenderables) separate = separate || sr->isLightCapSeparate();
    std::vector<uint32> ranges(4 * (size_t)mNumGroups);
    uint32 used = (uint32)indexBufferUsedSize, emitted = 0;
    void* dev = const_cast<void*>(ib->getDevicePtr(0));
    OGRECU_THROW_IF(ogrecu_shadow_volume_dev(mEdges, static_cast<const char*>(mFacings), directional ? 1 : 0, mcguire, (uint32)flags, separate ? 1 : 0,
                                             (uint32)indexBuffer->getIndexSize(), dev, (uint32)indexBuffer->getNumIndexes(), &used, ranges.data(), &emitted));
    if (emitted)
        OGRECU_THROW_IF(ogrecu_buf_mark_device_dirty(ib->getHandle(), (size_t)(used - emitted) * indexBuffer->getIndexSize(),
                                                     (size_t)emitted * indexBuffer->getIndexSize()));
    size_t g = 0;
    for (ShadowRenderable* sr : shadowRenderables)
    {
        IndexData* id = sr->getRenderOperationForUpdate()->indexData;
        if (id->indexBuffer != indexBuffer)
        {
            sr->rebindIndexBuffer(indexBuffer);
            id = sr->getRenderOperationForUpdate()->indexData;
        }
        id->indexStart = ranges[4 * g]; id->indexCount = ranges[4 * g + 1];
        if ((flags & SRF_INCLUDE_LIGHT_CAP) && sr->isLightCapSeparate())
        {
            IndexData* ic = sr->getLightCapRenderable()->getRenderOperationForUpdate()->indexData;
            ic->indexStart = ranges[4 * g + 2]; ic->indexCount = ranges[4 * g + 3];
        }
        ++g;
    }
    indexBufferUsedSize = used;
}
} // namespace Ogre
