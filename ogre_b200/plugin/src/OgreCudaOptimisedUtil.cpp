#include "OgreCudaHostTiming.h"
#include "OgreCudaOptimisedUtil.h"
#include "OgreCudaRenderSystem.h"
#include "OgreEdgeListBuilder.h"
#include "OgreMatrix4.h"
#include "OgreVector.h"

namespace Ogre
{
OptimisedUtil* CudaOptimisedUtil::msPrevious = 0;
CudaOptimisedUtil CudaOptimisedUtil::msInstance;

static_assert(sizeof(Affine3) == 64, "Affine3 must be 16 floats");
static_assert(sizeof(EdgeData::Triangle) == 32, "EdgeData::Triangle must be 32 bytes");
static_assert(sizeof(Vector4) == 16, "Vector4 must be 4 floats");

void CudaOptimisedUtil::install()
{
    if (msImplementation == &msInstance) return;
    msPrevious = msImplementation;
    msImplementation = &msInstance;
}

void CudaOptimisedUtil::uninstall()
{
    if (msImplementation == &msInstance) msImplementation = msPrevious;
}

ogrecu_ctx* CudaOptimisedUtil::ctx(bool flush) const
{
    CudaRenderSystem* rs = CudaRenderSystem::getInstance();
    OgreAssert(rs, "CUDA render system is not installed");
    // results land in buffers that recorded draws may read: run those first.  (softwareVertexSkinning does not ask for it:
    // libogrecu checks its destination against the recorded draws itself and queues the call, see "optutil_batch".)
    if (flush) CudaRenderSystem::flushPendingDraws();
    return rs->getContext();
}

// OptimisedUtil functions return void and cannot fail in the reference; a device failure here is
// fatal by design (no CPU fallback), reported as an OGRE exception.
void CudaOptimisedUtil::softwareVertexSkinning(const float* srcPosPtr, float* destPosPtr, const float* srcNormPtr,
                                               float* destNormPtr, const float* blendWeightPtr,
                                               const unsigned char* blendIndexPtr, const Affine3* const* blendMatrices,
                                               size_t srcPosStride, size_t destPosStride, size_t srcNormStride,
                                               size_t destNormStride, size_t blendWeightStride, size_t blendIndexStride,
                                               size_t numWeightsPerVertex, size_t numVertices)
{
    OGRECU_TIME(HT_SKIN);
    OGRECU_THROW_IF(ogrecu_software_vertex_skinning_host(
        ctx(false), srcPosPtr, destPosPtr, srcNormPtr, destNormPtr, blendWeightPtr, blendIndexPtr,
        reinterpret_cast<const float* const*>(blendMatrices), srcPosStride, destPosStride, srcNormStride, destNormStride,
        blendWeightStride, blendIndexStride, numWeightsPerVertex, numVertices));
}

void CudaOptimisedUtil::softwareVertexMorph(float t, const float* srcPos1, const float* srcPos2, float* dstPos,
                                            size_t pos1VSize, size_t pos2VSize, size_t dstVSize, size_t numVertices,
                                            bool morphNormals)
{
    OGRECU_TIME(HT_OTHER_UTIL);
    OGRECU_THROW_IF(ogrecu_software_vertex_morph_host(ctx(), t, srcPos1, srcPos2, dstPos, pos1VSize, pos2VSize, dstVSize,
                                                      numVertices, morphNormals ? 1 : 0));
}

void CudaOptimisedUtil::concatenateAffineMatrices(const Affine3& baseMatrix, const Affine3* srcMatrices,
                                                  Affine3* dstMatrices, size_t numMatrices)
{
    OGRECU_THROW_IF(ogrecu_concatenate_affine_matrices_host(ctx(), baseMatrix[0], srcMatrices[0][0], dstMatrices[0][0],
                                                            numMatrices));
}

void CudaOptimisedUtil::calculateFaceNormals(const float* positions, const EdgeData::Triangle* triangles,
                                             Vector4* faceNormals, size_t numTriangles)
{
    OGRECU_THROW_IF(ogrecu_calculate_face_normals_host(ctx(), positions, 0, triangles, faceNormals->ptr(), numTriangles));
}

void CudaOptimisedUtil::calculateLightFacing(const Vector4& lightPos, const Vector4* faceNormals, char* lightFacings,
                                             size_t numFaces)
{
    OGRECU_THROW_IF(ogrecu_calculate_light_facing_host(ctx(), lightPos.ptr(), faceNormals->ptr(), lightFacings, numFaces));
}

void CudaOptimisedUtil::extrudeVertices(const Vector4& lightPos, Real extrudeDist, const float* srcPositions,
                                        float* destPositions, size_t numVertices)
{
    OGRECU_THROW_IF(ogrecu_extrude_vertices_host(ctx(), lightPos.ptr(), extrudeDist, srcPositions, destPositions, numVertices));
}
} // namespace Ogre
