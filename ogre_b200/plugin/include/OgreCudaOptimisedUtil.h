// CUDA-backed Ogre::OptimisedUtil (OgreMain/include/OgreOptimisedUtil.h:47-207): the six virtuals
// normally served by OgreOptimisedUtilSSE.cpp forward to the *_host entry points of libogrecu.so.
// Installed by swapping the protected static OptimisedUtil::msImplementation (.h:57) from a derived
// class, so stock libOgreMain.so needs no change; every call site goes through getImplementation().
#ifndef __OgreCudaOptimisedUtil_H__
#define __OgreCudaOptimisedUtil_H__

#include "OgreCudaPrerequisites.h"
#include "OgreOptimisedUtil.h"

namespace Ogre
{
class _OgreCudaExport CudaOptimisedUtil : public OptimisedUtil
{
    static OptimisedUtil* msPrevious;
    static CudaOptimisedUtil msInstance;
    ogrecu_ctx* ctx(bool flush = true) const;

public:
    /// make this the process-wide implementation; remembers the one it replaces
    static void install();
    /// restore the implementation that was active before install()
    static void uninstall();
    static OptimisedUtil* getPrevious() { return msPrevious; }

    void softwareVertexSkinning(const float* srcPosPtr, float* destPosPtr, const float* srcNormPtr, float* destNormPtr,
                                const float* blendWeightPtr, const unsigned char* blendIndexPtr,
                                const Affine3* const* blendMatrices, size_t srcPosStride, size_t destPosStride,
                                size_t srcNormStride, size_t destNormStride, size_t blendWeightStride,
                                size_t blendIndexStride, size_t numWeightsPerVertex, size_t numVertices) override;
    void softwareVertexMorph(float t, const float* srcPos1, const float* srcPos2, float* dstPos, size_t pos1VSize,
                             size_t pos2VSize, size_t dstVSize, size_t numVertices, bool morphNormals) override;
    void concatenateAffineMatrices(const Affine3& baseMatrix, const Affine3* srcMatrices, Affine3* dstMatrices,
                                   size_t numMatrices) override;
    void calculateFaceNormals(const float* positions, const EdgeData::Triangle* triangles, Vector4* faceNormals,
                              size_t numTriangles) override;
    void calculateLightFacing(const Vector4& lightPos, const Vector4* faceNormals, char* lightFacings,
                              size_t numFaces) override;
    void extrudeVertices(const Vector4& lightPos, Real extrudeDist, const float* srcPositions, float* destPositions,
                         size_t numVertices) override;
};
} // namespace Ogre
#endif
