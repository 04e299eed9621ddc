// HardwareBufferManager + HardwareBuffer delegate of the CUDA render system.
// Mirrors OgreMain's DefaultHardwareBufferManager (OgreMain/src/OgreDefaultHardwareBufferManager.cpp:33-96)
// with the delegate pattern of RenderSystems/GL3Plus/src/OgreGL3PlusHardwareBufferManager.cpp:56-80.
#ifndef __OgreCudaHardwareBufferManager_H__
#define __OgreCudaHardwareBufferManager_H__

#include "OgreCudaPrerequisites.h"
#include "OgreHardwareBufferManager.h"
#include "OgreHardwareBuffer.h"

namespace Ogre
{
/** Device-resident buffer with a pinned host shadow (ogrecu_buf).  lock() hands out the shadow —
    the pointer stays valid for the life of the buffer, which is what TinyRenderSystem::getData
    (OgreTinyRenderSystem.cpp:335-347) and Mesh::softwareVertexBlend (OgreMesh.cpp:1827-1896) rely on —
    unlock() after a write schedules the upload on the context stream. */
class CudaHardwareBuffer : public HardwareBuffer
{
    ogrecu_buf* mBuf;
    const uint8* mDevice;        // device address of byte 0 (fixed for the buffer's life)
    void* lockImpl(size_t offset, size_t length, LockOptions options) override;
    void unlockImpl(void) override;

public:
    CudaHardwareBuffer(ogrecu_ctx* ctx, size_t sizeInBytes, Usage usage);
    ~CudaHardwareBuffer();
    void readData(size_t offset, size_t length, void* pDest) override;
    void writeData(size_t offset, size_t length, const void* pSource, bool discardWholeBuffer = false) override;
    void copyData(HardwareBuffer& srcBuffer, size_t srcOffset, size_t dstOffset, size_t length,
                  bool discardWholeBuffer = false) override;
    bool isSystemMemory(void) const override { return false; }

    ogrecu_buf* getHandle() const { return mBuf; }
    /// device address of byte `offset`
    const void* getDevicePtr(size_t offset = 0) const { return mDevice + offset; }
};

class CudaHardwareBufferManager : public HardwareBufferManager
{
    ogrecu_ctx* mCtx;

public:
    explicit CudaHardwareBufferManager(ogrecu_ctx* ctx) : mCtx(ctx) {}
    ~CudaHardwareBufferManager()
    {
        destroyAllDeclarations();
        destroyAllBindings();
    }
    HardwareVertexBufferSharedPtr createVertexBuffer(size_t vertexSize, size_t numVerts, HardwareBuffer::Usage usage,
                                                     bool useShadowBuffer = false) override;
    HardwareIndexBufferSharedPtr createIndexBuffer(HardwareIndexBuffer::IndexType itype, size_t numIndexes,
                                                   HardwareBuffer::Usage usage, bool useShadowBuffer = false) override;
    HardwareBufferPtr createUniformBuffer(size_t sizeBytes, HardwareBufferUsage usage = HBU_CPU_ONLY,
                                          bool useShadowBuffer = false) override;
};
} // namespace Ogre
#endif
