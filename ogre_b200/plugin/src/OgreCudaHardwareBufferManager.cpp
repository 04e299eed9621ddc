#include "OgreCudaHostTiming.h"
#include "OgreCudaHardwareBufferManager.h"
#include "OgreCudaRenderSystem.h"
#include "OgreDefaultHardwareBufferManager.h"

namespace Ogre
{
CudaHardwareBuffer::CudaHardwareBuffer(ogrecu_ctx* ctx, size_t sizeInBytes, Usage usage)
    : HardwareBuffer(usage, false), mBuf(0), mDevice(0)   // the pinned shadow lives inside ogrecu_buf, not in OgreMain
{
    mSizeInBytes = sizeInBytes;
    OGRECU_THROW_IF(ogrecu_buf_create(ctx, sizeInBytes, &mBuf));
    void* d = 0;
    OGRECU_THROW_IF(ogrecu_buf_device_ptr(mBuf, &d));
    mDevice = static_cast<const uint8*>(d);
}

CudaHardwareBuffer::~CudaHardwareBuffer()
{
    if (mBuf)
    {
        CudaRenderSystem::flushPendingDraws();   // recorded draws may still reference the device copy
        ogrecu_buf_destroy(mBuf);
    }
}

static int mapMode(HardwareBuffer::LockOptions options)
{
    switch (options)
    {
    case HardwareBuffer::HBL_READ_ONLY: return OGRECU_MAP_READ;
    case HardwareBuffer::HBL_DISCARD: return OGRECU_MAP_WRITE | OGRECU_MAP_DISCARD;
    case HardwareBuffer::HBL_NO_OVERWRITE:
    case HardwareBuffer::HBL_WRITE_ONLY:   // partial writes must keep the rest of the shadow current
    case HardwareBuffer::HBL_NORMAL:
    default: return OGRECU_MAP_READ | OGRECU_MAP_WRITE;
    }
}

void* CudaHardwareBuffer::lockImpl(size_t offset, size_t length, LockOptions options)
{
    // Draws are only recorded by _render and softwareVertexSkinning calls may be queued; a lock that would let the host change
    // a buffer they reference (or read one they produce) runs them first.  libogrecu tracks which buffers recorded draws and
    // queued jobs touch (ogrecu_buf_map), so locking entity B's buffers does not flush entity A's draws.
    OGRECU_TIME(HT_LOCK);
    void* p = 0;
    OGRECU_THROW_IF(ogrecu_buf_map(mBuf, offset, length, mapMode(options), &p));
    return p;
}

void CudaHardwareBuffer::unlockImpl(void)
{
    OGRECU_TIME(HT_UNLOCK);
    OGRECU_THROW_IF(ogrecu_buf_unmap(mBuf));
}

void CudaHardwareBuffer::readData(size_t offset, size_t length, void* pDest)
{
    void* p = lock(offset, length, HBL_READ_ONLY);
    memcpy(pDest, p, length);
    unlock();
}

void CudaHardwareBuffer::writeData(size_t offset, size_t length, const void* pSource, bool discardWholeBuffer)
{
    void* p = lock(offset, length, (discardWholeBuffer && offset == 0 && length == mSizeInBytes) ? HBL_DISCARD : HBL_NORMAL);
    memcpy(p, pSource, length);
    unlock();
}

void CudaHardwareBuffer::copyData(HardwareBuffer& srcBuffer, size_t srcOffset, size_t dstOffset, size_t length,
                                  bool discardWholeBuffer)
{
    // srcBuffer is the other buffer's delegate (HardwareBuffer::copyData, OgreHardwareBuffer.h:326-336)
    if (auto other = dynamic_cast<CudaHardwareBuffer*>(&srcBuffer))
    {
        // both live on the device, whose copy is the current one in stream order (uploads are enqueued at unlock): copy there,
        // the destination's shadow is refreshed when the host next locks it
        if (!length) return;
        OgreAssert(srcOffset + length <= other->getSizeInBytes() && dstOffset + length <= mSizeInBytes, "copy out of range");
        CudaRenderSystem::flushPendingDraws();
        OGRECU_THROW_IF(ogrecu_memcpy_d2d(CudaRenderSystem::getInstance()->getContext(), const_cast<void*>(getDevicePtr(dstOffset)),
                                          other->getDevicePtr(srcOffset), length));
        OGRECU_THROW_IF(ogrecu_buf_mark_device_dirty(mBuf, dstOffset, length));
        return;
    }
    const void* src = srcBuffer.lock(srcOffset, length, HBL_READ_ONLY);
    writeData(dstOffset, length, src, discardWholeBuffer);
    srcBuffer.unlock();
}

HardwareVertexBufferSharedPtr CudaHardwareBufferManager::createVertexBuffer(size_t vertexSize, size_t numVerts,
                                                                            HardwareBuffer::Usage usage, bool)
{
    auto impl = new CudaHardwareBuffer(mCtx, vertexSize * numVerts, usage);
    auto buf = std::make_shared<HardwareVertexBuffer>(this, vertexSize, numVerts, impl);
    {
        OGRE_LOCK_MUTEX(mVertexBuffersMutex);
        mVertexBuffers.insert(buf.get());
    }
    return buf;
}

HardwareIndexBufferSharedPtr CudaHardwareBufferManager::createIndexBuffer(HardwareIndexBuffer::IndexType itype,
                                                                          size_t numIndexes, HardwareBuffer::Usage usage, bool)
{
    auto impl = new CudaHardwareBuffer(mCtx, HardwareIndexBuffer::indexSize(itype) * numIndexes, usage);
    return std::make_shared<HardwareIndexBuffer>(this, itype, numIndexes, impl);
}

HardwareBufferPtr CudaHardwareBufferManager::createUniformBuffer(size_t sizeBytes, HardwareBufferUsage, bool)
{
    // uniforms are consumed on the host (applyFixedFunctionParams); plain memory is right
    return std::make_shared<DefaultHardwareBuffer>(sizeBytes);
}
} // namespace Ogre
