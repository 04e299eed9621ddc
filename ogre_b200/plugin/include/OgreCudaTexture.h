// Mirrors RenderSystems/Tiny: TinyTexture + TinyHardwarePixelBuffer + TinyTextureManager
// (src/OgreTinyTexture.cpp:30-57, OgreTinyHardwarePixelBuffer.cpp:9-53, OgreTinyTextureManager.cpp:26-37),
// and adds what Tiny lacks: TU_RENDERTARGET textures get a real RenderTexture per slice
// (pattern: RenderSystems/GL3Plus/src/OgreGL3PlusTextureBuffer.cpp:79-102).
#ifndef __OgreCudaTexture_H__
#define __OgreCudaTexture_H__

#include "OgreCudaPrerequisites.h"
#include "OgreTexture.h"
#include "OgreTextureManager.h"
#include "OgreHardwarePixelBuffer.h"
#include "OgreRenderTexture.h"
#include "OgreImage.h"

namespace Ogre
{
class CudaTexture : public Texture
{
public:
    CudaTexture(ResourceManager* creator, const String& name, ResourceHandle handle, const String& group, bool isManual,
                ManualResourceLoader* loader, ogrecu_ctx* ctx);
    virtual ~CudaTexture();

    Image* getImage() { return &mBuffer; }
    void markDirty() { mDirty = true; }
    /// RGBA8 device texture, uploaded (or resolved from the render target) on demand
    ogrecu_tex* getDeviceTexture();
    /// render-target textures own an ogrecu_rt
    ogrecu_rt* getTarget() const { return mTarget; }
    void markTargetRendered() { mTargetDirty = true; }

protected:
    Image mBuffer;          // host staging, always PF_BYTE_RGBA (TinyTextureManager::getNativeFormat)
    ogrecu_ctx* mCtx;
    ogrecu_tex* mDevice;
    ogrecu_rt* mTarget;
    bool mDirty;            // host image newer than the device texture
    bool mTargetDirty;      // render target newer than the device texture
    void createInternalResourcesImpl(void) override;
    void freeInternalResourcesImpl(void) override;
    HardwarePixelBufferPtr createSurface(uint32 face, uint32 mipmap, uint32 width, uint32 height, uint32 depth) override;
};

class CudaHardwarePixelBuffer : public HardwarePixelBuffer
{
    PixelBox mBuffer;
    CudaTexture* mParent;
    bool mLockedForWrite;

public:
    CudaHardwarePixelBuffer(const PixelBox& data, Usage usage, CudaTexture* parent, bool renderTarget);
    PixelBox lockImpl(const Box& lockBox, LockOptions options) override;
    void unlockImpl(void) override;
    void blitFromMemory(const PixelBox& src, const Box& dstBox) override;
    void blitToMemory(const Box& srcBox, const PixelBox& dst) override;
    CudaTexture* getParent() const { return mParent; }
};

/// RenderTexture over a CudaTexture's ogrecu_rt
class CudaRenderTexture : public RenderTexture
{
public:
    CudaRenderTexture(const String& name, CudaHardwarePixelBuffer* buffer, uint32 zoffset);
    bool requiresTextureFlipping() const override { return true; }   // same convention as the window
    void getCustomAttribute(const String& name, void* pData) override;
    ogrecu_rt* getTarget() const;
    CudaTexture* getTexture() const;
};

class CudaTextureManager : public TextureManager
{
    ogrecu_ctx* mCtx;

public:
    explicit CudaTextureManager(ogrecu_ctx* ctx);
    virtual ~CudaTextureManager();
    PixelFormat getNativeFormat(TextureType ttype, PixelFormat format, int usage) override;

protected:
    Resource* createImpl(const String& name, ResourceHandle handle, const String& group, bool isManual,
                         ManualResourceLoader* loader, const NameValuePairList* createParams) override;
};
} // namespace Ogre
#endif
