#include "OgreCudaTexture.h"
#include "OgreCudaRenderSystem.h"
#include "OgreRoot.h"
#include "OgreResourceGroupManager.h"

namespace Ogre
{
CudaTexture::CudaTexture(ResourceManager* creator, const String& name, ResourceHandle handle, const String& group,
                         bool isManual, ManualResourceLoader* loader, ogrecu_ctx* ctx)
    : Texture(creator, name, handle, group, isManual, loader), mCtx(ctx), mDevice(0), mTarget(0), mDirty(true),
      mTargetDirty(false)
{
    mMipmapsHardwareGenerated = false;
}

CudaTexture::~CudaTexture()
{
    // have to call this here rather than in Resource destructor (virtual call)
    unload();
}

HardwarePixelBufferPtr CudaTexture::createSurface(uint32 face, uint32 mipmap, uint32, uint32, uint32)
{
    return std::make_shared<CudaHardwarePixelBuffer>(mBuffer.getPixelBox(face, mipmap), (HardwareBuffer::Usage)mUsage, this,
                                                     (mUsage & TU_RENDERTARGET) != 0 && face == 0 && mipmap == 0);
}

void CudaTexture::createInternalResourcesImpl(void)
{
    if ((mUsage & TU_RENDERTARGET) && (mUsage & ~TU_RENDERTARGET) == 0)
        mUsage |= HardwareBuffer::HBU_DYNAMIC;
    mFormat = TextureManager::getSingleton().getNativeFormat(mTextureType, mFormat, mUsage);
    mNumMipmaps = mNumRequestedMipmaps = 0;   // like Tiny: no mip chain (OgreTinyTexture.cpp:45)
    mBuffer.create(mFormat, mWidth, mHeight, mDepth, getNumFaces(), mNumMipmaps);
    if (mUsage & TU_RENDERTARGET)
        OGRECU_THROW_IF(ogrecu_rt_create_placed(mCtx, mWidth, mHeight, CudaRenderSystem::getInstance()->placementForRenderTexture(), &mTarget));
    mDirty = true;
    createSurfaceList();
}

void CudaTexture::freeInternalResourcesImpl(void)
{
    CudaRenderSystem::flushPendingDraws();   // recorded draws may sample this texture
    mSurfaceList.clear();                    // destroys the RenderTextures before their ogrecu_rt
    if (mDevice) { ogrecu_tex_destroy(mDevice); mDevice = 0; }
    if (mTarget) { CudaRenderSystem::forgetTarget(mTarget); ogrecu_rt_destroy(mTarget); mTarget = 0; }
}

ogrecu_tex* CudaTexture::getDeviceTexture()
{
    if (!mDevice)
    {
        OGRECU_THROW_IF(ogrecu_tex_create(mCtx, mWidth, mHeight, &mDevice));
        mDirty = true;
    }
    if (mTarget)
    {
        if (mTargetDirty || mDirty)
        {
            CudaRenderSystem::flushPendingDraws();   // draws recorded earlier sample the previous contents
            OGRECU_THROW_IF(ogrecu_tex_copy_from_rt(mDevice, mTarget));   // RGB8 plane -> RGBA8, alpha 255, on device
            mTargetDirty = mDirty = false;
        }
    }
    else if (mDirty)
    {
        CudaRenderSystem::flushPendingDraws();
        OGRECU_THROW_IF(ogrecu_tex_upload(mDevice, mBuffer.getData(), (size_t)mWidth * 4));
        mDirty = false;
    }
    return mDevice;
}

/* ------------------------------------------------------------------ pixel buffer ---- */

CudaHardwarePixelBuffer::CudaHardwarePixelBuffer(const PixelBox& data, Usage usage, CudaTexture* parent, bool renderTarget)
    : HardwarePixelBuffer(data.getWidth(), data.getHeight(), data.getDepth(), data.format, usage, false), mBuffer(data),
      mParent(parent), mLockedForWrite(false)
{
    if (renderTarget)
    {
        // one RenderTexture per slice, attached to the render system (priority group of RTTs)
        mSliceTRT.reserve(mDepth);
        for (uint32 zoffset = 0; zoffset < mDepth; ++zoffset)
        {
            String name = "rtt/" + StringConverter::toString((size_t)this) + "/" + parent->getName();
            RenderTexture* trt = new CudaRenderTexture(name, this, zoffset);
            mSliceTRT.push_back(trt);
            Root::getSingleton().getRenderSystem()->attachRenderTarget(*trt);
        }
    }
}

PixelBox CudaHardwarePixelBuffer::lockImpl(const Box& lockBox, LockOptions options)
{
    mLockedForWrite = options != HBL_READ_ONLY;
    if (mParent->getTarget() && options != HBL_DISCARD)
    {
        // bring the rendered image into the host copy first
        PixelBox whole = mBuffer;
        CudaRenderSystem::readColour(mParent->getTarget(), Box(0, 0, mWidth, mHeight), whole);
    }
    return mBuffer.getSubVolume(lockBox);
}

void CudaHardwarePixelBuffer::unlockImpl(void)
{
    if (mLockedForWrite) mParent->markDirty();
}

void CudaHardwarePixelBuffer::blitFromMemory(const PixelBox& src, const Box& dstBox)
{
    if (!mBuffer.contains(dstBox))
        OGRE_EXCEPT(Exception::ERR_INVALIDPARAMS, "Destination box out of range");
    // host-side conversion / scaling exactly as TinyHardwarePixelBuffer::blitFromMemory
    // (OgreTinyHardwarePixelBuffer.cpp:14-34); the device copy is refreshed lazily
    PixelBox scaled = mBuffer.getSubVolume(dstBox);
    if (src.getSize() != dstBox.getSize())
        Image::scale(src, scaled, Image::FILTER_BILINEAR);
    else
        PixelUtil::bulkPixelConversion(src, scaled);
    mParent->markDirty();
}

void CudaHardwarePixelBuffer::blitToMemory(const Box& srcBox, const PixelBox& dst)
{
    if (!mBuffer.contains(srcBox))
        OGRE_EXCEPT(Exception::ERR_INVALIDPARAMS, "source box out of range");
    if (mParent->getTarget())
    {
        // render-to-texture readback (RenderTexture::copyContentsToMemory, OgreMain/src/OgreRenderTexture.cpp:50-56)
        if (srcBox.getSize() == dst.getSize())
        {
            CudaRenderSystem::readColour(mParent->getTarget(), srcBox, dst);
            return;
        }
        if (srcBox.getDepth() == 1 && dst.getDepth() == 1)
        {
            // scaled read-back: the resampling runs on the device, only the destination's pixels cross the bus
            CudaRenderSystem::readColourScaled(mParent->getTarget(), srcBox, dst);
            return;
        }
        PixelBox whole = mBuffer;
        CudaRenderSystem::readColour(mParent->getTarget(), Box(0, 0, mWidth, mHeight), whole);
    }
    if (srcBox.getSize() != dst.getSize())
        Image::scale(mBuffer.getSubVolume(srcBox), dst, Image::FILTER_BILINEAR);
    else
        PixelUtil::bulkPixelConversion(mBuffer.getSubVolume(srcBox), dst);
}

/* ---------------------------------------------------------------- render texture ---- */

CudaRenderTexture::CudaRenderTexture(const String& name, CudaHardwarePixelBuffer* buffer, uint32 zoffset)
    : RenderTexture(buffer, zoffset)
{
    mName = name;
}

ogrecu_rt* CudaRenderTexture::getTarget() const { return getTexture()->getTarget(); }
CudaTexture* CudaRenderTexture::getTexture() const { return static_cast<CudaHardwarePixelBuffer*>(mBuffer)->getParent(); }

void CudaRenderTexture::getCustomAttribute(const String& name, void* pData)
{
    if (name == "OGRECU_RT")
    {
        *static_cast<ogrecu_rt**>(pData) = getTarget();
        return;
    }
    RenderTexture::getCustomAttribute(name, pData);
}

/* --------------------------------------------------------------- texture manager ---- */

CudaTextureManager::CudaTextureManager(ogrecu_ctx* ctx) : mCtx(ctx)
{
    ResourceGroupManager::getSingleton()._registerResourceManager(mResourceType, this);
}

CudaTextureManager::~CudaTextureManager() { ResourceGroupManager::getSingleton()._unregisterResourceManager(mResourceType); }

Resource* CudaTextureManager::createImpl(const String& name, ResourceHandle handle, const String& group, bool isManual,
                                         ManualResourceLoader* loader, const NameValuePairList*)
{
    return new CudaTexture(this, name, handle, group, isManual, loader, mCtx);
}

PixelFormat CudaTextureManager::getNativeFormat(TextureType, PixelFormat, int)
{
    return PF_BYTE_RGBA;   // only byte formats, like Tiny (OgreTinyTextureManager.cpp:33-37)
}
} // namespace Ogre
