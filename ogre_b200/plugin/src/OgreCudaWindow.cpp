#include "OgreCudaWindow.h"
#include "OgrePixelFormat.h"
#include "OgreCudaRenderSystem.h"

namespace Ogre
{
CudaWindow::CudaWindow(ogrecu_ctx* ctx) : mCtx(ctx), mTarget(0)
{
    mIsFullScreen = false;
    mActive = true;
}

CudaWindow::~CudaWindow() { destroy(); }

void CudaWindow::create(const String& name, uint width, uint height, bool, const NameValuePairList*)
{
    mName = name;
    resize(width, height);
}

void CudaWindow::destroy(void)
{
    if (mTarget)
    {
        ogrecu_rt_destroy(mTarget);
        mTarget = 0;
    }
}

void CudaWindow::resize(uint width, uint height)
{
    destroy();
    OGRECU_THROW_IF(ogrecu_rt_create(mCtx, width, height, &mTarget));   // zero-filled like Image::create
    RenderWindow::resize(width, height);
}

void CudaWindow::swapBuffers()
{
    // offscreen only: make the frame's work start now (asynchronous); readback synchronises
    if (mTarget) OGRECU_THROW_IF(ogrecu_flush(mTarget));
}

void CudaWindow::copyContentsToMemory(const Box& src, const PixelBox& dst, FrameBuffer)
{
    if (mClosed) return;
    // same validation as TinyWindow::copyContentsToMemory (OgreTinyWindow.cpp:73-77)
    if (src.right > mWidth || src.bottom > mHeight || src.front != 0 || src.back != 1 || dst.getWidth() != src.getWidth() ||
        dst.getHeight() != src.getHeight() || dst.getDepth() != 1)
    {
        OGRE_EXCEPT(Exception::ERR_INVALIDPARAMS, "Invalid box");
    }
    CudaRenderSystem::readColour(mTarget, src, dst);
}

void CudaWindow::getCustomAttribute(const String& name, void* pData)
{
    if (name == "OGRECU_RT")
    {
        *static_cast<ogrecu_rt**>(pData) = mTarget;
        return;
    }
    RenderWindow::getCustomAttribute(name, pData);
}
} // namespace Ogre
