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
        // a clear or draw may be recorded on the target and not flushed yet (manual Viewport::clear, an exception mid-frame):
        // the render system must not keep a pointer to it
        CudaRenderSystem::forgetTarget(mTarget);
        ogrecu_rt_destroy(mTarget);
        mTarget = 0;
    }
}

void CudaWindow::resize(uint width, uint height)
{
    ogrecu_rt* fresh = 0;
    OGRECU_THROW_IF(ogrecu_rt_create_placed(mCtx, width, height, CudaRenderSystem::getInstance()->placementForWindow(), &fresh));   // zero-filled like Image::create
    if (mTarget)
    {
        CudaRenderSystem::forgetTarget(mTarget, fresh);                // an active window keeps rendering, into the new target
        ogrecu_rt_destroy(mTarget);
    }
    mTarget = fresh;
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
