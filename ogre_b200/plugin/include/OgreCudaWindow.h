// Mirrors RenderSystems/Tiny/include/OgreTinyWindow.h: an offscreen window whose colour (RGB8,
// tightly packed like Tiny's PF_BYTE_RGB Image) and depth (f32) planes live in HBM.
#ifndef __OgreCudaWindow_H__
#define __OgreCudaWindow_H__

#include "OgreCudaPrerequisites.h"
#include "OgreRenderWindow.h"

namespace Ogre
{
class CudaWindow : public RenderWindow
{
public:
    explicit CudaWindow(ogrecu_ctx* ctx);
    ~CudaWindow();

    void create(const String& name, unsigned int width, unsigned int height, bool fullScreen,
                const NameValuePairList* miscParams) override;
    void destroy(void) override;
    void resize(unsigned int width, unsigned int height) override;
    void copyContentsToMemory(const Box& src, const PixelBox& dst, FrameBuffer buffer) override;
    bool requiresTextureFlipping() const override { return true; }
    void swapBuffers() override;
    /// "OGRECU_RT" -> ogrecu_rt* (depth readback for tests goes through the C ABI)
    void getCustomAttribute(const String& name, void* pData) override;

    ogrecu_rt* getTarget() const { return mTarget; }

protected:
    ogrecu_ctx* mCtx;
    ogrecu_rt* mTarget;
};
} // namespace Ogre
#endif
