// CudaRenderSystem — mirrors RenderSystems/Tiny/include/OgreTinyRenderSystem.h: same fixed-function
// contract (5 auto-constants, 1 texture unit, [-1,1] depth range, identity projection conversion,
// texture flipping on), same state honoured / ignored.  The difference: _render
// RECORDS a draw descriptor (state + uniform snapshot + device buffer addresses) on the target's
// ogrecu_rt; the binned tile rasteriser in libogrecu.so runs at _endFrame / readback / buffer lock.
#ifndef __CudaRenderSystem_H__
#define __CudaRenderSystem_H__

#include "OgreCudaPrerequisites.h"
#include "OgreRenderSystem.h"
#include "OgreRenderWindow.h"
#include "OgreMatrix4.h"

namespace Ogre
{
class HardwareBufferManager;

class CudaRenderSystem : public RenderSystem
{
    ogrecu_ctx* mCtx;
    int mDeviceIndex;
    Matrix4 mVP;   // viewport transform (TinyRenderSystem::mVP)

    ogrecu_rt* mActiveTarget;
    CudaTexture* mActiveTargetTexture;   // non-null when rendering into a RenderTexture

    // DefaultShader uniforms (RenderSystems/Tiny/include/OgreTinyRenderSystem.h:57-75)
    Matrix4 mUniformMVP, mUniformTex, mUniformMVIT;
    Vector3 mUniformLightDir;
    ColourValue mUniformAmbient;
    bool mDoLighting;
    CudaTexture* mTexture;   // unit 0 or NULL

    bool mDepthTest, mDepthWrite, mBlendAdd;

    // Opt-in "beyond Tiny" mode (config option "Beyond Tiny" = "Yes"): state that Tiny's setters drop
    // (OgreTinyRenderSystem.cpp:208-297) is kept and sent with every draw as an ogrecu_draw_state.  Off by default: the
    // render system is then Tiny's twin.
    bool mBeyondTiny;
    CompareFunction mDepthFunc, mAlphaFunc;
    unsigned char mAlphaRef;
    bool mScissorOn;
    bool mBatchOptimisedUtil = true;   // config option "OptimisedUtil Batching"
    // config options "Devices" (a device group driven by this process: "0,1,2,3", "All"; empty = the one "Device") and "Split Frame"
    // (windows are split over the group's devices in interleaved bands; render textures are placed round-robin either way)
    std::vector<int> mDevices;
    bool mSplitFrame = true;
    int mNextPlacement = 0;
    Rect mScissor;

    HardwareBufferManager* mHardwareBufferManager;
    bool mInitialised;
    std::vector<ogrecu_rt*> mDirtyTargets;   // targets with recorded, unflushed work

    static CudaRenderSystem* msInstance;
    void noteRecorded(ogrecu_rt* rt);

public:
    CudaRenderSystem();
    ~CudaRenderSystem();

    /// the device context; created on first use, throws when no CUDA device is usable (no fallback)
    ogrecu_ctx* getContext();
    static CudaRenderSystem* getInstance() { return msInstance; }
    /// run everything recorded on any target (called before buffers/textures that draws may reference change)
    static void flushPendingDraws();
    /// a target is about to be destroyed: run what is recorded on it, then drop every reference the render system keeps to it
    /// (the dirty list, the active target); `replacement` takes over as the active target when `rt` was active (window resize)
    static void forgetTarget(ogrecu_rt* rt, ogrecu_rt* replacement = 0);
    /// flush + read a colour box into an arbitrary-format PixelBox (readback path of windows and RTTs)
    static void readColour(ogrecu_rt* rt, const Box& src, const PixelBox& dst);
    /// read-back into a destination of another size: Image::scale(FILTER_BILINEAR) on the device (ogrecu_rt_read_colour_scaled)
    static void readColourScaled(ogrecu_rt* rt, const Box& src, const PixelBox& dst);

    Real getMinimumDepthInputValue(void) override { return -1.0f; }   // Range [-1.0f, 1.0f], as Tiny
    Real getMaximumDepthInputValue(void) override { return 1.0f; }
    void _convertProjectionMatrix(const Matrix4& matrix, Matrix4& dest, bool) override { dest = matrix; }

    void initConfigOptions() override;
    void setConfigOption(const String& name, const String& value) override;

    const String& getName(void) const override;
    RenderSystemCapabilities* createRenderSystemCapabilities() const override;
    void shutdown(void) override;

    RenderWindow* _createRenderWindow(const String& name, unsigned int width, unsigned int height, bool fullScreen,
                                      const NameValuePairList* miscParams = 0) override;
    DepthBuffer* _createDepthBufferFor(RenderTarget* renderTarget) override;
    MultiRenderTarget* createMultiRenderTarget(const String& name) override;

    void _setTexture(size_t unit, bool enabled, const TexturePtr& tex) override;
    void _setSampler(size_t unit, Sampler& sampler) override {}
    void setLightingEnabled(bool enabled) override { mDoLighting = enabled; }
    void _setViewport(Viewport* vp) override;
    void _endFrame(void) override;
    void _setCullingMode(CullingMode mode) override { mCullingMode = mode; }   // used in "Beyond Tiny" mode only
    bool getBeyondTiny() const { return mBeyondTiny; }
    void _setDepthBufferParams(bool depthTest = true, bool depthWrite = true,
                               CompareFunction depthFunction = CMPF_LESS_EQUAL) override;
    void _setDepthBias(float, float) override {}
    void setColourBlendState(const ColourBlendState& state) override;
    void _setPolygonMode(PolygonMode) override {}
    void setStencilState(const StencilState&) override {}
    void applyFixedFunctionParams(const GpuProgramParametersPtr& params, uint16 variabilityMask) override;
    void _render(const RenderOperation& op) override;
    void setScissorTest(bool enabled, const Rect& rect = Rect()) override { mScissorOn = enabled; mScissor = rect; }
    void clearFrameBuffer(unsigned int buffers, const ColourValue& colour = ColourValue::Black, float depth = 1.0f,
                          unsigned short stencil = 0) override;
    HardwareOcclusionQuery* createHardwareOcclusionQuery(void) override { return NULL; }
    void _setRenderTarget(RenderTarget* target) override;
    void bindGpuProgramParameters(GpuProgramType, const GpuProgramParametersPtr&, uint16) override {}
    void _setAlphaRejectSettings(CompareFunction func, unsigned char value, bool) override { mAlphaFunc = func; mAlphaRef = value; }
    // NVTX ranges / markers around whatever the application brackets (OgreRenderSystem.h:1104-1115)
    // where the next target goes on the device group: windows are split ("Split Frame") or live on the root, render textures
    // take the devices in turn (the batched offscreen views of BASELINE config 4 then run side by side)
    int placementForWindow() const;
    int placementForRenderTexture();
    void beginProfileEvent(const String& eventName) override { if (mCtx) ogrecu_profile_begin(mCtx, eventName.c_str()); }
    void endProfileEvent(void) override { if (mCtx) ogrecu_profile_end(mCtx); }
    void markProfileEvent(const String& event) override { if (mCtx) ogrecu_profile_mark(mCtx, event.c_str()); }
};
} // namespace Ogre
#endif
