#include "OgreCudaHostTiming.h"
#include "OgreCudaRenderSystem.h"

#include "OgreCudaDepthBuffer.h"
#include "OgreCudaHardwareBufferManager.h"
#include "OgreCudaTexture.h"
#include "OgreCudaWindow.h"
#include "OgreGpuProgramManager.h"
#include "OgreLogManager.h"
#include "OgreRoot.h"
#include "OgreStringConverter.h"
#include "OgreViewport.h"

namespace Ogre
{
CudaRenderSystem* CudaRenderSystem::msInstance = 0;

CudaRenderSystem::CudaRenderSystem()
    : mCtx(0), mDeviceIndex(0), mActiveTarget(0), mActiveTargetTexture(0), mDoLighting(false), mTexture(0), mDepthTest(true),
      mDepthWrite(true), mBlendAdd(false), mBeyondTiny(false), mDepthFunc(CMPF_LESS_EQUAL), mAlphaFunc(CMPF_ALWAYS_PASS), mAlphaRef(0),
      mScissorOn(false), mHardwareBufferManager(0), mInitialised(false)
{
    LogManager::getSingleton().logMessage(getName() + " created.");
    msInstance = this;
    if (const char* dev = getenv("OGRECU_DEVICE")) mDeviceIndex = atoi(dev);
    initConfigOptions();

    // the same five auto-constants as TinyRenderSystem (OgreTinyRenderSystem.cpp:35-42): OgreMain
    // then produces bit-identical uniform values for both back ends
    GpuLogicalBufferStructPtr logicalBufferStruct(new GpuLogicalBufferStruct());
    mFixedFunctionParams.reset(new GpuProgramParameters);
    mFixedFunctionParams->_setLogicalIndexes(logicalBufferStruct);
    mFixedFunctionParams->setAutoConstant(0, GpuProgramParameters::ACT_WORLDVIEWPROJ_MATRIX);
    mFixedFunctionParams->setAutoConstant(4, GpuProgramParameters::ACT_TEXTURE_MATRIX);
    mFixedFunctionParams->setAutoConstant(8, GpuProgramParameters::ACT_DERIVED_AMBIENT_LIGHT_COLOUR);
    mFixedFunctionParams->setAutoConstant(9, GpuProgramParameters::ACT_LIGHT_POSITION);
    mFixedFunctionParams->setAutoConstant(10, GpuProgramParameters::ACT_INVERSE_TRANSPOSE_WORLDVIEW_MATRIX);

    mActiveRenderTarget = 0;
}

CudaRenderSystem::~CudaRenderSystem()
{
    shutdown();
    if (mCtx)
    {
        ogrecu_shutdown(mCtx);
        mCtx = 0;
    }
    if (msInstance == this) msInstance = 0;
}

ogrecu_ctx* CudaRenderSystem::getContext()
{
    if (!mCtx)
    {
        String where;
        if (mDevices.size() > 1)
        {
            OGRECU_THROW_IF(ogrecu_init_devices((int)mDevices.size(), mDevices.data(), &mCtx));
            where = "device group";
            for (int d : mDevices) where += " " + StringConverter::toString(d);
            where += mSplitFrame ? " (windows split over the group)" : " (windows on the first device)";
        }
        else
        {
            if (mDevices.size() == 1) mDeviceIndex = mDevices[0];
            OGRECU_THROW_IF(ogrecu_init(mDeviceIndex, &mCtx));
            where = "device " + StringConverter::toString(mDeviceIndex);
        }
        OGRECU_THROW_IF(ogrecu_set_config_option(mCtx, "optutil_batch", mBatchOptimisedUtil ? 1 : 0));
        LogManager::getSingleton().logMessage("CUDA RenderSystem: libogrecu context on " + where +
                                              (mBatchOptimisedUtil ? ", OptimisedUtil calls batched per frame" : ""));
    }
    return mCtx;
}

void CudaRenderSystem::initConfigOptions()
{
    RenderSystem::initConfigOptions();
    ConfigOption dev;
    dev.name = "Device";
    dev.currentValue = StringConverter::toString(mDeviceIndex);
    dev.immutable = false;
    int n = 0;
    ogrecu_device_count(&n);
    for (int i = 0; i < n; ++i) dev.possibleValues.push_back(StringConverter::toString(i));
    mOptions[dev.name] = dev;
    ConfigOption devs;
    devs.name = "Devices";           // several GPUs driven by this process (libogrecu device group); empty = only "Device"
    devs.currentValue = "";
    devs.immutable = false;
    devs.possibleValues.push_back("");
    if (n > 1)
    {
        String all;
        for (int i = 0; i < n; ++i)
        {
            all += (i ? "," : "") + StringConverter::toString(i);
            if (i) devs.possibleValues.push_back(all);
        }
        devs.possibleValues.push_back("All");
    }
    mOptions[devs.name] = devs;
    ConfigOption split;
    split.name = "Split Frame";      // with "Devices": windows are rasterised by all devices in interleaved bands of tile rows
    split.currentValue = mSplitFrame ? "Yes" : "No";
    split.possibleValues = {"No", "Yes"};
    split.immutable = false;
    mOptions[split.name] = split;
    ConfigOption beyond;
    beyond.name = "Beyond Tiny";      // honour depth function, culling mode, scissor, alpha rejection and vertex colours
    beyond.currentValue = mBeyondTiny ? "Yes" : "No";
    beyond.possibleValues = {"No", "Yes"};
    beyond.immutable = false;
    mOptions[beyond.name] = beyond;
    ConfigOption batch;
    batch.name = "OptimisedUtil Batching";   // queue softwareVertexSkinning calls, one batched launch per frame (ogrecu "optutil_batch")
    batch.currentValue = mBatchOptimisedUtil ? "Yes" : "No";
    batch.possibleValues = {"No", "Yes"};
    batch.immutable = false;
    mOptions[batch.name] = batch;
}

void CudaRenderSystem::setConfigOption(const String& name, const String& value)
{
    auto it = mOptions.find(name);
    if (it != mOptions.end()) it->second.currentValue = value;
    if (name == "Device")
    {
        OgreAssert(!mCtx, "the device cannot change once the context exists");
        mDeviceIndex = StringConverter::parseInt(value);
    }
    if (name == "Devices")
    {
        OgreAssert(!mCtx, "the devices cannot change once the context exists");
        mDevices.clear();
        if (value == "All" || value == "all")
        {
            int n = 0;
            ogrecu_device_count(&n);
            for (int i = 0; i < n; ++i) mDevices.push_back(i);
        }
        else
            for (const String& tok : StringUtil::split(value, ", ")) mDevices.push_back(StringConverter::parseInt(tok));
    }
    if (name == "Split Frame") mSplitFrame = StringConverter::parseBool(value);
    if (name == "Beyond Tiny") mBeyondTiny = StringConverter::parseBool(value);
    if (name == "OptimisedUtil Batching")
    {
        mBatchOptimisedUtil = StringConverter::parseBool(value);
        if (mCtx) OGRECU_THROW_IF(ogrecu_set_config_option(mCtx, "optutil_batch", mBatchOptimisedUtil ? 1 : 0));
    }
}

int CudaRenderSystem::placementForWindow() const { return (mDevices.size() > 1 && mSplitFrame) ? OGRECU_PLACE_SPLIT : 0; }

int CudaRenderSystem::placementForRenderTexture()
{
    if (mDevices.size() <= 1) return 0;
    const int k = mNextPlacement;
    mNextPlacement = (mNextPlacement + 1) % (int)mDevices.size();
    return k;
}

const String& CudaRenderSystem::getName(void) const
{
    static String strName("CUDA Rendering Subsystem");
    return strName;
}

RenderSystemCapabilities* CudaRenderSystem::createRenderSystemCapabilities() const
{
    // mirrors OgreTinyRenderSystem.cpp:94-131 so OgreMain takes the same code paths (FFP, no RTSS)
    RenderSystemCapabilities* rsc = OGRE_NEW RenderSystemCapabilities();
    rsc->setDriverVersion(mDriverVersion);
    rsc->setRenderSystemName(getName());
    rsc->setDeviceName("NVIDIA B200 (sm_100a) via libogrecu");
    rsc->setNumTextureUnits(1);
    rsc->setCapability(RSC_FIXED_FUNCTION);
    rsc->setCapability(RSC_MAPBUFFER);
    rsc->setCapability(RSC_NON_POWER_OF_2_TEXTURES);
    rsc->setCapability(RSC_USER_CLIP_PLANES);
    rsc->setCapability(RSC_TEXTURE_1D);
    rsc->setCapability(RSC_TEXTURE_3D);
    rsc->setCapability(RSC_TEXTURE_2D_ARRAY);
    rsc->setCapability(RSC_32BIT_INDEX);
    rsc->setCapability(RSC_DEPTH_CLAMP);
    rsc->setCapability(RSC_VERTEX_BUFFER_INSTANCE_DATA);
    rsc->setCapability(RSC_TEXTURE_FLOAT);
    rsc->setCapability(RSC_VERTEX_TEXTURE_FETCH);
    return rsc;
}

void CudaRenderSystem::shutdown(void)
{
    if (mCtx) flushPendingDraws();
    RenderSystem::shutdown();   // destroys render targets (windows, render textures)
    mDirtyTargets.clear();
    mActiveTarget = 0;
    mActiveTargetTexture = 0;
    mTexture = 0;

    OGRE_DELETE mHardwareBufferManager;
    mHardwareBufferManager = 0;
    OGRE_DELETE mTextureManager;
    mTextureManager = 0;
    mInitialised = false;
}

RenderWindow* CudaRenderSystem::_createRenderWindow(const String& name, unsigned int width, unsigned int height,
                                                    bool fullScreen, const NameValuePairList* miscParams)
{
    RenderSystem::_createRenderWindow(name, width, height, fullScreen, miscParams);

    RenderWindow* win = new CudaWindow(getContext());
    win->create(name, width, height, fullScreen, miscParams);
    attachRenderTarget(*win);

    if (!mInitialised)
    {
        LogManager::getSingleton().logMessage("***  CUDA (B200) software RenderSystem started  ***");
        mRealCapabilities = createRenderSystemCapabilities();
        if (!mUseCustomCapabilities) mCurrentCapabilities = mRealCapabilities;
        fireEvent("RenderSystemCapabilitiesCreated");
        mHardwareBufferManager = new CudaHardwareBufferManager(getContext());
        mTextureManager = new CudaTextureManager(getContext());
        mInitialised = true;
    }

    if (win->getDepthBufferPool() != RBP_NONE)
    {
        auto* depthBuffer = new CudaDepthBuffer(win->getDepthBufferPool(), win->getWidth(), win->getHeight(), win->getFSAA(), true);
        mDepthBufferPool[depthBuffer->getPoolId()].push_back(depthBuffer);
        win->attachDepthBuffer(depthBuffer);
    }
    return win;
}

DepthBuffer* CudaRenderSystem::_createDepthBufferFor(RenderTarget* rt)
{
    return new CudaDepthBuffer(rt->getDepthBufferPool(), rt->getWidth(), rt->getHeight(), rt->getFSAA(), false);
}

MultiRenderTarget* CudaRenderSystem::createMultiRenderTarget(const String&) { return NULL; }

void CudaRenderSystem::_setTexture(size_t stage, bool enabled, const TexturePtr& texPtr)
{
    if (stage > 0) return;   // one texture unit (OgreTinyRenderSystem.cpp:208-220)
    mTexture = (!enabled || !texPtr) ? NULL : static_cast<CudaTexture*>(texPtr.get());
}

void CudaRenderSystem::_setViewport(Viewport* vp)
{
    if (!vp)
    {
        mActiveViewport = NULL;
        _setRenderTarget(NULL);
    }
    else if (vp != mActiveViewport || vp->_isUpdated())
    {
        RenderTarget* target = vp->getTarget();
        _setRenderTarget(target);
        mActiveViewport = vp;

        // identical arithmetic to TinyRenderSystem::_setViewport (OgreTinyRenderSystem.cpp:230-262)
        Rect vpRect = vp->getActualDimensions();
        if (!target->requiresTextureFlipping())
        {
            std::swap(vpRect.top, vpRect.bottom);
            vpRect.top = target->getHeight() - vpRect.top;
            vpRect.bottom = target->getHeight() - vpRect.bottom;
        }
        mVP.makeTransform({vpRect.left + vpRect.width() / 2.f, vpRect.top + vpRect.height() / 2.f, 0.5},
                          {vpRect.width() / 2.f, vpRect.height() / 2.f, 0.5}, Ogre::Quaternion::IDENTITY);
        vp->_clearUpdatedFlag();
    }
}

void CudaRenderSystem::noteRecorded(ogrecu_rt* rt)
{
    if (std::find(mDirtyTargets.begin(), mDirtyTargets.end(), rt) == mDirtyTargets.end()) mDirtyTargets.push_back(rt);
}

void CudaRenderSystem::flushPendingDraws()
{
    CudaRenderSystem* rs = msInstance;
    if (!rs) return;
    OGRECU_TIME(HT_FLUSH);
    if (rs->mCtx) OGRECU_THROW_IF(ogrecu_optutil_flush(rs->mCtx));   // queued OptimisedUtil jobs: callers are about to use raw device data
    if (rs->mDirtyTargets.empty()) return;
    std::vector<ogrecu_rt*> targets;
    targets.swap(rs->mDirtyTargets);
    for (ogrecu_rt* rt : targets) OGRECU_THROW_IF(ogrecu_flush(rt));
}

void CudaRenderSystem::forgetTarget(ogrecu_rt* rt, ogrecu_rt* replacement)
{
    CudaRenderSystem* rs = msInstance;
    if (!rs || !rt) return;
    auto it = std::find(rs->mDirtyTargets.begin(), rs->mDirtyTargets.end(), rt);
    if (it != rs->mDirtyTargets.end())
    {
        rs->mDirtyTargets.erase(it);
        ogrecu_flush(rt);   // recorded clears / draws still run (their results die with the target); never throws from a destructor path
    }
    if (rs->mActiveTarget == rt)
    {
        rs->mActiveTarget = replacement;
        if (!replacement) rs->mActiveTargetTexture = 0;
    }
}

void CudaRenderSystem::_endFrame(void)
{
    OGRECU_TIME(HT_END_FRAME);
    // launch the frame's recorded work now; it runs asynchronously while OgreMain goes on
    // (Root::_fireFrameRenderingQueued, OgreRoot.cpp:1156, is the designed overlap window)
    flushPendingDraws();
}

void CudaRenderSystem::_setDepthBufferParams(bool depthTest, bool depthWrite, CompareFunction depthFunction)
{
    mDepthTest = depthTest;
    mDepthWrite = depthWrite;
    mDepthFunc = depthFunction;   // Tiny drops it (OgreTinyRenderSystem.cpp:270-274); sent only in "Beyond Tiny" mode
}

void CudaRenderSystem::setColourBlendState(const ColourBlendState& state) { mBlendAdd = state.destFactor == SBF_ONE; }

void CudaRenderSystem::applyFixedFunctionParams(const GpuProgramParametersPtr& params, uint16 mask)
{
    // same slots, same conversions as TinyRenderSystem::applyFixedFunctionParams (OgreTinyRenderSystem.cpp:49-82)
    for (const auto& ac : params->getAutoConstants())
    {
        if (!(ac.variability & mask)) continue;
        const float* ptr = params->getFloatPointer(ac.physicalIndex);
        switch (ac.paramType)
        {
        case GpuProgramParameters::ACT_WORLDVIEWPROJ_MATRIX: mUniformMVP = Matrix4(ptr); break;
        case GpuProgramParameters::ACT_TEXTURE_MATRIX: mUniformTex = Matrix4(ptr); break;
        case GpuProgramParameters::ACT_DERIVED_AMBIENT_LIGHT_COLOUR: memcpy(mUniformAmbient.ptr(), ptr, sizeof(float) * 4); break;
        case GpuProgramParameters::ACT_LIGHT_POSITION:
            mUniformLightDir = ptr[3] ? Vector3(ptr).normalisedCopy() : -Vector3(ptr);
            break;
        case GpuProgramParameters::ACT_INVERSE_TRANSPOSE_WORLDVIEW_MATRIX: mUniformMVIT = Matrix4(ptr); break;
        default: break;
        }
    }
}

namespace
{
struct Stream
{
    const void* ptr;
    uint32 stride;
    ogrecu_buf* buf;
};

// device address of the first element of a vertex element, the way TinyRenderSystem's getData() computes the
// host address (OgreTinyRenderSystem.cpp:335-347): base + elemOffset + vertexStart*stride
Stream deviceStream(const RenderOperation& op, const VertexElement* element)
{
    Stream s = {0, 0, 0};
    if (!element) return s;
    s.stride = (uint32)op.vertexData->vertexDeclaration->getVertexSize(element->getSource());
    auto buf = op.vertexData->vertexBufferBinding->getBuffer(element->getSource());
    auto impl = dynamic_cast<CudaHardwareBuffer*>(buf->_getImpl<HardwareBuffer>());
    OgreAssert(impl, "vertex buffer was not created by CudaHardwareBufferManager");
    s.ptr = impl->getDevicePtr(element->getOffset() + op.vertexData->vertexStart * s.stride);
    s.buf = impl->getHandle();
    return s;
}
} // namespace

void CudaRenderSystem::_render(const RenderOperation& op)
{
    OGRECU_TIME(HT_RENDER);
    RenderSystem::_render(op);   // face / batch / vertex statistics (OgreMain/src/OgreRenderSystem.cpp:622)

    bool isStrip = op.operationType == RenderOperation::OT_TRIANGLE_STRIP;
    if (!isStrip && op.operationType != RenderOperation::OT_TRIANGLE_LIST)   // only triangle list / strip supported
        return;
    if (!mActiveTarget) return;

    // one walk over the declaration (findElementBySemantic walks it once per semantic; this runs once per sub-entity and frame)
    const VertexElement *ePos = 0, *eUv = 0, *eNrm = 0, *eCol = 0;
    for (const VertexElement& el : op.vertexData->vertexDeclaration->getElements())
    {
        if (el.getIndex() != 0) continue;                     // findElementBySemantic(sem, 0): the first element of index 0
        switch (el.getSemantic())
        {
        case VES_POSITION: if (!ePos) ePos = &el; break;
        case VES_TEXTURE_COORDINATES: if (!eUv) eUv = &el; break;
        case VES_NORMAL: if (!eNrm) eNrm = &el; break;
        case VES_DIFFUSE: if (!eCol) eCol = &el; break;
        default: break;
        }
    }
    Stream pos = deviceStream(op, ePos);
    OgreAssert(pos.ptr, "VES_POSITION required");
    Stream uv = deviceStream(op, eUv);
    Stream nrm = deviceStream(op, eNrm);
    const ogrecu_buf* refs[5] = {pos.buf, uv.buf, nrm.buf, 0, 0};

    mDoLighting &= bool(nrm.ptr);   // sticky, as in Tiny (OgreTinyRenderSystem.cpp:369)

    ogrecu_draw_desc d;
    memset(&d, 0, sizeof d);
    d.positions = pos.ptr; d.pos_stride = pos.stride;
    d.uvs = uv.ptr; d.uv_stride = uv.stride;
    d.normals = nrm.ptr; d.normal_stride = nrm.stride;
    d.index_type = OGRECU_INDEX_NONE;
    d.draw_count = (uint32_t)op.vertexData->vertexCount;
    d.vertex_count = (uint32_t)op.vertexData->vertexCount;   // indices address [vertexStart, vertexStart + vertexCount)
    if (op.useIndexes)
    {
        auto& ib = op.indexData->indexBuffer;
        auto impl = dynamic_cast<CudaHardwareBuffer*>(ib->_getImpl<HardwareBuffer>());
        OgreAssert(impl, "index buffer was not created by CudaHardwareBufferManager");
        d.index_type = ib->getIndexSize() == 2 ? OGRECU_INDEX_U16 : OGRECU_INDEX_U32;
        d.indices = impl->getDevicePtr(op.indexData->indexStart * ib->getIndexSize());
        d.draw_count = (uint32_t)op.indexData->indexCount;
        refs[3] = impl->getHandle();
    }
    d.prim_type = isStrip ? OGRECU_PRIM_TRIANGLE_STRIP : OGRECU_PRIM_TRIANGLE_LIST;
    d.flags = (mDepthTest ? OGRECU_DRAW_DEPTH_TEST : 0) | (mDepthWrite ? OGRECU_DRAW_DEPTH_WRITE : 0) |
              (mBlendAdd ? OGRECU_DRAW_BLEND_ADD : 0) | (mDoLighting ? OGRECU_DRAW_LIGHTING : 0);
    d.texture = mTexture ? mTexture->getDeviceTexture() : NULL;
    memcpy(d.mvp, mUniformMVP[0], 64);
    memcpy(d.tex_matrix, mUniformTex[0], 64);
    memcpy(d.mvit, mUniformMVIT[0], 64);
    memcpy(d.viewport, mVP[0], 64);
    memcpy(d.light_dir, mUniformLightDir.ptr(), 12);
    memcpy(d.ambient, mUniformAmbient.ptr(), 16);

    ogrecu_draw_state st;
    memset(&st, 0, sizeof st);
    if (mBeyondTiny)
    {
        // the Ogre enums ARE the C ABI's values (include/ogrecu.h)
        st.enable = OGRECU_EXT_DEPTH_FUNC | OGRECU_EXT_CULL_MODE;
        st.depth_func = mDepthFunc;
        CullingMode cull = mCullingMode;
        if (mInvertVertexWinding && cull != CULL_NONE) cull = cull == CULL_CLOCKWISE ? CULL_ANTICLOCKWISE : CULL_CLOCKWISE;
        st.cull_mode = cull;
        if (mAlphaFunc != CMPF_ALWAYS_PASS)
        {
            st.enable |= OGRECU_EXT_ALPHA_REJECT;
            st.alpha_func = mAlphaFunc; st.alpha_ref = mAlphaRef;
        }
        if (mScissorOn)
        {
            st.enable |= OGRECU_EXT_SCISSOR;
            st.scissor[0] = std::max<int32>(0, mScissor.left); st.scissor[1] = std::max<int32>(0, mScissor.top);
            st.scissor[2] = std::max<int32>(0, mScissor.right); st.scissor[3] = std::max<int32>(0, mScissor.bottom);
        }
        if (eCol && eCol->getType() == VET_UBYTE4_NORM)
        {
            Stream col = deviceStream(op, eCol);
            refs[4] = col.buf;
            st.enable |= OGRECU_EXT_VERTEX_COLOUR;
            st.colours = col.ptr; st.colour_stride = col.stride;
        }
    }

    do
    {
        OGRECU_THROW_IF(ogrecu_draw_refs(mActiveTarget, &d, mBeyondTiny ? &st : NULL, refs, 5));
    } while (updatePassIterationRenderState());
    noteRecorded(mActiveTarget);
    if (mActiveTargetTexture) mActiveTargetTexture->markTargetRendered();
}

void CudaRenderSystem::clearFrameBuffer(unsigned int buffers, const ColourValue& colour, float depth, unsigned short)
{
    if (!mActiveTarget) return;
    // whole target, like Tiny (OgreTinyRenderSystem.cpp:421-433); ordered with the recorded draws
    unsigned mask = ((buffers & FBT_COLOUR) ? OGRECU_CLEAR_COLOUR : 0) | ((buffers & FBT_DEPTH) ? OGRECU_CLEAR_DEPTH : 0);
    if (!mask) return;
    OGRECU_THROW_IF(ogrecu_rt_clear(mActiveTarget, mask, colour.ptr(), depth));
    noteRecorded(mActiveTarget);
    if (mActiveTargetTexture) mActiveTargetTexture->markTargetRendered();
}

void CudaRenderSystem::_setRenderTarget(RenderTarget* target)
{
    mActiveRenderTarget = target;
    mActiveTarget = 0;
    mActiveTargetTexture = 0;
    if (!target) return;

    if (auto win = dynamic_cast<CudaWindow*>(target))
        mActiveTarget = win->getTarget();
    else if (auto rtt = dynamic_cast<CudaRenderTexture*>(target))
    {
        mActiveTarget = rtt->getTarget();
        mActiveTargetTexture = rtt->getTexture();
    }

    // depth buffer pool bookkeeping exactly as Tiny (OgreTinyRenderSystem.cpp:447-456)
    auto* depthBuffer = target->getDepthBuffer();
    if (target->getDepthBufferPool() != RBP_NONE && !depthBuffer)
        setDepthBufferFor(target);
}

void CudaRenderSystem::readColour(ogrecu_rt* rt, const Box& src, const PixelBox& dst)
{
    flushPendingDraws();
    const size_t w = src.getWidth(), h = src.getHeight();
    if (!w || !h) return;
    if (dst.format == PF_BYTE_RGB && dst.isConsecutive())
    {
        OGRECU_THROW_IF(ogrecu_rt_read_colour(rt, src.left, src.top, src.right, src.bottom, dst.data, w * 3));
        return;
    }
    // 32-bit destinations with 8 bits per colour channel (PF_BYTE_RGBA / PF_BYTE_BGRA / PF_A8R8G8B8 / PF_X8R8G8B8 / ...): converted
    // on the device, the host receives the final bytes (the reference converts on the CPU, OgreTinyWindow.cpp:79)
    {
        int depths[4];
        unsigned char shifts[4];
        uint64 masks[4];
        PixelUtil::getBitDepths(dst.format, depths);
        PixelUtil::getBitShifts(dst.format, shifts);
        PixelUtil::getBitMasks(dst.format, masks);
        const unsigned flags = PixelUtil::getFlags(dst.format);
        if (PixelUtil::getNumElemBytes(dst.format) == 4 && (flags & PFF_NATIVEENDIAN) && !(flags & (PFF_FLOAT | PFF_INTEGER | PFF_COMPRESSED | PFF_LUMINANCE | PFF_DEPTH)) &&
            depths[0] == 8 && depths[1] == 8 && depths[2] == 8 && (depths[3] == 8 || depths[3] == 0) && dst.getDepth() == 1)
        {
            // alpha: unpackColour gives 1.0 for a source without alpha; the reference's converters also set the X byte of the
            // X8 formats (checked against the real PixelUtil in tests/test_plugin_gpu.py)
            const uint32 fill = ~(uint32)(masks[0] | masks[1] | masks[2]);
            uint8* top = dst.getTopLeftFrontPixelPtr();
            OGRECU_THROW_IF(ogrecu_rt_read_colour_packed32(rt, src.left, src.top, src.right, src.bottom, top, dst.rowPitch * 4, shifts[0], shifts[1],
                                                           shifts[2], fill));
            return;
        }
    }
    std::vector<uint8> tmp(w * h * 3);
    OGRECU_THROW_IF(ogrecu_rt_read_colour(rt, src.left, src.top, src.right, src.bottom, tmp.data(), w * 3));
    PixelBox box(w, h, 1, PF_BYTE_RGB, tmp.data());
    PixelUtil::bulkPixelConversion(box, dst);   // format conversion as TinyWindow::copyContentsToMemory (OgreTinyWindow.cpp:79)
}

// true + shifts/fill when `format` is a native-endian 32-bit format with 8 bits per colour channel (what pack32 can write)
static bool packed32Layout(PixelFormat format, unsigned char shifts[4], uint32& fill)
{
    int depths[4];
    uint64 masks[4];
    PixelUtil::getBitDepths(format, depths);
    PixelUtil::getBitShifts(format, shifts);
    PixelUtil::getBitMasks(format, masks);
    const unsigned flags = PixelUtil::getFlags(format);
    fill = ~(uint32)(masks[0] | masks[1] | masks[2]);
    return PixelUtil::getNumElemBytes(format) == 4 && (flags & PFF_NATIVEENDIAN) &&
           !(flags & (PFF_FLOAT | PFF_INTEGER | PFF_COMPRESSED | PFF_LUMINANCE | PFF_DEPTH)) && depths[0] == 8 && depths[1] == 8 &&
           depths[2] == 8 && (depths[3] == 8 || depths[3] == 0);
}

void CudaRenderSystem::readColourScaled(ogrecu_rt* rt, const Box& src, const PixelBox& dst)
{
    flushPendingDraws();
    const uint32 dw = dst.getWidth(), dh = dst.getHeight();
    if (!dw || !dh || !src.getWidth() || !src.getHeight()) return;
    uint8* top = dst.getTopLeftFrontPixelPtr();
    unsigned char shifts[4];
    uint32 fill;
    if (dst.format == PF_BYTE_RGB && dst.getDepth() == 1)
    {
        OGRECU_THROW_IF(ogrecu_rt_read_colour_scaled(rt, src.left, src.top, src.right, src.bottom, dw, dh, top, dst.rowPitch * 3, 3, 0, 0, 0, 0));
        return;
    }
    if (dst.getDepth() == 1 && packed32Layout(dst.format, shifts, fill))
    {
        OGRECU_THROW_IF(ogrecu_rt_read_colour_scaled(rt, src.left, src.top, src.right, src.bottom, dw, dh, top, dst.rowPitch * 4, 4, shifts[0],
                                                     shifts[1], shifts[2], fill));
        return;
    }
    // any other destination format: scaled on the device in the source format, converted by PixelUtil on the host — the
    // temp-buffer tail of Image::scale (OgreImage.cpp:436-457)
    std::vector<uint8> tmp((size_t)dw * dh * 3);
    OGRECU_THROW_IF(ogrecu_rt_read_colour_scaled(rt, src.left, src.top, src.right, src.bottom, dw, dh, tmp.data(), (size_t)dw * 3, 3, 0, 0, 0, 0));
    PixelBox box(dw, dh, 1, PF_BYTE_RGB, tmp.data());
    PixelUtil::bulkPixelConversion(box, dst);
}
} // namespace Ogre

#ifdef OGRECU_HOST_TIMING
namespace Ogre { uint64_t HostTimer::acc[HT_COUNT] = {}; uint64_t HostTimer::calls[HT_COUNT] = {}; }
// ticks and calls per bucket since the last read (buckets: render, skin, other OptimisedUtil, lock, unlock, flush, endFrame)
extern "C" __attribute__((visibility("default"))) void ogrecu_plugin_host_timing(uint64_t* acc, uint64_t* calls)
{
    for (int i = 0; i < Ogre::HT_COUNT; ++i) {
        acc[i] = Ogre::HostTimer::acc[i]; calls[i] = Ogre::HostTimer::calls[i];
        Ogre::HostTimer::acc[i] = 0; Ogre::HostTimer::calls[i] = 0;
    }
}
#endif
