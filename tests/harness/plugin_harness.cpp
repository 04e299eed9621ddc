// plugin_harness.cpp — drives stock OgreMain with a render-system PLUGIN chosen at run time
// (RenderSystem_CUDA.so under test, or RenderSystem_Tiny.so as the oracle), exactly the way an OGRE
// application does: Root::loadPlugin -> setRenderSystem -> createRenderWindow -> renderOneFrame.
// TEST INFRASTRUCTURE (tests/ only).  Plain C entry points for ctypes; one Root per process.
#include <Ogre.h>
#include <OgreOptimisedUtil.h>
#include <OgreSubEntity.h>
#include <OgreSubMesh.h>
#include <OgreSkeletonInstance.h>
#include <OgreHardwarePixelBuffer.h>
#include <OgreRenderTexture.h>
#include "OgreCudaSkeletonPalettes.h"
#include "OgreCudaShadowVolume.h"
#include "OgreCudaPoseAnimation.h"
#include <OgreShadowCaster.h>
#include <OgreEdgeListBuilder.h>

#include <xmmintrin.h>
#include <chrono>
#include <cmath>
#include <csignal>
#include <execinfo.h>
#include <unistd.h>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace Ogre;
#define API extern "C" __attribute__((visibility("default")))

static OptimisedUtil* g_stock_util = OptimisedUtil::getImplementation();   // before any plugin is loaded

struct Host {
    std::unique_ptr<Root> root;
    RenderWindow* win = nullptr;
    SceneManager* sm = nullptr;
    Camera* cam = nullptr;
    Entity* ent = nullptr;
    std::string media;
    uint32_t w = 0, h = 0;
    TexturePtr rtt;
};
static Host* g = nullptr;
static std::string g_err;

API const char* ph_error(void) { return g_err.c_str(); }

#define TRY try {
#define CATCH(rc) } catch (std::exception& e) { g_err = e.what(); fprintf(stderr, "plugin_harness: %s\n", e.what()); return rc; }

static void crash_handler(int sig)
{
    // native backtrace of a crash inside the render system under test (the reference can crash too)
    void* frames[64];
    int n = backtrace(frames, 64);
    const char msg[] = "plugin_harness: fatal signal, native backtrace:\n";
    (void)!write(2, msg, sizeof msg - 1);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

API int ph_open(const char* plugin_so, const char* rs_name_part, const char* stbi_so, const char* media_dir, uint32_t w, uint32_t h)
{
    TRY
    if (g) { g_err = "already open"; return 1; }
    signal(SIGFPE, crash_handler);
    signal(SIGSEGV, crash_handler);
    g = new Host();
    LogManager* lm = new LogManager();
    lm->createLog("plugin_harness.log", true, getenv("PH_LOG") != nullptr, true);   // PH_LOG=1: OGRE log to stderr
    g->root.reset(new Root("", "", ""));
    g->root->loadPlugin(plugin_so);
    g->root->loadPlugin(stbi_so);
    RenderSystem* rs = nullptr;
    for (auto* r : g->root->getAvailableRenderers())
        if (r->getName().find(rs_name_part) != String::npos) rs = r;
    if (!rs) { g_err = "render system not found"; return 2; }
    g->root->setRenderSystem(rs);
    if (getenv("PH_BEYOND_TINY")) rs->setConfigOption("Beyond Tiny", "Yes");   // CUDA plugin only: opt-in state Tiny ignores
    if (const char* devs = getenv("PH_DEVICES")) rs->setConfigOption("Devices", devs);          // CUDA plugin only: a device group, e.g. "0,1" or "All"
    if (const char* split = getenv("PH_SPLIT_FRAME")) rs->setConfigOption("Split Frame", split);
    g->root->initialise(false);
    g->win = g->root->createRenderWindow("harness", w, h, false);
    g->w = w; g->h = h;
    g->media = media_dir;
    ResourceGroupManager& rgm = ResourceGroupManager::getSingleton();
    rgm.addResourceLocation(g->media + "/models", "FileSystem");
    rgm.addResourceLocation(g->media + "/scripts", "FileSystem");
    rgm.addResourceLocation(g->media + "/textures", "FileSystem");
    rgm.addResourceLocation(g->media + "/packs/Sinbad.zip", "Zip");
    rgm.initialiseAllResourceGroups();
    return 0;
    CATCH(3)
}

API const char* ph_render_system_name(void) { return g->root->getRenderSystem()->getName().c_str(); }
API int ph_optimised_util_replaced(void) { return OptimisedUtil::getImplementation() != g_stock_util; }

// the scene of SURVEY §8d C1 (same as oracle/ref_driver.cpp ref_scene_open)
API int ph_scene_open(const char* mesh, float cam_z)
{
    TRY
    if (g->sm) { g->win->removeAllViewports(); g->root->destroySceneManager(g->sm); g->sm = nullptr; g->ent = nullptr; }
    SceneManager* sm = g->root->createSceneManager();
    g->sm = sm;
    sm->setAmbientLight(ColourValue(0.3f, 0.3f, 0.3f));
    Light* light = sm->createLight("l");
    SceneNode* ln = sm->getRootSceneNode()->createChildSceneNode();
    ln->setPosition(20, 80, 50);
    ln->attachObject(light);
    Camera* cam = sm->createCamera("c");
    cam->setNearClipDistance(5);
    cam->setAutoAspectRatio(true);
    SceneNode* cn = sm->getRootSceneNode()->createChildSceneNode();
    cn->setPosition(0, 0, cam_z);
    cn->lookAt(Vector3::ZERO, Node::TS_WORLD);
    cn->attachObject(cam);
    g->cam = cam;
    Viewport* vp = g->win->addViewport(cam);
    vp->setBackgroundColour(ColourValue(0.1f, 0.2f, 0.3f));
    Entity* e = sm->createEntity(mesh);
    sm->getRootSceneNode()->createChildSceneNode()->attachObject(e);
    g->ent = e;
    return 0;
    CATCH(1)
}

// Pass state of every sub-entity's material (shared materials: test scene only): culling mode (Ogre::CullingMode, 0 = keep),
// depth function (Ogre::CompareFunction, -1 = keep), alpha rejection (function -1 = keep).  Tiny ignores all of it.
API int ph_set_pass_state(int cull_mode, int depth_func, int alpha_func, int alpha_ref)
{
    TRY
    if (!g->ent) { g_err = "no entity"; return 2; }
    for (SubEntity* se : g->ent->getSubEntities()) {
        Pass* p = se->getMaterial()->getTechnique(0)->getPass(0);
        if (cull_mode) p->setCullingMode((CullingMode)cull_mode);
        if (depth_func >= 0) p->setDepthFunction((CompareFunction)depth_func);
        if (alpha_func >= 0) p->setAlphaRejectSettings((CompareFunction)alpha_func, (unsigned char)alpha_ref);
    }
    return 0;
    CATCH(1)
}

API int ph_render(void)
{
    TRY
    g->root->renderOneFrame();
    return 0;
    CATCH(1)
}

API size_t ph_triangles(void) { return g->win->getStatistics().triangleCount; }
API size_t ph_batches(void) { return g->win->getStatistics().batchCount; }

// public readback path: RenderTarget::copyContentsToMemory into PF_BYTE_RGB (fmt 0) or PF_BYTE_RGBA (fmt 1)
API int ph_read_colour(uint8_t* dst, int fmt)
{
    TRY
    PixelBox box(g->w, g->h, 1, fmt ? PF_BYTE_RGBA : PF_BYTE_RGB, dst);
    g->win->copyContentsToMemory(box, box, RenderTarget::FB_AUTO);
    return 0;
    CATCH(1)
}

// Window read-back in an arbitrary OGRE pixel format through the plugin (device-side conversion for the 32-bit 8-bit-per-
// channel formats, host PixelUtil otherwise) next to the reference's way: PF_BYTE_RGB read-back converted by the real
// PixelUtil::bulkPixelConversion.  Returns the number of differing bytes (-1 on error); a sub-box with a padded
// destination pitch is read too.
API long ph_read_colour_format_check(const char* format_name)
{
    TRY
    const PixelFormat pf = PixelUtil::getFormatFromName(format_name, true);
    if (pf == PF_UNKNOWN) { g_err = "unknown pixel format"; return -1; }
    const size_t bpp = PixelUtil::getNumElemBytes(pf);
    const uint32_t w = g->w, h = g->h;
    std::vector<uint8_t> rgb((size_t)w * h * 3), want((size_t)w * h * bpp), got((size_t)w * h * bpp, 0xEE);
    PixelBox prgb(w, h, 1, PF_BYTE_RGB, rgb.data());
    g->win->copyContentsToMemory(Box(0, 0, w, h), prgb, RenderTarget::FB_AUTO);
    PixelBox pwant(w, h, 1, pf, want.data());
    PixelUtil::bulkPixelConversion(prgb, pwant);
    PixelBox pgot(w, h, 1, pf, got.data());
    g->win->copyContentsToMemory(Box(0, 0, w, h), pgot, RenderTarget::FB_AUTO);
    long bad = 0;
    for (size_t i = 0; i < want.size(); ++i) bad += want[i] != got[i];
    // sub-box into a destination with a row pitch wider than the box
    const uint32_t l = 3, t = 5, r = w - 7, b = h - 2, bw = r - l, bh = b - t, pitch = bw + 9;
    std::vector<uint8_t> sub((size_t)pitch * bh * bpp, 0xEE);
    PixelBox psub(bw, bh, 1, pf, sub.data());
    psub.rowPitch = pitch; psub.slicePitch = (size_t)pitch * bh;
    g->win->copyContentsToMemory(Box(l, t, r, b), psub, RenderTarget::FB_AUTO);
    for (uint32_t y = 0; y < bh; ++y) {
        for (size_t k = 0; k < (size_t)bw * bpp; ++k) bad += sub[(size_t)y * pitch * bpp + k] != want[((size_t)(t + y) * w + l) * bpp + k];
        for (size_t k = (size_t)bw * bpp; k < (size_t)pitch * bpp; ++k) bad += sub[(size_t)y * pitch * bpp + k] != 0xEE;      // padding untouched
    }
    return bad;
    CATCH(-1)
}

API int ph_read_colour_box(uint32_t l, uint32_t t, uint32_t r, uint32_t b, uint8_t* dst)
{
    TRY
    PixelBox box(r - l, b - t, 1, PF_BYTE_RGB, dst);
    g->win->copyContentsToMemory(Box(l, t, r, b), box, RenderTarget::FB_AUTO);
    return 0;
    CATCH(1)
}

// CUDA plugin only: the window's ogrecu_rt (depth is read through the C ABI by the test)
API void* ph_window_native_target(void)
{
    void* rt = nullptr;
    try { g->win->getCustomAttribute("OGRECU_RT", &rt); } catch (...) { rt = nullptr; }
    return rt;
}

// Render the current scene into a TU_RENDERTARGET texture of the window's size through the same
// camera and read it back with RenderTexture::copyContentsToMemory (-> HardwarePixelBuffer::blitToMemory).
API int ph_render_to_texture(uint8_t* dst_rgb)
{
    TRY
    if (!g->rtt) {
        g->rtt = TextureManager::getSingleton().createManual("harness_rtt", ResourceGroupManager::DEFAULT_RESOURCE_GROUP_NAME,
                                                             TEX_TYPE_2D, g->w, g->h, 0, PF_BYTE_RGBA, TU_RENDERTARGET);
        RenderTexture* rt = g->rtt->getBuffer()->getRenderTarget();
        rt->setAutoUpdated(false);
        Viewport* vp = rt->addViewport(g->cam);
        vp->setBackgroundColour(ColourValue(0.1f, 0.2f, 0.3f));
        vp->setOverlaysEnabled(false);
    }
    RenderTexture* rt = g->rtt->getBuffer()->getRenderTarget();
    rt->update();
    PixelBox box(g->w, g->h, 1, PF_BYTE_RGB, dst_rgb);
    rt->copyContentsToMemory(box, box, RenderTarget::FB_AUTO);
    return 0;
    CATCH(1)
}

// Scaled read-back of the render texture (RenderTexture::copyContentsToMemory into a PixelBox of another size ->
// HardwarePixelBuffer::blitToMemory -> scaling): through the plugin (resampled on the device) next to the reference's way —
// the unscaled PF_BYTE_RGB read-back resampled by the real Image::scale(FILTER_BILINEAR).  Source box (l,t,r,b), destination
// dw x dh in `format_name`, destination rows `pad` pixels wider than dw.  Returns differing bytes (-1 on error).
API long ph_rtt_scaled_check(uint32_t l, uint32_t t, uint32_t r, uint32_t b, uint32_t dw, uint32_t dh, uint32_t pad, const char* format_name)
{
    TRY
    if (!g->rtt) { g_err = "no rtt"; return -1; }
    const PixelFormat pf = PixelUtil::getFormatFromName(format_name, true);
    if (pf == PF_UNKNOWN) { g_err = "unknown pixel format"; return -1; }
    const size_t bpp = PixelUtil::getNumElemBytes(pf);
    RenderTexture* rt = g->rtt->getBuffer()->getRenderTarget();
    const uint32_t w = g->w, h = g->h, pitch = dw + pad;
    std::vector<uint8_t> rgb((size_t)w * h * 3), want((size_t)pitch * dh * bpp, 0xEE), got((size_t)pitch * dh * bpp, 0xEE);
    PixelBox prgb(w, h, 1, PF_BYTE_RGB, rgb.data());
    rt->copyContentsToMemory(Box(0, 0, w, h), prgb, RenderTarget::FB_AUTO);
    PixelBox pwant(dw, dh, 1, pf, want.data());
    pwant.rowPitch = pitch; pwant.slicePitch = (size_t)pitch * dh;
    Image::scale(prgb.getSubVolume(Box(l, t, r, b)), pwant, Image::FILTER_BILINEAR);
    PixelBox pgot(dw, dh, 1, pf, got.data());
    pgot.rowPitch = pitch; pgot.slicePitch = (size_t)pitch * dh;
    rt->copyContentsToMemory(Box(l, t, r, b), pgot, RenderTarget::FB_AUTO);
    long bad = 0;
    for (size_t i = 0; i < want.size(); ++i) bad += want[i] != got[i];
    return bad;
    CATCH(-1)
}

/* -------- pose animation: a hand-built pose-animated Entity (as oracle/ref_driver.cpp ref_pose_entity) evaluated twice -------- */
// `engine`: OGRE's own Entity::_updateAnimation on buffers of the render system under test (host arithmetic through our
// lock / unlock); `device`: CudaPoseAnimation (plugin class over ogrecu_pose_*) on a clone of the same vertex data.
// out6 arrays: nverts x (pos, normal).  CUDA plugin only.
API int ph_pose_entity(uint32_t nverts, const float* posnorm6, uint32_t nposes, const uint32_t* pose_off, const uint32_t* idx, const float* off3,
                       const float* nrm3, const float* influence, int with_normals, float* engine6, float* device6)
{
    TRY
    static int serial = 0;
    const String name = "ph_pose_" + StringConverter::toString(serial++);
    MeshPtr m = MeshManager::getSingleton().createManual(name, ResourceGroupManager::DEFAULT_RESOURCE_GROUP_NAME);
    SubMesh* sub = m->createSubMesh();
    sub->useSharedVertices = false;
    sub->vertexData = OGRE_NEW VertexData();
    sub->vertexData->vertexCount = nverts;
    sub->vertexData->vertexDeclaration->addElement(0, 0, VET_FLOAT3, VES_POSITION);
    sub->vertexData->vertexDeclaration->addElement(0, 12, VET_FLOAT3, VES_NORMAL);
    auto vb = HardwareBufferManager::getSingleton().createVertexBuffer(24, nverts, HBU_CPU_TO_GPU);
    vb->writeData(0, (size_t)nverts * 24, posnorm6, true);
    sub->vertexData->vertexBufferBinding->setBinding(0, vb);
    const uint16_t tri[3] = {0, (uint16_t)(nverts > 1 ? 1 : 0), (uint16_t)(nverts > 2 ? 2 : 0)};
    sub->indexData->indexBuffer = HardwareBufferManager::getSingleton().createIndexBuffer(HardwareIndexBuffer::IT_16BIT, 3, HBU_CPU_TO_GPU);
    sub->indexData->indexBuffer->writeData(0, 6, tri, true);
    sub->indexData->indexCount = 3;
    m->_setBounds(AxisAlignedBox(-100, -100, -100, 100, 100, 100));
    m->_setBoundingSphereRadius(200);
    for (uint32_t p = 0; p < nposes; ++p) {
        Pose* pose = m->createPose(1, "pose" + StringConverter::toString(p));
        for (uint32_t k = pose_off[p]; k < pose_off[p + 1]; ++k) {
            const Vector3f o(off3[3 * k], off3[3 * k + 1], off3[3 * k + 2]);
            if (with_normals) pose->addVertex(idx[k], o, Vector3f(nrm3[3 * k], nrm3[3 * k + 1], nrm3[3 * k + 2]));
            else pose->addVertex(idx[k], o);
        }
    }
    Animation* anim = m->createAnimation("manual", 0);
    VertexAnimationTrack* track = anim->createVertexTrack(1, VAT_POSE);
    VertexPoseKeyFrame* kf = track->createVertexPoseKeyFrame(0);
    for (uint32_t p = 0; p < nposes; ++p) kf->addPoseReference((ushort)p, influence[p]);
    m->load();
    Entity* e = g->sm->createEntity(m);
    g->sm->getRootSceneNode()->createChildSceneNode()->attachObject(e);
    AnimationState* as = e->getAnimationState("manual");
    as->setEnabled(true);
    as->setTimePosition(0);
    e->addSoftwareAnimationRequest(with_normals != 0);
    g->root->renderOneFrame();
    e->_updateAnimation();
    VertexData* vd = e->getSubEntity(0)->_getSoftwareVertexAnimVertexData();
    auto ob = vd->vertexBufferBinding->getBuffer(vd->vertexDeclaration->findElementBySemantic(VES_POSITION)->getSource());
    if (ob->getVertexSize() != 24) { g_err = "unexpected vertex size"; return 3; }
    ob->readData(0, (size_t)nverts * 24, engine6);
    // the same frame on the device, into a clone made the way Entity makes its own (Entity::cloneVertexDataRemoveBlendInfo)
    CudaPoseAnimation dev(m.get());
    std::unique_ptr<VertexData> clone(sub->vertexData->clone(true));
    dev.initialise(sub->vertexData, clone.get(), with_normals != 0);
    dev.applyKeyFrames(kf, kf, 0.0f, as->getWeight(), clone.get());
    if (with_normals) dev.finalise(sub->vertexData, clone.get());
    auto cb = clone->vertexBufferBinding->getBuffer(0);
    cb->readData(0, (size_t)nverts * 24, device6);
    e->setVisible(false);
    return 0;
    CATCH(1)
}

// Show the render texture on a full-window quad (textured, unlit) and render the window: exercises
// sampling a TU_RENDERTARGET texture.
API int ph_show_rtt_on_quad(void)
{
    TRY
    if (!g->rtt) { g_err = "no rtt"; return 2; }
    MaterialPtr mat = MaterialManager::getSingleton().getByName("harness_rtt_mat");
    if (!mat) {
        mat = MaterialManager::getSingleton().create("harness_rtt_mat", ResourceGroupManager::DEFAULT_RESOURCE_GROUP_NAME);
        Pass* p = mat->getTechnique(0)->getPass(0);
        p->setLightingEnabled(false);
        p->createTextureUnitState()->setTexture(g->rtt);
    }
    if (g->ent) g->ent->setVisible(false);
    ManualObject* mo = g->sm->hasManualObject("quad") ? g->sm->getManualObject("quad") : nullptr;
    if (!mo) {
        mo = g->sm->createManualObject("quad");
        mo->begin("harness_rtt_mat", RenderOperation::OT_TRIANGLE_LIST);
        const float s = 40.0f;
        mo->position(-s, -s, 0); mo->textureCoord(0, 1);
        mo->position(s, -s, 0); mo->textureCoord(1, 1);
        mo->position(s, s, 0); mo->textureCoord(1, 0);
        mo->position(-s, s, 0); mo->textureCoord(0, 0);
        mo->triangle(0, 2, 1); mo->triangle(0, 3, 2);    // Tiny culls cross>0: wind to survive
        mo->triangle(0, 1, 2); mo->triangle(0, 2, 3);
        mo->end();
        g->sm->getRootSceneNode()->createChildSceneNode()->attachObject(mo);
    }
    g->root->renderOneFrame();
    return 0;
    CATCH(1)
}

/* -------- skeletal entity: pose, let OGRE software-skin it (OptimisedUtil), read the result ---- */

static std::vector<std::vector<uint8_t>> g_skinned;

static void grab(const VertexData* skinned)
{
    const VertexElement* dp = skinned->vertexDeclaration->findElementBySemantic(VES_POSITION);
    auto db = skinned->vertexBufferBinding->getBuffer(dp->getSource());
    const uint8_t* p = (const uint8_t*)db->lock(HardwareBuffer::HBL_READ_ONLY);
    g_skinned.emplace_back(p + dp->getOffset(), p + db->getSizeInBytes());
    db->unlock();
}

API int ph_skel_pose(float time)
{
    TRY
    Entity* e = g->ent;
    if (!e || !e->hasSkeleton()) { g_err = "no skeletal entity"; return 2; }
    // the bone palette is refreshed once per frame number (Entity::cacheBoneMatrices): advance it
    g->root->renderOneFrame();
    AnimationStateSet* set = e->getAllAnimationStates();
    bool first = true;
    for (auto& kv : set->getAnimationStates()) {
        kv.second->setEnabled(first);
        if (first) { kv.second->setLoop(true); kv.second->setTimePosition(time); }
        first = false;
    }
    e->addSoftwareAnimationRequest(true);
    e->_updateAnimation();
    g_skinned.clear();
    MeshPtr m = e->getMesh();
    if (m->sharedVertexData) grab(e->_getSkelAnimVertexData());
    for (size_t i = 0; i < m->getNumSubMeshes(); ++i) {
        if (m->getSubMesh(i)->useSharedVertices) continue;
        grab(e->getSubEntity(i)->_getSkelAnimVertexData());
    }
    return 0;
    CATCH(1)
}

// Bone palettes of the entity's first animation at n time positions: `engine` from the real SkeletonInstance
// (setAnimationState + _getBoneMatrices, one pose after the other), `device` from the plugin's CudaSkeletonPalettes
// (all n AnimationStateSets in one launch).  Both n x numBones x 16 floats.  CUDA plugin only.
API int ph_skel_palettes(uint32_t n, const float* times, float* engine, float* device)
{
    TRY
    Entity* e = g->ent;
    if (!e || !e->hasSkeleton()) { g_err = "no skeletal entity"; return 2; }
    SkeletonInstance* sk = e->getSkeleton();
    const size_t nb = sk->getNumBones();
    AnimationStateSet* set = e->getAllAnimationStates();
    AnimationState* first = nullptr;
    for (auto& kv : set->getAnimationStates()) { kv.second->setEnabled(false); if (!first) first = kv.second; }
    first->setLoop(true); first->setEnabled(true); first->setWeight(1.0f);
    std::vector<std::unique_ptr<AnimationStateSet>> copies;
    std::vector<const AnimationStateSet*> ptrs;
    std::vector<float> mask(nb);
    for (uint32_t i = 0; i < n; ++i) {
        first->setTimePosition(times[i]);
        // every other pose with a blend mask on the state (Animation::apply with a mask, OgreAnimation.cpp:344-357)
        const bool masked = (i & 1u) != 0;
        if (masked) {
            for (size_t b = 0; b < nb; ++b) mask[b] = float((b * 7 + i) % 5) * 0.25f;
            first->createBlendMask(nb, 1.0f);
            for (size_t b = 0; b < nb; ++b) first->setBlendMaskEntry(b, mask[b]);
        }
        sk->setAnimationState(*set);
        sk->_getBoneMatrices(reinterpret_cast<Affine3*>(engine) + i * nb);
        copies.emplace_back(new AnimationStateSet(*set));
        if (masked) {                                            // (the copy constructor of AnimationState drops the mask)
            AnimationState* c = copies.back()->getAnimationState(first->getAnimationName());
            c->createBlendMask(nb, 1.0f);
            for (size_t b = 0; b < nb; ++b) c->setBlendMaskEntry(b, mask[b]);
            first->destroyBlendMask();
        }
        ptrs.push_back(copies.back().get());
    }
    CudaSkeletonPalettes pal(sk);
    pal.getBoneMatrices(ptrs.data(), n, reinterpret_cast<Affine3*>(device));
    return 0;
    CATCH(1)
}
API size_t ph_skel_num_bones(void) { return g->ent && g->ent->hasSkeleton() ? g->ent->getSkeleton()->getNumBones() : 0; }

// generateShadowVolume side by side on the entity's own edge list: the real ShadowCaster::generateShadowVolume (a
// ShadowCaster subclass reaches the protected member) into one index buffer, CudaShadowVolume::generate into another,
// both created by the render system under test.  Outputs: the two index buffers (n16 16-bit indices each) and the
// ShadowRenderable ranges (per group indexStart, indexCount, cap start, cap count).  CUDA plugin only.
struct HarnessCaster : public ShadowCaster {
    AxisAlignedBox box;
    bool getCastShadows() const override { return true; }
    const AxisAlignedBox& getWorldBoundingBox(bool) const override { return box; }
    Real getPointExtrusionDistance(const Light*) const override { return 0; }
    void run(EdgeData* ed, const HardwareIndexBufferSharedPtr& ib, size_t& used, const Light* l, ShadowRenderableList& list, unsigned long flags)
    {
        updateEdgeListLightFacing(ed, l->getAs4DVector());
        generateShadowVolume(ed, ib, used, l, list, flags);
    }
};

API int ph_shadow_volume(const float* light4, const float* bounds6, uint32_t flags, uint32_t separate, uint32_t n16, uint32_t used_in,
                         uint16_t* engine, uint16_t* device, uint32_t* ranges_engine, uint32_t* ranges_device, uint32_t* used_out2)
{
    TRY
    Entity* e = g->ent;
    MeshPtr m = e->getMesh();
    if (!m->isEdgeListBuilt()) m->buildEdgeList();
    EdgeData* ed = m->getEdgeList(0);
    if (!ed) { g_err = "no edge list"; return 2; }
    static Light* light = nullptr;
    static SceneNode* ln = nullptr;
    if (!light) { light = g->sm->createLight("shadow_probe"); ln = g->sm->getRootSceneNode()->createChildSceneNode(); ln->attachObject(light); }
    if (light4[3] == 0.0f) {
        light->setType(Light::LT_DIRECTIONAL);
        ln->setPosition(0, 0, 0);
        ln->setDirection(Vector3(-light4[0], -light4[1], -light4[2]), Node::TS_WORLD);
    } else {
        light->setType(Light::LT_POINT);
        ln->resetOrientation();
        ln->setPosition(light4[0], light4[1], light4[2]);
    }
    ln->_update(true, true);
    HarnessCaster caster;
    caster.box.setExtents(bounds6[0], bounds6[1], bounds6[2], bounds6[3], bounds6[4], bounds6[5]);
    HardwareIndexBufferSharedPtr ib[2];
    for (int k = 0; k < 2; ++k) {
        ib[k] = HardwareBufferManager::getSingleton().createIndexBuffer(HardwareIndexBuffer::IT_16BIT, n16, HBU_CPU_TO_GPU);
        uint16_t* p = (uint16_t*)ib[k]->lock(HardwareBuffer::HBL_DISCARD);
        for (uint32_t i = 0; i < n16; ++i) p[i] = 0xCDCD;
        ib[k]->unlock();
    }
    auto collect = [&](ShadowRenderableList& list, uint32_t* out) {
        size_t q = 0;
        for (auto* sr : list) {
            IndexData* id = sr->getRenderOperationForUpdate()->indexData;
            out[q++] = (uint32_t)id->indexStart; out[q++] = (uint32_t)id->indexCount;
            if (sr->isLightCapSeparate()) {
                IndexData* ic = sr->getLightCapRenderable()->getRenderOperationForUpdate()->indexData;
                out[q++] = (uint32_t)ic->indexStart; out[q++] = (uint32_t)ic->indexCount;
            } else { out[q++] = 0; out[q++] = 0; }
        }
    };
    for (int k = 0; k < 2; ++k) {
        ShadowRenderableList list;
        for (auto& eg : ed->edgeGroups) list.push_back(OGRE_NEW ShadowRenderable(nullptr, ib[k], eg.vertexData, separate != 0));
        size_t used = used_in;
        if (k == 0) caster.run(ed, ib[0], used, light, list, flags);
        else { CudaShadowVolume sv(ed); sv.generate(ib[1], used, light, caster.box, list, flags); }
        used_out2[k] = (uint32_t)used;
        collect(list, k == 0 ? ranges_engine : ranges_device);
        const uint16_t* p = (const uint16_t*)ib[k]->lock(HardwareBuffer::HBL_READ_ONLY);
        memcpy(k == 0 ? engine : device, p, (size_t)n16 * 2);
        ib[k]->unlock();
        for (auto* sr : list) OGRE_DELETE sr;
    }
    return 0;
    CATCH(1)
}
API size_t ph_shadow_num_groups(void)
{
    MeshPtr m = g->ent->getMesh();
    if (!m->isEdgeListBuilt()) m->buildEdgeList();
    EdgeData* ed = m->getEdgeList(0);
    return ed ? ed->edgeGroups.size() : 0;
}

API size_t ph_skel_num_subs(void) { return g_skinned.size(); }
API size_t ph_skel_sub_size(size_t i) { return g_skinned[i].size(); }
API const uint8_t* ph_skel_sub_data(size_t i) { return g_skinned[i].data(); }

API double ph_frame_time(int frames)
{
    g->root->renderOneFrame();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < frames; ++i) g->root->renderOneFrame();
    // include the device work: a 1-pixel readback synchronises
    uint8_t px[3];
    PixelBox box(1, 1, 1, PF_BYTE_RGB, px);
    g->win->copyContentsToMemory(Box(0, 0, 1, 1), box, RenderTarget::FB_AUTO);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / frames;
}

// Seconds per frame of the skeletally animated entity: every frame advances the first animation state by dt, so
// Entity::_updateAnimation re-skins every sub-mesh through OptimisedUtil before the scene is rendered.
API double ph_anim_frame_time(int frames, float dt)
{
    Entity* e = g->ent;
    if (!e || !e->hasSkeleton()) return -1.0;
    AnimationState* st = nullptr;
    for (auto& kv : e->getAllAnimationStates()->getAnimationStates()) {
        kv.second->setEnabled(st == nullptr);
        if (!st) { st = kv.second; st->setLoop(true); }
    }
    e->addSoftwareAnimationRequest(true);
    uint8_t px[3];
    PixelBox box(1, 1, 1, PF_BYTE_RGB, px);
    for (int i = 0; i < 3; ++i) { st->addTime(dt); g->root->renderOneFrame(); }
    g->win->copyContentsToMemory(Box(0, 0, 1, 1), box, RenderTarget::FB_AUTO);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < frames; ++i) { st->addTime(dt); g->root->renderOneFrame(); }
    g->win->copyContentsToMemory(Box(0, 0, 1, 1), box, RenderTarget::FB_AUTO);      // includes the device work
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / frames;
}

// A crowd: `count` entities of one skeletally animated mesh on a square grid (`spacing` apart), every one with its own
// animation state (time offset i * 0.037 s), seen from above and behind.  The scene BASELINE config 2 describes, through the
// engine: Root::renderOneFrame -> SceneManager -> Entity::_updateAnimation (skeleton on the CPU, OptimisedUtil per sub-entity)
// -> _render per sub-entity.
static std::vector<Entity*> g_crowd;
static std::vector<AnimationState*> g_crowd_states;

API int ph_crowd_open(const char* mesh, uint32_t count, float spacing)
{
    TRY
    if (ph_scene_open(mesh, 100.0f) != 0) return 1;
    g_crowd.clear(); g_crowd_states.clear();
    SceneManager* sm = g->sm;
    const uint32_t side = (uint32_t)std::ceil(std::sqrt((double)count));
    const float half = 0.5f * spacing * (side - 1);
    g->ent->getParentSceneNode()->setPosition(-half, 0, -half);
    g_crowd.push_back(g->ent);
    for (uint32_t i = 1; i < count; ++i) {
        Entity* e = sm->createEntity(mesh);
        SceneNode* n = sm->getRootSceneNode()->createChildSceneNode();
        n->setPosition(spacing * (i % side) - half, 0, spacing * (i / side) - half);
        n->attachObject(e);
        g_crowd.push_back(e);
    }
    for (size_t i = 0; i < g_crowd.size(); ++i) {
        Entity* e = g_crowd[i];
        if (!e->hasSkeleton()) return 2;
        AnimationState* st = nullptr;
        for (auto& kv : e->getAllAnimationStates()->getAnimationStates()) {
            kv.second->setEnabled(st == nullptr);
            if (!st) { st = kv.second; st->setLoop(true); }
        }
        st->setTimePosition(0.037f * (float)i);
        e->addSoftwareAnimationRequest(true);
        // pinned submission order (see ph_pin_submission_order): sub-entity k of every entity in priority group 100 + (n-1-k)
        const size_t nsub = e->getNumSubEntities();
        for (size_t k = 0; k < nsub; ++k) e->getSubEntity(k)->setRenderQueueGroupAndPriority(RENDER_QUEUE_MAIN, (ushort)(100 + nsub - 1 - k));
        g_crowd_states.push_back(st);
    }
    SceneNode* cn = g->cam->getParentSceneNode();
    const float d = std::max(60.0f, 1.6f * half + 40.0f);
    cn->setPosition(0, 0.8f * d, d);
    cn->lookAt(Vector3(0, 0, 0), Node::TS_WORLD);
    g->cam->setFarClipDistance(0);      // infinite
    return 0;
    CATCH(3)
}

API size_t ph_crowd_vertices(void)
{
    size_t n = 0;
    for (Entity* e : g_crowd)
        for (size_t k = 0; k < e->getNumSubEntities(); ++k) {
            const SubMesh* sm = e->getSubEntity(k)->getSubMesh();
            n += sm->useSharedVertices ? 0 : sm->vertexData->vertexCount;
        }
    if (!g_crowd.empty() && g_crowd[0]->getMesh()->sharedVertexData) n += g_crowd.size() * g_crowd[0]->getMesh()->sharedVertexData->vertexCount;
    return n;
}

// seconds per frame of the crowd: every entity's animation advances by dt, so every frame re-evaluates every skeleton and
// re-skins every sub-mesh; ends with a 1-pixel read-back so that the device work is inside the clock
API double ph_crowd_frame_time(int frames, float dt)
{
    if (g_crowd.empty()) return -1.0;
    uint8_t px[3];
    PixelBox box(1, 1, 1, PF_BYTE_RGB, px);
    for (int i = 0; i < 2; ++i) {
        for (AnimationState* st : g_crowd_states) st->addTime(dt);
        g->root->renderOneFrame();
    }
    g->win->copyContentsToMemory(Box(0, 0, 1, 1), box, RenderTarget::FB_AUTO);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < frames; ++i) {
        for (AnimationState* st : g_crowd_states) st->addTime(dt);
        g->root->renderOneFrame();
    }
    g->win->copyContentsToMemory(Box(0, 0, 1, 1), box, RenderTarget::FB_AUTO);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / frames;
}

// one more crowd frame at a fixed pose (every state set to offset + t), then the whole frame as RGB
API int ph_crowd_render_at(float t, uint8_t* dst_rgb)
{
    TRY
    for (size_t i = 0; i < g_crowd_states.size(); ++i) g_crowd_states[i]->setTimePosition(0.037f * (float)i + t);
    g->root->renderOneFrame();
    PixelBox box(g->win->getWidth(), g->win->getHeight(), 1, PF_BYTE_RGB, dst_rgb);
    g->win->copyContentsToMemory(Box(0, 0, g->win->getWidth(), g->win->getHeight()), box, RenderTarget::FB_AUTO);
    return 0;
    CATCH(1)
}

// CUDA plugin only: its "OptimisedUtil Batching" option at run time
API int ph_set_rs_option(const char* name, const char* value)
{
    TRY
    g->root->getRenderSystem()->setConfigOption(name, value);
    return 0;
    CATCH(1)
}

// Submission order of the entity's sub-entities.  OGRE's render queue groups renderables by Pass and orders the groups by
// pass hash, then by Pass* ADDRESS (QueuedRenderableCollection::PassGroupLess, OgreMain/src/OgreRenderQueueSortingGrouping.cpp):
// passes whose hashes tie are drawn in an order that depends on the heap layout of the run.  Tiny (and this plugin) let a later
// fragment of EQUAL depth win, so a pixel where two sub-meshes coincide depends on that order.  Distinct priorities pin it:
// mode 0 = sub-entity i gets priority 100 + i (ascending), 1 = descending.
API int ph_pin_submission_order(int mode)
{
    TRY
    if (!g->ent) return 1;
    const size_t n = g->ent->getNumSubEntities();
    for (size_t i = 0; i < n; ++i)
        g->ent->getSubEntity(i)->setRenderQueueGroupAndPriority(RENDER_QUEUE_MAIN, (ushort)(100 + (mode ? n - 1 - i : i)));
    return 0;
    CATCH(2)
}

// The SSE control/status register of this thread: rounding mode, flush-to-zero (bit 15) and denormals-are-zero (bit 6).
// Libraries loaded into the host process may change it; the reference's arithmetic depends on it (tools/tiny_host_check.py).
API unsigned ph_get_mxcsr(void) { return _mm_getcsr(); }
API void ph_set_mxcsr(unsigned v) { _mm_setcsr(v); }

API void ph_close(void)
{
    if (!g) return;
    g->rtt.reset();
    g->root.reset();
    delete g;
    g = nullptr;
}
