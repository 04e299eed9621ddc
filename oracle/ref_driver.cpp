// ref_driver.cpp — drives the UNMODIFIED reference (compiled by oracle/build_ref.py) for the
// tests and for bench.py's reference arm.  TEST INFRASTRUCTURE ONLY: the product never links
// or loads this.
//
// What it exposes (plain C, loaded with ctypes):
//   * ref_skin/ref_morph/...   the real Ogre::OptimisedUtil implementations: SSE is the object
//                              OptimisedUtil::getImplementation() returns in stock libOgreMain.so;
//                              General is OgreMain/src/OgreOptimisedUtilGeneral.cpp compiled into
//                              this driver from where it lies (its accessor is hidden in the .so).
//   * ref_tiny_*               the real TinyRenderSystem: RenderSystems/Tiny/src/*.cpp are compiled
//                              into this driver (with -fno-access-control so a subclass can set and
//                              read the private DefaultShader uniforms) and _render() is called with
//                              caller-supplied state -> pins oracle/ogre_oracle.c draw for draw.
//   * ref_scene_*              a real Root/SceneManager scene (ogrehead) rendered by
//                              that TinyRenderSystem while every _render call is captured (buffers,
//                              uniforms, state) -> tests/golden fixtures.
//   * ref_skel_*               a real skeletally animated Entity (Sinbad / jaiqua): source streams,
//                              blend data, bone palettes and OGRE's own software-skinned output.
//   * ref_bench_*              timing loops for the reference arm (host threads stated by caller).
#include <Ogre.h>
#include <OgreOptimisedUtil.h>
#include <OgreEdgeListBuilder.h>
#include <OgreDefaultHardwareBufferManager.h>
#include <OgreSkeletonInstance.h>
#include <OgreSubMesh.h>
#include <OgreSubEntity.h>
#include <OgreTagPoint.h>

#include "OgreTinyRenderSystem.h"
#include "OgreTinyWindow.h"
#include "OgreTinyDepthBuffer.h"
#include "OgreTinyTexture.h"

#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <thread>
#include <vector>

#include "ogre_oracle.h"

namespace Ogre { extern OptimisedUtil* _getOptimisedUtilGeneral(void); }

using namespace Ogre;

#define API extern "C" __attribute__((visibility("default")))

/* ===================================================================== OptimisedUtil */

static OptimisedUtil* g_sse = OptimisedUtil::getImplementation();   // before anything replaces it

static OptimisedUtil* impl(int which) { return which == 0 ? g_sse : _getOptimisedUtilGeneral(); }

API int ref_is_sse_active(void)
{
    // the stock selection picked SSE iff the pointer differs from what a General-only build returns;
    // all we can say from outside is that it is non-null
    return g_sse != nullptr;
}

API void ref_skin(int which, const float* sp, float* dp, const float* sn, float* dn, const float* bw, const uint8_t* bi,
                  const float* const* mats, size_t sps, size_t dps, size_t sns, size_t dns, size_t bws, size_t bis, size_t nw, size_t n)
{
    impl(which)->softwareVertexSkinning(sp, dp, sn, dn, bw, bi, (const Affine3* const*)mats, sps, dps, sns, dns, bws, bis, nw, n);
}

API void ref_morph(int which, float t, const float* a, const float* b, float* d, size_t s1, size_t s2, size_t sd, size_t n, int normals)
{
    impl(which)->softwareVertexMorph(t, a, b, d, s1, s2, sd, n, normals != 0);
}

API void ref_concat(int which, const float* base, const float* src, float* dst, size_t n)
{
    impl(which)->concatenateAffineMatrices(*(const Affine3*)base, (const Affine3*)src, (Affine3*)dst, n);
}

API void ref_face_normals(int which, const float* pos, const void* tris, float* out, size_t n)
{
    impl(which)->calculateFaceNormals(pos, (const EdgeData::Triangle*)tris, (Vector4*)out, n);
}

API void ref_light_facing(int which, const float* l, const float* fn, char* out, size_t n)
{
    impl(which)->calculateLightFacing(*(const Vector4*)l, (const Vector4*)fn, out, n);
}

API void ref_extrude(int which, const float* l, float dist, const float* src, float* dst, size_t n)
{
    impl(which)->extrudeVertices(*(const Vector4*)l, dist, src, dst, n);
}

API size_t ref_sizeof_edge_triangle(void) { return sizeof(EdgeData::Triangle); }

/* ========================================================================== Tiny ==== */

struct CapturedDraw {
    oracle_draw d;                  // pointers into the copies below
    std::vector<uint8_t> vbuf[3];   // position / uv / normal source buffers (whole)
    std::vector<uint8_t> ibuf;
    int tex_id;                     // index into g_tex_copies or -1
    size_t elem_off[3];             // byte offset of the first element inside vbuf[k]
};

struct TexCopy { std::vector<uint8_t> rgba; uint32_t w, h; const Image* src; };

static std::vector<std::unique_ptr<CapturedDraw>> g_captured;
static std::vector<TexCopy> g_tex_copies;
static bool g_capture_on = false;

class CaptureTiny : public TinyRenderSystem
{
public:
    void _render(const RenderOperation& op) override
    {
        if (g_capture_on) capture(op);
        TinyRenderSystem::_render(op);
    }

    // state exactly as TinyRenderSystem holds it when _render runs (private members; this TU is
    // compiled with -fno-access-control)
    void capture(const RenderOperation& op)
    {
        bool isStrip = op.operationType == RenderOperation::OT_TRIANGLE_STRIP;
        if (!isStrip && op.operationType != RenderOperation::OT_TRIANGLE_LIST) return;
        std::unique_ptr<CapturedDraw> c(new CapturedDraw());
        memset(&c->d, 0, sizeof c->d);
        const VertexElementSemantic sems[3] = {VES_POSITION, VES_TEXTURE_COORDINATES, VES_NORMAL};
        const uint8_t* ptrs[3] = {0, 0, 0};
        uint32_t strides[3] = {0, 0, 0};
        for (int k = 0; k < 3; ++k) {
            auto el = op.vertexData->vertexDeclaration->findElementBySemantic(sems[k]);
            if (!el) continue;
            auto buf = op.vertexData->vertexBufferBinding->getBuffer(el->getSource());
            size_t step = op.vertexData->vertexDeclaration->getVertexSize(el->getSource());
            const uint8_t* base = (const uint8_t*)buf->lock(HardwareBuffer::HBL_READ_ONLY);
            c->vbuf[k].assign(base, base + buf->getSizeInBytes());
            buf->unlock();
            c->elem_off[k] = el->getOffset() + op.vertexData->vertexStart * step;
            ptrs[k] = c->vbuf[k].data() + c->elem_off[k];
            strides[k] = (uint32_t)step;
        }
        c->d.positions = ptrs[0]; c->d.pos_stride = strides[0];
        c->d.uvs = ptrs[1]; c->d.uv_stride = strides[1];
        c->d.normals = ptrs[2]; c->d.normal_stride = strides[2];
        c->d.draw_count = (uint32_t)op.vertexData->vertexCount;
        if (op.useIndexes) {
            auto ib = op.indexData->indexBuffer;
            const uint8_t* base = (const uint8_t*)ib->lock(HardwareBuffer::HBL_READ_ONLY);
            c->ibuf.assign(base, base + ib->getSizeInBytes());
            ib->unlock();
            c->d.index_type = (uint32_t)ib->getIndexSize();
            c->d.indices = c->ibuf.data() + op.indexData->indexStart * ib->getIndexSize();
            c->d.draw_count = (uint32_t)op.indexData->indexCount;
        }
        c->d.is_strip = isStrip;
        c->d.depth_test = mDepthTest; c->d.depth_write = mDepthWrite; c->d.blend_add = mBlendAdd;
        c->d.lighting = mDefaultShader.uniform_doLighting && ptrs[2];
        c->tex_id = -1;
        if (mDefaultShader.image) {
            const Image* img = mDefaultShader.image;
            for (size_t i = 0; i < g_tex_copies.size(); ++i) if (g_tex_copies[i].src == img) c->tex_id = (int)i;
            if (c->tex_id < 0) {
                TexCopy t;
                t.src = img; t.w = img->getWidth(); t.h = img->getHeight();
                t.rgba.assign(img->getData(), img->getData() + (size_t)t.w * t.h * 4);
                g_tex_copies.push_back(std::move(t));
                c->tex_id = (int)g_tex_copies.size() - 1;
            }
            c->d.tex_w = img->getWidth(); c->d.tex_h = img->getHeight();
        }
        memcpy(c->d.mvp, mDefaultShader.uniform_MVP[0], 64);
        memcpy(c->d.tex_matrix, mDefaultShader.uniform_Tex[0], 64);
        memcpy(c->d.mvit, mDefaultShader.uniform_MVIT[0], 64);
        memcpy(c->d.viewport, mVP[0], 64);
        memcpy(c->d.light_dir, mDefaultShader.uniform_lightDir.ptr(), 12);
        memcpy(c->d.ambient, mDefaultShader.uniform_ambientCol.ptr(), 16);
        g_captured.push_back(std::move(c));
    }
};

struct TinyHost {
    std::unique_ptr<Root> root;
    CaptureTiny* rs = nullptr;
    RenderWindow* win = nullptr;
    uint32_t w = 0, h = 0;
    bool resources_ready = false;
};
static TinyHost* g_host = nullptr;
static std::string g_media;

static void host_open(uint32_t w, uint32_t h)
{
    if (g_host && g_host->w == w && g_host->h == h) return;
    if (!g_host) {
        g_host = new TinyHost();
        // silent log (no file, no stderr) created before Root so that Root adopts it
        LogManager* lm = new LogManager();
        lm->createLog("ref_driver.log", true, false, true);
        g_host->root.reset(new Root("", "", ""));       // no plugins.cfg, no ogre.cfg
        g_host->rs = new CaptureTiny();
        g_host->root->addRenderSystem(g_host->rs);
        g_host->root->setRenderSystem(g_host->rs);
        g_host->root->initialise(false);
    } else if (g_host->win) {
        g_host->root->destroyRenderTarget(g_host->win);
        g_host->win = nullptr;
    }
    static int serial = 0;
    g_host->win = g_host->root->createRenderWindow("ref" + std::to_string(serial++), w, h, false);
    g_host->w = w; g_host->h = h;
}

API int ref_tiny_open(uint32_t w, uint32_t h)
{
    try { host_open(w, h); g_host->rs->_setRenderTarget(g_host->win); return 0; }
    catch (std::exception& e) { fprintf(stderr, "ref_tiny_open: %s\n", e.what()); return 1; }
}

API void ref_tiny_clear(unsigned buffers, const float* rgba, float depth)
{
    g_host->rs->_setRenderTarget(g_host->win);
    unsigned fb = ((buffers & 1) ? FBT_COLOUR : 0) | ((buffers & 2) ? FBT_DEPTH : 0);
    g_host->rs->clearFrameBuffer(fb, ColourValue(rgba[0], rgba[1], rgba[2], rgba[3]), depth, 0);
}

// upload initial contents of the planes (for no-clear tests)
API void ref_tiny_write(const uint8_t* colour_rgb, const float* depth)
{
    g_host->rs->_setRenderTarget(g_host->win);
    if (colour_rgb) memcpy(g_host->rs->mActiveColourBuffer->getData(), colour_rgb, (size_t)g_host->w * g_host->h * 3);
    if (depth) memcpy(g_host->rs->mActiveDepthBuffer->getData(), depth, (size_t)g_host->w * g_host->h * 4);
}

API void ref_tiny_read(uint8_t* colour_rgb, float* depth)
{
    // colour through the public readback path, depth through TinyDepthBuffer::getImage()
    if (colour_rgb) {
        PixelBox dst(g_host->w, g_host->h, 1, PF_BYTE_RGB, colour_rgb);
        g_host->win->copyContentsToMemory(dst, dst, RenderTarget::FB_AUTO);
    }
    if (depth) {
        Image* z = static_cast<TinyDepthBuffer*>(g_host->win->getDepthBuffer())->getImage();
        memcpy(depth, z->getData(), (size_t)g_host->w * g_host->h * 4);
    }
}

// One real TinyRenderSystem::_render with the given state (buffers are created through the
// render system's own DefaultHardwareBufferManager).
API int ref_tiny_draw(const oracle_draw* d, size_t num_vertices_in_buffers)
{
    try {
        CaptureTiny* rs = g_host->rs;
        rs->_setRenderTarget(g_host->win);
        HardwareBufferManager& hbm = HardwareBufferManager::getSingleton();
        // the oracle_draw streams may be interleaved in one allocation; rebuild one buffer per stream
        VertexData vd;
        vd.vertexStart = 0;
        vd.vertexCount = num_vertices_in_buffers;
        const uint8_t* src[3] = {d->positions, d->uvs, d->normals};
        const uint32_t stride[3] = {d->pos_stride, d->uv_stride, d->normal_stride};
        const VertexElementType types[3] = {VET_FLOAT3, VET_FLOAT2, VET_FLOAT3};
        const VertexElementSemantic sems[3] = {VES_POSITION, VES_TEXTURE_COORDINATES, VES_NORMAL};
        const size_t elsize[3] = {12, 8, 12};
        unsigned short source = 0;
        for (int k = 0; k < 3; ++k) {
            if (!src[k]) continue;
            // keep the caller's stride: element at offset 0 of a `stride`-byte vertex
            size_t bytes = (num_vertices_in_buffers - 1) * stride[k] + elsize[k];
            size_t nverts = (bytes + stride[k] - 1) / stride[k];
            auto vb = hbm.createVertexBuffer(stride[k], nverts, HBU_CPU_TO_GPU);
            void* p = vb->lock(HardwareBuffer::HBL_DISCARD);
            memcpy(p, src[k], bytes);
            vb->unlock();
            vd.vertexDeclaration->addElement(source, 0, types[k], sems[k]);
            // getVertexSize(source) must equal the stride: pad with a dummy element when wider
            size_t have = elsize[k];
            unsigned short pad = 0;
            while (have + 4 <= stride[k]) { vd.vertexDeclaration->addElement(source, have, VET_FLOAT1, VES_TEXTURE_COORDINATES, 1 + pad + 4 * k); have += 4; ++pad; }
            vd.vertexBufferBinding->setBinding(source, vb);
            ++source;
        }
        RenderOperation op;
        op.vertexData = &vd;
        op.operationType = d->is_strip ? RenderOperation::OT_TRIANGLE_STRIP : RenderOperation::OT_TRIANGLE_LIST;
        IndexData id;
        op.useIndexes = d->index_type != 0;
        if (op.useIndexes) {
            auto ib = hbm.createIndexBuffer(d->index_type == 2 ? HardwareIndexBuffer::IT_16BIT : HardwareIndexBuffer::IT_32BIT,
                                            d->draw_count, HBU_CPU_TO_GPU);
            void* p = ib->lock(HardwareBuffer::HBL_DISCARD);
            memcpy(p, d->indices, (size_t)d->draw_count * d->index_type);
            ib->unlock();
            id.indexBuffer = ib; id.indexStart = 0; id.indexCount = d->draw_count;
            op.indexData = &id;
        } else {
            vd.vertexCount = d->draw_count;
        }
        // state + uniforms, written where Tiny keeps them
        rs->mDepthTest = d->depth_test; rs->mDepthWrite = d->depth_write; rs->mBlendAdd = d->blend_add;
        rs->mDefaultShader.uniform_doLighting = d->lighting;
        rs->mDefaultShader.uniform_MVP = Matrix4(d->mvp);
        rs->mDefaultShader.uniform_Tex = Matrix4(d->tex_matrix);
        rs->mDefaultShader.uniform_MVIT = Matrix4(d->mvit);
        rs->mVP = Matrix4(d->viewport);
        rs->mDefaultShader.uniform_lightDir = Vector3(d->light_dir[0], d->light_dir[1], d->light_dir[2]);
        rs->mDefaultShader.uniform_ambientCol = ColourValue(d->ambient[0], d->ambient[1], d->ambient[2], d->ambient[3]);
        Image tex;
        if (d->tex_rgba) {
            tex.loadDynamicImage(const_cast<uchar*>(d->tex_rgba), d->tex_w, d->tex_h, 1, PF_BYTE_RGBA, false);
            rs->mDefaultShader.image = &tex;
        } else rs->mDefaultShader.image = NULL;
        rs->mCurrentPassIterationCount = 0;
        bool cap = g_capture_on; g_capture_on = false;
        rs->_render(op);
        g_capture_on = cap;
        rs->mDefaultShader.image = NULL;
        op.vertexData = 0; op.indexData = 0;   // stack objects
        return 0;
    } catch (std::exception& e) { fprintf(stderr, "ref_tiny_draw: %s\n", e.what()); return 1; }
}

/* ================================================================= real scenes ==== */

API void ref_set_media_dir(const char* dir) { g_media = dir; }

static void ensure_resources()
{
    if (g_host->resources_ready) return;
    // Codec_STBI (jpg/png) is a plugin; it sits next to libOgreMain.so
    std::string lib = g_media + "/../lib/Codec_STBI.so";
    g_host->root->loadPlugin(lib);
    ResourceGroupManager& rgm = ResourceGroupManager::getSingleton();
    rgm.addResourceLocation(g_media + "/models", "FileSystem");
    rgm.addResourceLocation(g_media + "/scripts", "FileSystem");
    rgm.addResourceLocation(g_media + "/textures", "FileSystem");
    rgm.addResourceLocation(g_media + "/packs/Sinbad.zip", "Zip");
    rgm.initialiseAllResourceGroups();
    g_host->resources_ready = true;
}

struct SceneState { SceneManager* sm = nullptr; Camera* cam = nullptr; Entity* ent = nullptr; Viewport* vp = nullptr; };
static SceneState g_scene;

// ogrehead, camera (0,0,140) -> origin, near 5, auto aspect, point light
// (20,80,50), ambient 0.3, background (0.1,0.2,0.3)
API int ref_scene_open(const char* mesh, uint32_t w, uint32_t h, float cam_z)
{
    try {
        host_open(w, h);
        ensure_resources();
        if (g_scene.sm) { g_host->win->removeAllViewports(); g_host->root->destroySceneManager(g_scene.sm); g_scene = SceneState(); }
        SceneManager* sm = g_host->root->createSceneManager();
        g_scene.sm = sm;
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
        g_scene.cam = cam;
        Viewport* vp = g_host->win->addViewport(cam);
        vp->setBackgroundColour(ColourValue(0.1f, 0.2f, 0.3f));
        g_scene.vp = vp;
        Entity* e = sm->createEntity(mesh);
        sm->getRootSceneNode()->createChildSceneNode()->attachObject(e);
        g_scene.ent = e;
        return 0;
    } catch (std::exception& ex) { fprintf(stderr, "ref_scene_open: %s\n", ex.what()); return 1; }
}

API int ref_scene_render(int capture)
{
    try {
        if (capture) { g_captured.clear(); g_tex_copies.clear(); }
        g_capture_on = capture != 0;
        g_host->root->renderOneFrame();
        g_capture_on = false;
        return 0;
    } catch (std::exception& ex) { g_capture_on = false; fprintf(stderr, "ref_scene_render: %s\n", ex.what()); return 1; }
}

API size_t ref_scene_triangles(void) { return g_host->win->getStatistics().triangleCount; }
API size_t ref_scene_batches(void) { return g_host->win->getStatistics().batchCount; }

API size_t ref_capture_count(void) { return g_captured.size(); }
API const oracle_draw* ref_capture_draw(size_t i) { return &g_captured[i]->d; }
API const uint8_t* ref_capture_vbuf(size_t i, int k, size_t* size, size_t* elem_off)
{
    *size = g_captured[i]->vbuf[k].size(); *elem_off = g_captured[i]->elem_off[k];
    return g_captured[i]->vbuf[k].data();
}
API const uint8_t* ref_capture_ibuf(size_t i, size_t* size, size_t* start_off)
{
    auto& c = *g_captured[i];
    *size = c.ibuf.size();
    *start_off = c.ibuf.empty() ? 0 : (size_t)((const uint8_t*)c.d.indices - c.ibuf.data());
    return c.ibuf.data();
}
API int ref_capture_tex(size_t i) { return g_captured[i]->tex_id; }
API size_t ref_capture_tex_count(void) { return g_tex_copies.size(); }
API const uint8_t* ref_capture_tex_data(size_t t, uint32_t* w, uint32_t* h) { *w = g_tex_copies[t].w; *h = g_tex_copies[t].h; return g_tex_copies[t].rgba.data(); }

// wall-clock seconds per frame of the real Tiny path (after one warm-up frame)
API double ref_scene_time(int frames)
{
    g_host->root->renderOneFrame();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < frames; ++i) g_host->root->renderOneFrame();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / frames;
}

/* ============================================================ skeletal entities ==== */

struct SkelSub {
    std::vector<uint8_t> src, blend;       // source pos(+normal) stream, blend stream (as OGRE lays them out)
    uint32_t src_stride, norm_off, blend_stride, weight_off, index_off, nw, nverts;
    int has_norm;
    std::vector<uint16_t> bone_map;        // blendIndexToBoneIndexMap
    std::vector<uint8_t> skinned;          // OGRE's own software-skinned result for the current pose
    uint32_t dst_stride;
};
static std::vector<SkelSub> g_subs;
static std::vector<float> g_bones;         // numBones x 16 for the current pose
static Entity* g_skel_ent = nullptr;

static void dump_vertex_data(const VertexData* vd, const VertexData* skinned, const Mesh::IndexMap& map, SkelSub& s)
{
    const VertexDeclaration* decl = vd->vertexDeclaration;
    const VertexElement* ep = decl->findElementBySemantic(VES_POSITION);
    const VertexElement* en = decl->findElementBySemantic(VES_NORMAL);
    const VertexElement* ei = decl->findElementBySemantic(VES_BLEND_INDICES);
    const VertexElement* ew = decl->findElementBySemantic(VES_BLEND_WEIGHTS);
    auto pb = vd->vertexBufferBinding->getBuffer(ep->getSource());
    s.nverts = (uint32_t)vd->vertexCount;
    s.src_stride = (uint32_t)pb->getVertexSize();
    s.has_norm = en && en->getSource() == ep->getSource();
    s.norm_off = s.has_norm ? (uint32_t)(en->getOffset() - ep->getOffset()) : 0;
    {
        const uint8_t* p = (const uint8_t*)pb->lock(HardwareBuffer::HBL_READ_ONLY);
        s.src.assign(p + vd->vertexStart * s.src_stride + ep->getOffset(), p + pb->getSizeInBytes());
        pb->unlock();
    }
    auto bb = vd->vertexBufferBinding->getBuffer(ei->getSource());
    s.blend_stride = (uint32_t)bb->getVertexSize();
    s.index_off = (uint32_t)ei->getOffset();
    s.weight_off = (uint32_t)ew->getOffset();
    s.nw = VertexElement::getTypeCount(ew->getType());
    {
        const uint8_t* p = (const uint8_t*)bb->lock(HardwareBuffer::HBL_READ_ONLY);
        s.blend.assign(p + vd->vertexStart * s.blend_stride, p + bb->getSizeInBytes());
        bb->unlock();
    }
    s.bone_map.assign(map.begin(), map.end());
    if (skinned) {
        const VertexElement* dp = skinned->vertexDeclaration->findElementBySemantic(VES_POSITION);
        auto db = skinned->vertexBufferBinding->getBuffer(dp->getSource());
        s.dst_stride = (uint32_t)db->getVertexSize();
        const uint8_t* p = (const uint8_t*)db->lock(HardwareBuffer::HBL_READ_ONLY);
        s.skinned.assign(p + dp->getOffset(), p + db->getSizeInBytes());
        db->unlock();
    }
}

// Load `mesh` as an Entity in a scene, pose its first animation at `time` and let OGRE run
// its own software skinning (Entity::_updateAnimation -> Mesh::softwareVertexBlend ->
// OptimisedUtil::softwareVertexSkinning, OgreEntity.cpp:747-911, OgreMesh.cpp:1792-1897).
API int ref_skel_open(const char* mesh, uint32_t w, uint32_t h)
{
    int rc = ref_scene_open(mesh, w, h, 30.0f);
    if (rc) return rc;
    g_skel_ent = g_scene.ent;
    return g_skel_ent->hasSkeleton() ? 0 : 2;
}

API int ref_skel_pose(float time)
{
    try {
        Entity* e = g_skel_ent;
        // Entity::cacheBoneMatrices refreshes the palette once per frame number (OgreEntity.cpp:1307-1330):
        // advance the frame counter so that every pose is really evaluated
        g_host->root->renderOneFrame();
        AnimationStateSet* set = e->getAllAnimationStates();
        bool first = true;
        for (auto& kv : set->getAnimationStates()) {
            kv.second->setEnabled(first);
            if (first) { kv.second->setLoop(true); kv.second->setTimePosition(time); }
            first = false;
        }
        e->addSoftwareAnimationRequest(true);     // keep normals: the FFP path skins them anyway
        e->_updateAnimation();
        const Affine3* bm = e->_getBoneMatrices();
        size_t nb = e->_getNumBoneMatrices();
        g_bones.resize(nb * 16);
        memcpy(g_bones.data(), bm, nb * 64);
        g_subs.clear();
        MeshPtr m = e->getMesh();
        if (m->sharedVertexData) {
            g_subs.emplace_back();
            dump_vertex_data(m->sharedVertexData, e->_getSkelAnimVertexData(), m->sharedBlendIndexToBoneIndexMap, g_subs.back());
        }
        for (size_t i = 0; i < m->getNumSubMeshes(); ++i) {
            SubMesh* sm = m->getSubMesh(i);
            if (sm->useSharedVertices) continue;
            g_subs.emplace_back();
            dump_vertex_data(sm->vertexData, e->getSubEntity(i)->_getSkelAnimVertexData(), sm->blendIndexToBoneIndexMap, g_subs.back());
        }
        return 0;
    } catch (std::exception& ex) { fprintf(stderr, "ref_skel_pose: %s\n", ex.what()); return 1; }
}

API size_t ref_skel_num_bones(void) { return g_bones.size() / 16; }
API const float* ref_skel_bones(void) { return g_bones.data(); }
API size_t ref_skel_num_subs(void) { return g_subs.size(); }
struct SkelSubInfo { uint32_t src_stride, norm_off, blend_stride, weight_off, index_off, nw, nverts, has_norm, dst_stride, nmap; size_t src_size, blend_size, skinned_size; };
API void ref_skel_sub_info(size_t i, SkelSubInfo* o)
{
    const SkelSub& s = g_subs[i];
    o->src_stride = s.src_stride; o->norm_off = s.norm_off; o->blend_stride = s.blend_stride; o->weight_off = s.weight_off;
    o->index_off = s.index_off; o->nw = s.nw; o->nverts = s.nverts; o->has_norm = s.has_norm; o->dst_stride = s.dst_stride;
    o->nmap = (uint32_t)s.bone_map.size(); o->src_size = s.src.size(); o->blend_size = s.blend.size(); o->skinned_size = s.skinned.size();
}
API const uint8_t* ref_skel_sub_src(size_t i) { return g_subs[i].src.data(); }
API const uint8_t* ref_skel_sub_blend(size_t i) { return g_subs[i].blend.data(); }
API const uint16_t* ref_skel_sub_map(size_t i) { return g_subs[i].bone_map.data(); }
API const uint8_t* ref_skel_sub_skinned(size_t i) { return g_subs[i].skinned.data(); }

/* ============================================ skeleton definition + palette evaluation ==== */
// Exports what Skeleton::setAnimationState / _getBoneMatrices (OgreMain/src/OgreSkeleton.cpp:150-217,402-424) consume,
// so that the oracle restatement and the device kernel can be checked against the real pipeline on the real data.

static Animation* skel_anim(uint32_t a)
{
    SkeletonInstance* sk = g_skel_ent->getSkeleton();
    return a < sk->getNumAnimations() ? sk->getAnimation((unsigned short)a) : nullptr;
}

API int ref_skel_def_sizes(uint32_t* nbones, uint32_t* nanims, uint32_t* blend_mode)
{
    if (!g_skel_ent || !g_skel_ent->hasSkeleton()) return 1;
    SkeletonInstance* sk = g_skel_ent->getSkeleton();
    *nbones = sk->getNumBones();
    *nanims = sk->getNumAnimations();
    *blend_mode = (uint32_t)sk->getBlendMode();
    return 0;
}

API void ref_skel_set_blend_mode(uint32_t mode) { g_skel_ent->getSkeleton()->setBlendMode((SkeletonAnimationBlendMode)mode); }

// parent[n] (-1 = root); bind[n][10] = initial position xyz, orientation wxyz, scale xyz; inherit[n][2] = orientation, scale;
// bindinv[n][10] = the Bone::setBindingPose products (inverse position xyz, inverse orientation wxyz, inverse scale xyz)
API void ref_skel_def_bones(int32_t* parent, float* bind, uint8_t* inherit, float* bindinv)
{
    SkeletonInstance* sk = g_skel_ent->getSkeleton();
    for (unsigned short h = 0; h < sk->getNumBones(); ++h) {
        Bone* b = sk->getBone(h);
        Node* p = b->getParent();
        parent[h] = p ? (int32_t) static_cast<Bone*>(p)->getHandle() : -1;
        const Vector3& ip = b->getInitialPosition(); const Quaternion& iq = b->getInitialOrientation(); const Vector3& is = b->getInitialScale();
        float* o = bind + 10 * h;
        o[0] = ip.x; o[1] = ip.y; o[2] = ip.z; o[3] = iq.w; o[4] = iq.x; o[5] = iq.y; o[6] = iq.z; o[7] = is.x; o[8] = is.y; o[9] = is.z;
        inherit[2 * h] = b->getInheritOrientation(); inherit[2 * h + 1] = b->getInheritScale();
        const Vector3& bp = b->_getBindingPoseInversePosition(); const Quaternion& bq = b->_getBindingPoseInverseOrientation();
        const Vector3& bs = b->_getBindingPoseInverseScale();
        o = bindinv + 10 * h;
        o[0] = bp.x; o[1] = bp.y; o[2] = bp.z; o[3] = bq.w; o[4] = bq.x; o[5] = bq.y; o[6] = bq.z; o[7] = bs.x; o[8] = bs.y; o[9] = bs.z;
    }
}

API int ref_skel_anim_info(uint32_t a, float* length, uint32_t* interp, uint32_t* rot_interp, uint32_t* ntracks, uint32_t* nkeys, char* name64)
{
    Animation* an = skel_anim(a);
    if (!an) return 1;
    *length = an->getLength(); *interp = (uint32_t)an->getInterpolationMode(); *rot_interp = (uint32_t)an->getRotationInterpolationMode();
    *ntracks = (uint32_t)an->_getNodeTrackList().size();
    uint32_t k = 0;
    for (auto& t : an->_getNodeTrackList()) k += t.second->getNumKeyFrames();
    *nkeys = k;
    snprintf(name64, 64, "%s", an->getName().c_str());
    return 0;
}

// Animation::setInterpolationMode / setRotationInterpolationMode on the real animation (0 = IM_LINEAR / RIM_LINEAR,
// 1 = IM_SPLINE / RIM_SPHERICAL, OgreAnimation.h:68-84)
API int ref_skel_set_interpolation(uint32_t a, uint32_t interp, uint32_t rot_interp)
{
    Animation* an = skel_anim(a);
    if (!an) return 1;
    an->setInterpolationMode((Animation::InterpolationMode)interp);
    an->setRotationInterpolationMode((Animation::RotationInterpolationMode)rot_interp);
    return 0;
}

// track_bone[ntracks], track_shortest[ntracks], key_offset[ntracks + 1], keys[nkeys][11] = time, translate xyz, rotation wxyz, scale xyz
API void ref_skel_anim_data(uint32_t a, int32_t* track_bone, uint8_t* track_shortest, uint32_t* key_offset, float* keys)
{
    Animation* an = skel_anim(a);
    uint32_t t = 0, k = 0;
    for (auto& kv : an->_getNodeTrackList()) {
        NodeAnimationTrack* tr = kv.second;
        track_bone[t] = kv.first; track_shortest[t] = tr->getUseShortestRotationPath(); key_offset[t] = k;
        for (unsigned short i = 0; i < tr->getNumKeyFrames(); ++i, ++k) {
            TransformKeyFrame* kf = tr->getNodeKeyFrame(i);
            float* o = keys + 11 * k;
            const Vector3& tv = kf->getTranslate(); const Quaternion& q = kf->getRotation(); const Vector3& sc = kf->getScale();
            o[0] = kf->getTime(); o[1] = tv.x; o[2] = tv.y; o[3] = tv.z; o[4] = q.w; o[5] = q.x; o[6] = q.y; o[7] = q.z; o[8] = sc.x; o[9] = sc.y; o[10] = sc.z;
        }
        ++t;
    }
    key_offset[t] = k;
}

// Enable exactly the given animation states (index into the skeleton's animation list, time position as
// AnimationState::getTimePosition would return it, weight), run SkeletonInstance::setAnimationState and
// _getBoneMatrices, copy the numBones x 16 palette out.
API int ref_skel_pose_states(uint32_t n, const uint32_t* anim, const float* time, const float* weight, float* bones_out)
{
    try {
        Entity* e = g_skel_ent;
        SkeletonInstance* sk = e->getSkeleton();
        AnimationStateSet* set = e->getAllAnimationStates();
        for (auto& kv : set->getAnimationStates()) kv.second->setEnabled(false);
        for (uint32_t i = 0; i < n; ++i) {
            Animation* an = skel_anim(anim[i]);
            if (!an) return 2;
            AnimationState* st = set->getAnimationState(an->getName());
            st->setLoop(false);                       // clamp, not wrap: the caller passes the final time position
            st->setEnabled(true);
            st->setTimePosition(time[i]);
            st->setWeight(weight[i]);
        }
        sk->setAnimationState(*set);
        std::vector<Affine3> m(sk->getNumBones());
        sk->_getBoneMatrices(m.data());
        memcpy(bones_out, m.data(), m.size() * 64);
        return 0;
    } catch (std::exception& ex) { fprintf(stderr, "ref_skel_pose_states: %s\n", ex.what()); return 1; }
}

// ... with AnimationState blend masks (mask: [n][numBones], NULL = none; AnimationState::createBlendMask / setBlendMaskData) and
// manually controlled bones (manual_handles[k]: Bone::setManuallyControlled(true) + position / orientation / scale from
// manual10[k]; Skeleton::reset leaves such bones alone, OgreSkeleton.cpp:233-240).  Everything is restored afterwards.
API int ref_skel_pose_states_ex(uint32_t n, const uint32_t* anim, const float* time, const float* weight, const float* mask,
                                uint32_t nmanual, const uint32_t* manual_handles, const float* manual10, float* bones_out)
{
    try {
        Entity* e = g_skel_ent;
        SkeletonInstance* sk = e->getSkeleton();
        const size_t nb = sk->getNumBones();
        AnimationStateSet* set = e->getAllAnimationStates();
        for (auto& kv : set->getAnimationStates()) { kv.second->setEnabled(false); if (kv.second->hasBlendMask()) kv.second->destroyBlendMask(); }
        std::vector<AnimationState*> used;
        for (uint32_t i = 0; i < n; ++i) {
            Animation* an = skel_anim(anim[i]);
            if (!an) return 2;
            AnimationState* st = set->getAnimationState(an->getName());
            st->setLoop(false);
            st->setEnabled(true);
            st->setTimePosition(time[i]);
            st->setWeight(weight[i]);
            if (mask) { st->createBlendMask(nb, 1.0f); st->_setBlendMaskData(mask + (size_t)i * nb); }
            used.push_back(st);
        }
        for (uint32_t k = 0; k < nmanual; ++k) {
            Bone* b = sk->getBone((unsigned short)manual_handles[k]);
            const float* x = manual10 + 10 * (size_t)k;
            b->setManuallyControlled(true);
            b->setPosition(Vector3(x[0], x[1], x[2]));
            b->setOrientation(Quaternion(x[3], x[4], x[5], x[6]));
            b->setScale(Vector3(x[7], x[8], x[9]));
        }
        sk->setAnimationState(*set);
        std::vector<Affine3> m(nb);
        sk->_getBoneMatrices(m.data());
        memcpy(bones_out, m.data(), m.size() * 64);
        for (uint32_t k = 0; k < nmanual; ++k) {
            Bone* b = sk->getBone((unsigned short)manual_handles[k]);
            b->setManuallyControlled(false);
            b->reset();
        }
        for (AnimationState* st : used) if (st->hasBlendMask()) st->destroyBlendMask();
        return 0;
    } catch (std::exception& ex) { fprintf(stderr, "ref_skel_pose_states_ex: %s\n", ex.what()); return 1; }
}

// CPU timing of the same: `ninst` palette evaluations (one enabled state each, time[i]); seconds for one pass, 1 thread
// (SkeletonInstance is not thread-safe and OGRE evaluates entities serially on the render thread).
API double ref_bench_palettes(uint32_t anim, const float* time, size_t ninst, float* bones_out)
{
    Entity* e = g_skel_ent;
    SkeletonInstance* sk = e->getSkeleton();
    AnimationStateSet* set = e->getAllAnimationStates();
    for (auto& kv : set->getAnimationStates()) kv.second->setEnabled(false);
    AnimationState* st = set->getAnimationState(skel_anim(anim)->getName());
    st->setLoop(false); st->setEnabled(true); st->setWeight(1.0f);
    const size_t nb = sk->getNumBones();
    std::vector<Affine3> m(nb);
    auto t0 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < ninst; ++i) {
        st->setTimePosition(time[i]);
        sk->setAnimationState(*set);
        sk->_getBoneMatrices(m.data());
        if (bones_out) memcpy(bones_out + i * nb * 16, m.data(), nb * 64);
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* ========================================= silhouette / shadow-volume index generation ==== */
// ShadowCaster::generateShadowVolume (OgreMain/src/OgreShadowCaster.cpp:191-515) driven on the edge list of the open
// mesh: the step after calculateLightFacing / extrudeVertices in the stencil-shadow chain.

struct RefCaster : public ShadowCaster {
    AxisAlignedBox box;
    bool getCastShadows() const override { return true; }
    const AxisAlignedBox& getWorldBoundingBox(bool) const override { return box; }
    Real getPointExtrusionDistance(const Light*) const override { return 0; }
    void run(EdgeData* ed, const HardwareIndexBufferSharedPtr& ib, size_t& used, const Light* l, ShadowRenderableList& list, unsigned long flags)
    {
        generateShadowVolume(ed, ib, used, l, list, flags);
    }
};

struct ShadowState {
    EdgeData* ed = nullptr;
    Light* light = nullptr;
    SceneNode* lnode = nullptr;
    std::vector<uint16_t> out;
    std::vector<uint32_t> ranges;          // per group: indexStart, indexCount, capStart, capCount
};
static ShadowState g_shadow;

API int ref_shadow_open(const char* mesh, uint32_t* ngroups, uint32_t* nedges, uint32_t* ntris)
{
    try {
        int rc = ref_scene_open(mesh, 64, 64, 100.0f);
        if (rc) return rc;
        MeshPtr m = g_scene.ent->getMesh();
        m->freeEdgeList();
        m->buildEdgeList();
        g_shadow = ShadowState();
        g_shadow.ed = m->getEdgeList(0);
        if (!g_shadow.ed) return 2;
        g_shadow.light = g_scene.sm->createLight("shadow_probe");
        g_shadow.lnode = g_scene.sm->getRootSceneNode()->createChildSceneNode();
        g_shadow.lnode->attachObject(g_shadow.light);
        uint32_t ne = 0;
        for (auto& eg : g_shadow.ed->edgeGroups) ne += (uint32_t)eg.edges.size();
        *ngroups = (uint32_t)g_shadow.ed->edgeGroups.size(); *nedges = ne; *ntris = (uint32_t)g_shadow.ed->triangles.size();
        return 0;
    } catch (std::exception& ex) { fprintf(stderr, "ref_shadow_open: %s\n", ex.what()); return 1; }
}

// groups[g] = {edge_start, edge_count, tri_start, tri_count, vertex_count}; edges[e] = {tri0, tri1, v0, v1, degenerate};
// tris[t] = vertIndex[3]
API void ref_shadow_edges(uint32_t* groups, uint32_t* edges, uint32_t* tris)
{
    EdgeData* ed = g_shadow.ed;
    uint32_t e = 0, g = 0;
    for (auto& eg : ed->edgeGroups) {
        uint32_t* o = groups + 5 * g++;
        o[0] = e; o[1] = (uint32_t)eg.edges.size(); o[2] = eg.triStart; o[3] = eg.triCount; o[4] = (uint32_t)eg.vertexData->vertexCount;
        for (auto& ed1 : eg.edges) {
            uint32_t* q = edges + 5 * e++;
            q[0] = ed1.triIndex[0]; q[1] = ed1.triIndex[1]; q[2] = ed1.vertIndex[0]; q[3] = ed1.vertIndex[1]; q[4] = ed1.degenerate ? 1u : 0u;
        }
    }
    for (size_t t = 0; t < ed->triangles.size(); ++t)
        for (int k = 0; k < 3; ++k) tris[3 * t + k] = ed->triangles[t].vertIndex[k];
}

// light4 = position (w=1, point light) or direction-to-light (w=0, directional) as Light::getAs4DVector gives it;
// bounds6 = the caster's light-cap bounds (min xyz, max xyz) for the McGuire test; `used_in` = indices already in the
// shared buffer.  Returns the number of indices written (from used_in on, or from 0 when the buffer wrapped), -1 on error;
// facings_out gets the per-triangle light facings the real EdgeData computed.
API long ref_shadow_volume(const float* light4, const float* bounds6, uint32_t flags, uint32_t separate_light_cap, uint32_t buffer_indices,
                           uint32_t used_in, int8_t* facings_out, uint32_t* used_out)
{
    try {
        EdgeData* ed = g_shadow.ed;
        Light* l = g_shadow.light;
        if (light4[3] == 0.0f) {
            l->setType(Light::LT_DIRECTIONAL);
            g_shadow.lnode->setPosition(0, 0, 0);
            g_shadow.lnode->setDirection(Vector3(-light4[0], -light4[1], -light4[2]), Node::TS_WORLD);
        } else {
            l->setType(Light::LT_POINT);
            g_shadow.lnode->resetOrientation();
            g_shadow.lnode->setPosition(light4[0], light4[1], light4[2]);
        }
        g_shadow.lnode->_update(true, true);
        const Vector4 lv = l->getAs4DVector();
        ed->updateTriangleLightFacing(lv);
        memcpy(facings_out, ed->triangleLightFacings.data(), ed->triangleLightFacings.size());
        HardwareIndexBufferSharedPtr ib = HardwareBufferManager::getSingleton().createIndexBuffer(HardwareIndexBuffer::IT_16BIT, buffer_indices,
                                                                                                  HBU_CPU_TO_GPU);
        {   // recognisable background so that untouched ranges show
            uint16_t* p = (uint16_t*)ib->lock(HardwareBuffer::HBL_DISCARD);
            for (uint32_t i = 0; i < buffer_indices; ++i) p[i] = 0xCDCD;
            ib->unlock();
        }
        RefCaster caster;
        caster.box.setExtents(bounds6[0], bounds6[1], bounds6[2], bounds6[3], bounds6[4], bounds6[5]);
        ShadowRenderableList list;
        for (auto& eg : ed->edgeGroups) list.push_back(OGRE_NEW ShadowRenderable(nullptr, ib, eg.vertexData, separate_light_cap != 0));
        size_t used = used_in;
        caster.run(ed, ib, used, l, list, flags);
        *used_out = (uint32_t)used;
        g_shadow.out.resize(buffer_indices);
        const uint16_t* p = (const uint16_t*)ib->lock(HardwareBuffer::HBL_READ_ONLY);
        memcpy(g_shadow.out.data(), p, buffer_indices * 2);
        ib->unlock();
        g_shadow.ranges.clear();
        for (auto* sr : list) {
            IndexData* id = sr->getRenderOperationForUpdate()->indexData;
            g_shadow.ranges.push_back((uint32_t)id->indexStart); g_shadow.ranges.push_back((uint32_t)id->indexCount);
            if (sr->isLightCapSeparate()) {
                IndexData* ic = sr->getLightCapRenderable()->getRenderOperationForUpdate()->indexData;
                g_shadow.ranges.push_back((uint32_t)ic->indexStart); g_shadow.ranges.push_back((uint32_t)ic->indexCount);
            } else { g_shadow.ranges.push_back(0); g_shadow.ranges.push_back(0); }
        }
        for (auto* sr : list) OGRE_DELETE sr;
        return (long)buffer_indices;
    } catch (std::exception& ex) { fprintf(stderr, "ref_shadow_volume: %s\n", ex.what()); return -1; }
}
API const uint16_t* ref_shadow_indices(void) { return g_shadow.out.data(); }
API const uint32_t* ref_shadow_ranges(void) { return g_shadow.ranges.data(); }

/* ================================================================ pose blending ==== */
// Mesh::softwareVertexPoseBlend (OgreMain/src/OgreMesh.cpp:1948-2007, public static) on a vertex buffer of `nverts`
// vertices, `floats_per_vertex` floats each (position at 0, normal at 3 when with_normals).  SURVEY §8f rank 3.
API int ref_pose_blend(float weight, uint32_t n_off, const uint32_t* off_idx, const float* off3, uint32_t n_nrm, const uint32_t* nrm_idx,
                       const float* nrm3, float* buf, uint32_t nverts, uint32_t floats_per_vertex, int with_normals)
{
    try {
        host_open(64, 64);
        std::map<uint32, Vector3f> offsets, normals;
        for (uint32_t i = 0; i < n_off; ++i) offsets[off_idx[i]] = Vector3f(off3[3 * i], off3[3 * i + 1], off3[3 * i + 2]);
        for (uint32_t i = 0; i < n_nrm; ++i) normals[nrm_idx[i]] = Vector3f(nrm3[3 * i], nrm3[3 * i + 1], nrm3[3 * i + 2]);
        VertexData vd;
        vd.vertexCount = nverts;
        vd.vertexDeclaration->addElement(0, 0, VET_FLOAT3, VES_POSITION);
        if (with_normals) vd.vertexDeclaration->addElement(0, 12, VET_FLOAT3, VES_NORMAL);
        auto vb = HardwareBufferManager::getSingleton().createVertexBuffer(floats_per_vertex * 4, nverts, HBU_CPU_TO_GPU);
        vb->writeData(0, (size_t)nverts * floats_per_vertex * 4, buf, true);
        vd.vertexBufferBinding->setBinding(0, vb);
        Mesh::softwareVertexPoseBlend(weight, offsets, normals, &vd);
        vb->readData(0, (size_t)nverts * floats_per_vertex * 4, buf);
        return 0;
    } catch (std::exception& ex) { fprintf(stderr, "ref_pose_blend: %s\n", ex.what()); return 1; }
}

// The whole pose-animation path of an Entity on a hand-built mesh (one sub-mesh, position + normal in one buffer):
// Entity::_updateAnimation -> initialisePoseVertexData (copy, normals zeroed when the poses carry normals,
// OgreEntity.cpp:1164-1205) -> Animation::apply -> VertexAnimationTrack::applyPoseToVertexData ->
// Mesh::softwareVertexPoseBlend per referenced pose (OgreAnimationTrack.cpp:925-944) -> Entity::finalisePoseNormals
// (OgreEntity.cpp:1207-1261).  pose p owns entries [pose_off[p], pose_off[p+1]) of idx / off3 / nrm3.
API int ref_pose_entity(uint32_t nverts, const float* posnorm6, uint32_t nposes, const uint32_t* pose_off, const uint32_t* idx, const float* off3,
                        const float* nrm3, const float* influence, int with_normals, float* out6)
{
    try {
        if (!g_scene.sm) { int rc = ref_scene_open("ogrehead.mesh", 64, 64, 100.0f); if (rc) return rc; }
        static int serial = 0;
        const String name = "pose_probe_" + StringConverter::toString(serial++);
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
            Pose* pose = m->createPose(1, "pose" + StringConverter::toString(p));      // target 1 = sub-mesh 0
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
        Entity* e = g_scene.sm->createEntity(m);
        g_scene.sm->getRootSceneNode()->createChildSceneNode()->attachObject(e);
        AnimationState* as = e->getAnimationState("manual");
        as->setEnabled(true);
        as->setTimePosition(0);
        e->addSoftwareAnimationRequest(with_normals != 0);
        g_host->root->renderOneFrame();          // advance the frame number: animation is evaluated once per frame
        e->_updateAnimation();
        VertexData* vd = e->getSubEntity(0)->_getSoftwareVertexAnimVertexData();
        auto ob = vd->vertexBufferBinding->getBuffer(vd->vertexDeclaration->findElementBySemantic(VES_POSITION)->getSource());
        if (ob->getVertexSize() != 24) { fprintf(stderr, "ref_pose_entity: unexpected vertex size %zu\n", ob->getVertexSize()); return 3; }
        ob->readData(0, (size_t)nverts * 24, out6);
        // The probe entity and mesh are deliberately kept alive: OGRE pools the temporary blend buffers by the ADDRESS of
        // their source buffer, and a freed probe buffer's address is reused by the next probe of another size.
        e->setVisible(false);
        return 0;
    } catch (std::exception& ex) { fprintf(stderr, "ref_pose_entity: %s\n", ex.what()); return 1; }
}

/* ==================================================================== timing ==== */

// `reps` calls of the real SSE softwareVertexSkinning over `ninst` instances (each with its own
// palette and destination), split over `threads` host threads; returns seconds for ONE pass
// over all instances (best of reps).
API double ref_bench_skin_instances(const float* sp, const float* sn, const float* bw, const uint8_t* bi, size_t sps, size_t sns,
                                    size_t bws, size_t bis, size_t nw, size_t nverts, const float* palettes, size_t nmat,
                                    size_t ninst, float* dst, size_t dst_inst_stride_floats, int threads, int reps)
{
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t)
            pool.emplace_back([=]() {
                std::vector<const float*> pal(nmat);
                for (size_t i = t; i < ninst; i += threads) {
                    for (size_t m = 0; m < nmat; ++m) pal[m] = palettes + (i * nmat + m) * 16;
                    float* d = dst + i * dst_inst_stride_floats;
                    g_sse->softwareVertexSkinning(sp, d, sn, sn ? d + 3 : nullptr, bw, bi, (const Affine3* const*)pal.data(), sps, 24,
                                                  sns, 24, bws, bis, nw, nverts);
                }
            });
        for (auto& th : pool) th.join();
        best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    return best;
}

// One sub-mesh of the config-2 step on the reference: per instance the real SSE
// softwareVertexMorph (two shared keyframes -> the instance's animated source, t = tarr[i]) and
// then the real SSE softwareVertexSkinning (instance source + own palette -> instance output),
// exactly the order of Entity::updateAnimation (OgreEntity.cpp:747-911).  pos+normal stride 24.
API double ref_bench_anim_sub(const float* key1, const float* key2, const float* tarr, const float* bw, const uint8_t* bi,
                              size_t bws, size_t bis, size_t nw, size_t nverts, const float* palettes, size_t pal_stride_floats,
                              const uint16_t* bone_map, size_t nmap, size_t ninst, float* tmp, float* dst, int threads, int do_morph)
{
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([=]() {
            std::vector<const float*> pal(nmap);
            for (size_t i = t; i < ninst; i += threads) {
                for (size_t m = 0; m < nmap; ++m) pal[m] = palettes + i * pal_stride_floats + (size_t)bone_map[m] * 16;
                float* src = tmp + i * nverts * 6;
                float* d = dst + i * nverts * 6;
                const float* s = key1;
                if (do_morph) { g_sse->softwareVertexMorph(tarr[i], key1, key2, src, 24, 24, 24, nverts, true); s = src; }
                g_sse->softwareVertexSkinning(s, d, s + 3, d + 3, bw, bi, (const Affine3* const*)pal.data(), 24, 24, 24, 24, bws, bis, nw, nverts);
            }
        });
    for (auto& th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

API double ref_bench_morph(float t, const float* a, const float* b, float* d, size_t stride, size_t n, int normals, int threads, int reps)
{
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        size_t per = (n / threads) & ~(size_t)3;
        for (int k = 0; k < threads; ++k)
            pool.emplace_back([=]() {
                size_t v0 = k * per, cnt = (k == threads - 1) ? n - v0 : per;
                size_t fo = v0 * stride / 4;
                g_sse->softwareVertexMorph(t, a + fo, b + fo, d + fo, stride, stride, stride, cnt, normals != 0);
            });
        for (auto& th : pool) th.join();
        best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    return best;
}

/* -------- the real Image::scale / PixelUtil::bulkPixelConversion on caller memory (no Root needed) -------- */

// Image::scale(src box of a `src_plane_w`-pixel-wide plane in format `src_fmt_name`, dst, FILTER_BILINEAR), the call
// TinyHardwarePixelBuffer::blitToMemory makes when the sizes differ (OgreTinyHardwarePixelBuffer.cpp:43-47).
// Pitches in pixels.  Returns 0, or 1 for an unknown format name.
API int ref_image_scale(const uint8_t* src, uint32_t sw, uint32_t sh, size_t src_pitch_px, const char* src_fmt_name, uint8_t* dst,
                        uint32_t dw, uint32_t dh, size_t dst_pitch_px, const char* dst_fmt_name)
{
    const PixelFormat sf = PixelUtil::getFormatFromName(src_fmt_name, true), df = PixelUtil::getFormatFromName(dst_fmt_name, true);
    if (sf == PF_UNKNOWN || df == PF_UNKNOWN) return 1;
    PixelBox ps(sw, sh, 1, sf, const_cast<uint8_t*>(src));
    ps.rowPitch = src_pitch_px; ps.slicePitch = src_pitch_px * sh;
    PixelBox pd(dw, dh, 1, df, dst);
    pd.rowPitch = dst_pitch_px; pd.slicePitch = dst_pitch_px * dh;
    Image::scale(ps, pd, Image::FILTER_BILINEAR);
    return 0;
}

API size_t ref_pixel_format_bytes(const char* fmt_name)
{
    const PixelFormat f = PixelUtil::getFormatFromName(fmt_name, true);
    return f == PF_UNKNOWN ? 0 : PixelUtil::getNumElemBytes(f);
}

API void ref_shutdown(void)
{
    if (g_host) { g_captured.clear(); g_tex_copies.clear(); g_host->root.reset(); delete g_host; g_host = nullptr; }
}
