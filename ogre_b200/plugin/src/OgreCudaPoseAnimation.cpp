#include "OgreCudaPoseAnimation.h"
#include "OgreCudaHardwareBufferManager.h"
#include "OgreCudaRenderSystem.h"
#include "OgrePose.h"

#include <vector>

namespace Ogre
{
namespace
{
struct Element { CudaHardwareBuffer* buf; size_t offset; uint32 stride; size_t bytes; };

// first float of `sem` of vertex vertexStart in the device copy of its buffer
Element element(const VertexData* vd, VertexElementSemantic sem)
{
    Element e = {0, 0, 0, 0};
    const VertexElement* el = vd->vertexDeclaration->findElementBySemantic(sem);
    if (!el) return e;
    auto buf = vd->vertexBufferBinding->getBuffer(el->getSource());
    e.buf = dynamic_cast<CudaHardwareBuffer*>(buf->_getImpl<HardwareBuffer>());
    OgreAssert(e.buf, "vertex buffer was not created by CudaHardwareBufferManager");
    e.stride = (uint32)buf->getVertexSize();
    e.offset = el->getOffset() + vd->vertexStart * e.stride;
    e.bytes = vd->vertexCount ? (vd->vertexCount - 1) * e.stride + 12 : 0;
    return e;
}

float* devicePtr(const Element& e) { return (float*)const_cast<void*>(e.buf->getDevicePtr(e.offset)); }

void upload(ogrecu_ctx* ctx, const void* host, size_t bytes, void** dev)
{
    *dev = 0;
    if (!bytes) return;
    OGRECU_THROW_IF(ogrecu_malloc(ctx, bytes, dev));
    OGRECU_THROW_IF(ogrecu_memcpy_h2d(ctx, *dev, host, bytes));
}
} // namespace

CudaPoseAnimation::CudaPoseAnimation(const Mesh* mesh) : mCtx(0)
{
    CudaRenderSystem* rs = CudaRenderSystem::getInstance();
    OgreAssert(rs, "CUDA render system is not installed");
    mCtx = rs->getContext();
    for (const Pose* pose : mesh->getPoseList())
    {
        // the std::maps flattened in key order: the order Mesh::softwareVertexPoseBlend walks them (OgreMesh.cpp:1967,1989)
        std::vector<uint32> oi, ni;
        std::vector<float> ov, nv;
        for (auto& kv : pose->getVertexOffsets())
        {
            oi.push_back((uint32)kv.first);
            ov.insert(ov.end(), {(float)kv.second.x, (float)kv.second.y, (float)kv.second.z});
        }
        for (auto& kv : pose->getNormals())
        {
            ni.push_back((uint32)kv.first);
            nv.insert(nv.end(), {(float)kv.second.x, (float)kv.second.y, (float)kv.second.z});
        }
        DevicePose d = {(uint32)oi.size(), (uint32)ni.size(), 0, 0, 0, 0};
        upload(mCtx, oi.data(), oi.size() * 4, &d.offsetIndex);
        upload(mCtx, ov.data(), ov.size() * 4, &d.offsets);
        upload(mCtx, ni.data(), ni.size() * 4, &d.normalIndex);
        upload(mCtx, nv.data(), nv.size() * 4, &d.normals);
        mPoses.push_back(d);
    }
    OGRECU_THROW_IF(ogrecu_sync(mCtx));   // the host vectors above go out of scope
}

CudaPoseAnimation::~CudaPoseAnimation()
{
    for (DevicePose& d : mPoses)
        for (void* p : {d.offsetIndex, d.offsets, d.normalIndex, d.normals})
            if (p) ogrecu_free(mCtx, p);
}

void CudaPoseAnimation::initialise(const VertexData* src, VertexData* dest, bool animateNormals)
{
    CudaRenderSystem::flushPendingDraws();   // recorded draws may still read dest
    const VertexElement* se = src->vertexDeclaration->findElementBySemantic(VES_POSITION);
    const VertexElement* de = dest->vertexDeclaration->findElementBySemantic(VES_POSITION);
    auto sbuf = src->vertexBufferBinding->getBuffer(se->getSource());
    auto dbuf = dest->vertexBufferBinding->getBuffer(de->getSource());
    // destBuffer->copyData(*origBuffer, 0, 0, size, true) (OgreEntity.cpp:1179): device to device for our buffers
    dbuf->copyData(*sbuf, 0, 0, dbuf->getSizeInBytes(), true);
    if (!animateNormals) return;
    Element n = element(dest, VES_NORMAL);
    if (!n.buf || !dest->vertexCount) return;
    OGRECU_THROW_IF(ogrecu_pose_normals_reset_dev(mCtx, devicePtr(n), (uint32)dest->vertexCount, n.stride, 1, 0));
    OGRECU_THROW_IF(ogrecu_buf_mark_device_dirty(n.buf->getHandle(), n.offset, n.bytes));
}

void CudaPoseAnimation::applyPose(ushort poseIndex, VertexData* dest, float influence)
{
    OgreAssert(poseIndex < mPoses.size(), "pose index out of range");
    if (influence == 0.0f) return;            // Mesh::softwareVertexPoseBlend: "Do nothing if no weight" (OgreMesh.cpp:1952-1954)
    const DevicePose& d = mPoses[poseIndex];
    // Addressing as Mesh::softwareVertexPoseBlend has it (OgreMesh.cpp:1958-2005): the position buffer is locked whole and
    // vertex i's position sits at pBase + i * vertexSize — neither the position element's offset nor vertexStart enter;
    // normals are blended only when they live in the SAME buffer as the positions (:1964), at pBase + normal offset.
    const VertexElement* posElem = dest->vertexDeclaration->findElementBySemantic(VES_POSITION);
    const VertexElement* normElem = dest->vertexDeclaration->findElementBySemantic(VES_NORMAL);
    OgreAssert(posElem, "no position element");
    auto buf = dest->vertexBufferBinding->getBuffer(posElem->getSource());
    CudaHardwareBuffer* hb = dynamic_cast<CudaHardwareBuffer*>(buf->_getImpl<HardwareBuffer>());
    OgreAssert(hb, "vertex buffer was not created by CudaHardwareBufferManager");
    const uint32 stride = (uint32)buf->getVertexSize();
    if (d.numOffsets)
    {
        OGRECU_THROW_IF(ogrecu_pose_blend_dev(mCtx, influence, (const uint32_t*)d.offsetIndex, (const float*)d.offsets, d.numOffsets,
                                              (float*)const_cast<void*>(hb->getDevicePtr(0)), stride));
        OGRECU_THROW_IF(ogrecu_buf_mark_device_dirty(hb->getHandle(), 0, buf->getSizeInBytes()));
    }
    const bool normals = normElem && d.numNormals && posElem->getSource() == normElem->getSource();
    if (normals)
    {
        OGRECU_THROW_IF(ogrecu_pose_blend_dev(mCtx, influence, (const uint32_t*)d.normalIndex, (const float*)d.normals, d.numNormals,
                                              (float*)const_cast<void*>(hb->getDevicePtr(normElem->getOffset())), stride));
        OGRECU_THROW_IF(ogrecu_buf_mark_device_dirty(hb->getHandle(), 0, buf->getSizeInBytes()));
    }
}

void CudaPoseAnimation::applyKeyFrames(const VertexPoseKeyFrame* kf1, const VertexPoseKeyFrame* kf2, float t, float weight,
                                       VertexData* dest)
{
    const VertexPoseKeyFrame::PoseRefList& l1 = kf1->getPoseReferences();
    const VertexPoseKeyFrame::PoseRefList& l2 = kf2->getPoseReferences();
    for (auto& p1 : l1)
    {
        float endInfluence = 0;
        for (auto& p2 : l2)
            if (p1.poseIndex == p2.poseIndex) { endInfluence = p2.influence; break; }
        float influence = Math::lerp(p1.influence, endInfluence, t);
        influence = weight * influence;
        applyPose(p1.poseIndex, dest, influence);
    }
    for (auto& p2 : l2)
    {
        bool found = false;
        for (auto& p1 : l1)
            if (p1.poseIndex == p2.poseIndex) { found = true; break; }
        if (found) continue;
        float influence = t * p2.influence;
        influence = weight * influence;
        applyPose(p2.poseIndex, dest, influence);
    }
}

void CudaPoseAnimation::finalise(const VertexData* src, VertexData* dest)
{
    Element dn = element(dest, VES_NORMAL), sn = element(src, VES_NORMAL);
    if (!dn.buf || !sn.buf || !dest->vertexCount) return;
    // the reference advances the source pointer by the DESTINATION's vertex size (OgreEntity.cpp:1257); the two are equal for
    // the clones Entity makes — required here
    OgreAssert(sn.stride == dn.stride || sn.buf != dn.buf, "source and destination strides differ");
    OGRECU_THROW_IF(ogrecu_pose_normals_finalise_dev(mCtx, devicePtr(dn), devicePtr(sn), (uint32)dest->vertexCount, dn.stride, sn.stride, 1, 0));
    OGRECU_THROW_IF(ogrecu_buf_mark_device_dirty(dn.buf->getHandle(), dn.offset, dn.bytes));
}
} // namespace Ogre
