// Mirrors RenderSystems/Tiny/include/OgreTinyDepthBuffer.h.  The f32 depth plane itself belongs to
// the render target's ogrecu_rt (each target owns one), so this object is the logical pool handle
// OgreMain's RenderSystem::setDepthBufferFor (OgreMain/src/OgreRenderSystem.cpp:512-552) shares out.
#ifndef __OgreCudaDepthBuffer_H__
#define __OgreCudaDepthBuffer_H__

#include "OgreCudaPrerequisites.h"
#include "OgreDepthBuffer.h"

namespace Ogre
{
class CudaDepthBuffer : public DepthBuffer
{
public:
    CudaDepthBuffer(uint16 poolId, uint32 width, uint32 height, uint32 fsaa, bool manual)
        : DepthBuffer(poolId, width, height, fsaa, manual)
    {
    }
};
} // namespace Ogre
#endif
