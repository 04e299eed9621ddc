// RenderSystems/CUDA — B200-native software-rendering RenderSystem for OGRE 14.6, a drop-in sibling
// of RenderSystems/Tiny.  All device work goes through the C ABI of libogrecu.so (include/ogrecu.h);
// there is no CPU rendering path in this plugin.
#ifndef __OgreCudaPrerequisites_H__
#define __OgreCudaPrerequisites_H__

#include "OgrePrerequisites.h"
#include "OgreException.h"
#include "ogrecu.h"

#define _OgreCudaExport __attribute__((visibility("default")))

namespace Ogre
{
class CudaRenderSystem;
class CudaTexture;
class CudaHardwareBuffer;

/// C-ABI status -> OGRE exception (OgreException.h:291); there is no fallback to fall back to
#define OGRECU_THROW_IF(expr)                                                                            \
    do {                                                                                                 \
        int rc__ = (expr);                                                                               \
        if (rc__ != 0)                                                                                   \
            OGRE_EXCEPT(Exception::ERR_RENDERINGAPI_ERROR, String(#expr ": ") + ogrecu_last_error_string()); \
    } while (0)
} // namespace Ogre
#endif
