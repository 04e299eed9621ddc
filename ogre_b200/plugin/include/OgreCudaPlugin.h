// Mirrors RenderSystems/Tiny/include/OgreTinyPlugin.h.
#ifndef __CudaPlugin_H__
#define __CudaPlugin_H__

#include "OgreCudaPrerequisites.h"
#include "OgrePlugin.h"

namespace Ogre
{
class _OgreCudaExport CudaPlugin : public Plugin
{
public:
    CudaPlugin();
    const String& getName() const override;
    void install() override;
    void initialise() override;
    void shutdown() override;
    void uninstall() override;

protected:
    RenderSystem* mRenderSystem;
};
} // namespace Ogre
#endif
