// Plugin entry (mirrors RenderSystems/Tiny/src/OgreTinyPlugin.cpp:13-61): dllStartPlugin creates the
// plugin, install() registers the render system and swaps in the CUDA OptimisedUtil.
#include "OgreCudaPlugin.h"
#include "OgreCudaOptimisedUtil.h"
#include "OgreCudaRenderSystem.h"
#include "OgreRoot.h"

namespace Ogre
{
static const String sPluginName = "CUDA RenderSystem";

CudaPlugin::CudaPlugin() : mRenderSystem(0) {}

const String& CudaPlugin::getName() const { return sPluginName; }

void CudaPlugin::install()
{
    mRenderSystem = OGRE_NEW CudaRenderSystem();
    Root::getSingleton().addRenderSystem(mRenderSystem);
    // OGRECU_OPTIMISED_UTIL=0 keeps the stock (SSE) implementation, e.g. to A/B the two in one process
    const char* env = getenv("OGRECU_OPTIMISED_UTIL");
    if (!env || atoi(env) != 0) CudaOptimisedUtil::install();
}

void CudaPlugin::initialise() {}
void CudaPlugin::shutdown() {}

void CudaPlugin::uninstall()
{
    CudaOptimisedUtil::uninstall();
    OGRE_DELETE mRenderSystem;
    mRenderSystem = 0;
}

static CudaPlugin* plugin;

extern "C" void _OgreCudaExport dllStartPlugin(void);
extern "C" void _OgreCudaExport dllStopPlugin(void);

extern "C" void _OgreCudaExport dllStartPlugin(void)
{
    plugin = OGRE_NEW CudaPlugin();
    Root::getSingleton().installPlugin(plugin);
}

extern "C" void _OgreCudaExport dllStopPlugin(void)
{
    Root::getSingleton().uninstallPlugin(plugin);
    OGRE_DELETE plugin;
}
} // namespace Ogre
