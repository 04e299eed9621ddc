#pragma once
// Host-side cost accounting of the plugin's entry points, compiled in only with -DOGRECU_HOST_TIMING (tools/plugin_host_timing.sh):
// TSC ticks and call counts per bucket, read through ogrecu_plugin_host_timing().  The shipped plugin has none of it.
#include <stdint.h>
#ifdef OGRECU_HOST_TIMING
#include <x86intrin.h>
namespace Ogre {
enum { HT_RENDER, HT_SKIN, HT_OTHER_UTIL, HT_LOCK, HT_UNLOCK, HT_FLUSH, HT_END_FRAME, HT_COUNT };
struct HostTimer {
    static uint64_t acc[HT_COUNT], calls[HT_COUNT];
    int k; uint64_t t0;
    explicit HostTimer(int bucket) : k(bucket), t0(__rdtsc()) {}
    ~HostTimer() { acc[k] += __rdtsc() - t0; ++calls[k]; }
};
}
#define OGRECU_TIME(bucket) ::Ogre::HostTimer ogrecu_host_timer_(::Ogre::bucket)
#else
#define OGRECU_TIME(bucket) ((void)0)
#endif
