#!/bin/sh
# Where a crowd frame's HOST time goes inside the plugin: builds an instrumented copy of RenderSystem_CUDA.so
# (-DOGRECU_HOST_TIMING, build/variants/timing/) and runs the Entity crowd through it.
#   sh tools/plugin_host_timing.sh [entities] [WxH]        (build here, run on the GPU box: the .so travels)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$ROOT/build/variants/timing"
mkdir -p "$OUT"
REF="${OGRE_REFERENCE:-/root/reference}"
if [ -d "$REF" ]; then
    FLAGS="-O3 -DNDEBUG -DOGRECU_HOST_TIMING -msse -std=c++14 -fPIC -pthread -fvisibility=hidden -fvisibility-inlines-hidden -Wno-deprecated-declarations"
    INC="-I$ROOT/ogre_b200/plugin/include -I$ROOT/include -I$ROOT/oracle/_ref/include -I$REF/OgreMain/include -I$REF/OgreMain/include/Threading"
    /usr/bin/g++ $FLAGS $INC -shared -o "$OUT/RenderSystem_CUDA.so" "$ROOT"/ogre_b200/plugin/src/*.cpp \
        -L"$ROOT/oracle/_ref/lib" -lOgreMain -L"$ROOT/ogre_b200" -logrecu \
        -Wl,-rpath,'$ORIGIN/..' -Wl,-rpath,'$ORIGIN/../../oracle/_ref/lib' -Wl,--no-undefined
fi
[ "$1" = "build" ] && exit 0
# the test harness links RenderSystem_CUDA.so by name, so the instrumented copy has to stand in its place: only done on a gpurun
# box, whose tree is a scratch snapshot
[ -n "$GRAFT_REPO_ROOT" ] || { echo "run this under gpurun (it replaces ogre_b200/plugin/RenderSystem_CUDA.so in the box's snapshot)"; exit 1; }
cp "$OUT/RenderSystem_CUDA.so" "$ROOT/ogre_b200/plugin/RenderSystem_CUDA.so"
python "$ROOT/tests/harness/run_crowd.py" --plugin cuda --count "${1:-1000}" --size "${2:-1280x720}" --frames 10 --host-timing 1
