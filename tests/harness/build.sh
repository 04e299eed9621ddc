#!/bin/sh
# Builds tests/harness/libplugin_harness.so (test infrastructure) against the reference's headers and
# the stock libOgreMain.so in oracle/_ref/lib.  Needs the reference tree; the .so travels to the GPU box.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${OGRE_REFERENCE:-/root/reference}"
[ -d "$REF" ] || { echo "harness/build.sh: $REF absent; keeping prebuilt harness"; exit 0; }
/usr/bin/g++ -O2 -DNDEBUG -msse -std=c++14 -fPIC -w -pthread -fvisibility=hidden \
    -I"$ROOT/oracle/_ref/include" -I"$REF/OgreMain/include" -I"$REF/OgreMain/include/Threading" \
    -I"$ROOT/ogre_b200/plugin/include" -I"$ROOT/include" \
    -shared -o "$HERE/libplugin_harness.so" "$HERE/plugin_harness.cpp" \
    -L"$ROOT/oracle/_ref/lib" -lOgreMain -L"$ROOT/ogre_b200/plugin" -l:RenderSystem_CUDA.so \
    -Wl,-rpath,'$ORIGIN/../../oracle/_ref/lib' -Wl,-rpath,'$ORIGIN/../../ogre_b200/plugin'
echo "harness/build.sh: ok"
