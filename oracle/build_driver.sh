#!/bin/sh
# Builds oracle/_ref/lib/libref_driver.so: the harness around the UNMODIFIED reference.
# Tiny's and OptimisedUtilGeneral's sources are compiled from where they lie under
# /root/reference (never copied); -fno-access-control lets the capture subclass read Tiny's
# private shader state.  Needs oracle/build_ref.py to have produced libOgreMain.so first.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${OGRE_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "build_driver: $REF absent; keeping prebuilt libref_driver.so"; exit 0; }
FLAGS="-O3 -DNDEBUG -msse -std=c++14 -fPIC -w -pthread -fno-access-control -fvisibility=hidden"
INC="-I$OUT/include -I$REF/OgreMain/include -I$REF/OgreMain/include/Threading -I$REF/OgreMain/src -I$REF/RenderSystems/Tiny/include -I$REF/RenderSystems/Tiny/src -I$HERE"
TINY="$REF/RenderSystems/Tiny/src/OgreTinyRenderSystem.cpp $REF/RenderSystems/Tiny/src/OgreTinyWindow.cpp $REF/RenderSystems/Tiny/src/OgreTinyTexture.cpp $REF/RenderSystems/Tiny/src/OgreTinyTextureManager.cpp $REF/RenderSystems/Tiny/src/OgreTinyHardwarePixelBuffer.cpp"
/usr/bin/g++ $FLAGS $INC -shared -o "$OUT/lib/libref_driver.so" "$HERE/ref_driver.cpp" $TINY \
    "$REF/OgreMain/src/OgreOptimisedUtilGeneral.cpp" -L"$OUT/lib" -lOgreMain -Wl,-rpath,'$ORIGIN'
echo "build_driver: ok"
