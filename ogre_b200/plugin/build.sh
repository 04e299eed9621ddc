#!/bin/sh
# Builds ogre_b200/plugin/RenderSystem_CUDA.so — the C++ drop-in RenderSystem + OptimisedUtil plugin —
# against the UNMODIFIED OgreMain headers in the reference tree and the stock libOgreMain.so that
# oracle/build_ref.py compiled from them.  Needs the reference tree (this container); on the GPU box
# the prebuilt .so travels with the repo snapshot.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${OGRE_REFERENCE:-/root/reference}"
OGRELIB="$ROOT/oracle/_ref/lib"
GEN="$ROOT/oracle/_ref/include"
[ -d "$REF" ] || { echo "plugin/build.sh: $REF absent; keeping prebuilt RenderSystem_CUDA.so"; exit 0; }
[ -f "$OGRELIB/libOgreMain.so" ] || { echo "plugin/build.sh: run oracle/build_ref.py first"; exit 1; }
# same arithmetic-defining flags as the reference build: -O3 -msse, no fast-math, no FMA contraction
FLAGS="-O3 -DNDEBUG -msse -std=c++14 -fPIC -Wall -Wno-unused-parameter -Wno-deprecated-declarations -pthread -fvisibility=hidden -fvisibility-inlines-hidden"
INC="-I$HERE/include -I$ROOT/include -I$GEN -I$REF/OgreMain/include -I$REF/OgreMain/include/Threading"
/usr/bin/g++ $FLAGS $INC -shared -o "$HERE/RenderSystem_CUDA.so" "$HERE"/src/*.cpp \
    -L"$OGRELIB" -lOgreMain -L"$ROOT/ogre_b200" -logrecu \
    -Wl,-rpath,'$ORIGIN/..' -Wl,-rpath,'$ORIGIN/../../oracle/_ref/lib' -Wl,--no-undefined
echo "plugin/build.sh: ok -> $HERE/RenderSystem_CUDA.so"
