#!/usr/bin/env python3
"""Build the UNMODIFIED reference (OGRE 14.6 OgreMain + RenderSystem_Tiny + Codec_STBI)
into oracle/_ref/ straight from the sources where they lie under /root/reference.

TEST INFRASTRUCTURE ONLY.  Nothing here is on the product path: the outputs are the
oracle the CUDA path is checked against (tests/, bench.py --impl reference /
cpu_baseline) and the stock host library the drop-in plugin links against.

The reference's own CMake build system is NOT run.  This script
  * writes the three tiny configuration headers CMake would have generated
    (OgreBuildSettings.h / OgreComponents.h / Ogre*Exports.h) into oracle/_ref/include,
    with the option values of a default Linux Release build
    (CMakeLists.txt:284-350, CMake/ConfigureBuild.cmake:35-77, CMake/Templates/*.in),
  * emits a build.ninja that runs g++ directly on OgreMain/src/*.cpp (the file set of
    OgreMain/CMakeLists.txt:22-150 for UNIX + zip), RenderSystems/Tiny/src/*.cpp and
    PlugIns/STBICodec/src/OgreSTBICodec.cpp with the flags that define the oracle's
    arithmetic: -O3 -DNDEBUG -msse -std=c++14, no -ffast-math, no FMA contraction
    (x86-64 baseline has no FMA),
  * copies the media the named configs need (ogrehead, Sinbad, jaiqua, their materials and
    textures) to oracle/_ref/media so the GPU box (which has no /root/reference) can run
    the same scenes.

No reference source is copied into the repository; oracle/_ref/ is git-ignored.
"""
import glob
import os
import shutil
import subprocess
import sys

REF = os.environ.get("OGRE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
INC = os.path.join(OUT, "include")
LIB = os.path.join(OUT, "lib")
OBJ = os.path.join(OUT, "build")
CXX = "/usr/bin/g++"
CC = "/usr/bin/gcc"

BUILD_SETTINGS = """\
#ifndef __Custom_Config_H_
#define __Custom_Config_H_
// written by oracle/build_ref.py: default Linux Release option values
#define OGRE_VERSION_MAJOR 14
#define OGRE_VERSION_MINOR 6
#define OGRE_VERSION_PATCH 0
#define OGRE_VERSION_SUFFIX ""
#define OGRE_VERSION_NAME "Tsathoggua"
#define OGRE_DEBUG_MODE 0
#define OGRE_RESOURCEMANAGER_STRICT 2
#define OGRE_CONFIG_LITTLE_ENDIAN
#define OGRE_DOUBLE_PRECISION 0
#define OGRE_NODE_INHERIT_TRANSFORM 0
#define OGRE_ASSERT_MODE 2
#define OGRE_THREAD_SUPPORT 3
#define OGRE_THREAD_PROVIDER 4
#define OGRE_NO_MESHLOD 0
#define OGRE_NO_DDS_CODEC 1
#define OGRE_NO_PVRTC_CODEC 1
#define OGRE_NO_ETC_CODEC 1
#define OGRE_NO_ASTC_CODEC 1
#define OGRE_NO_ZIP_ARCHIVE 0
#define OGRE_NO_TBB_SCHEDULER 1
#define OGRE_PROFILING 0
#define OGRE_NO_QUAD_BUFFER_STEREO 1
#define OGRE_NO_LOCALE_STRCONVERT 0
#endif
"""

COMPONENTS = """\
#ifndef __OGRE_COMPONENTS_Config_H_
#define __OGRE_COMPONENTS_Config_H_
#define OGRE_BUILD_RENDERSYSTEM_TINY
#define OGRE_BUILD_PLUGIN_STBI
#define OGRE_BITES_HAVE_SDL 0
#endif
"""


def export_header(guard, export, private=None, exports_define=None):
    s = "#ifndef %s\n#define %s\n" % (guard, guard)
    s += '#define %s __attribute__((visibility("default")))\n' % export
    if private:
        s += '#define %s __attribute__((visibility("hidden")))\n' % private
    if export == "_OgreExport":
        s += "#define OGRE_DEPRECATED __attribute__((__deprecated__))\n"
        s += "#define OGRE_DEPRECATED_EXPORT _OgreExport OGRE_DEPRECATED\n"
        s += "#define OGRE_DEPRECATED_NO_EXPORT _OgrePrivate OGRE_DEPRECATED\n"
    s += "#endif\n"
    return s


def write_if_changed(path, text):
    if os.path.exists(path) and open(path).read() == text:
        return
    with open(path, "w") as f:
        f.write(text)


def main():
    if not os.path.isdir(REF):
        print("build_ref: %s not present; keeping prebuilt oracle/_ref as is" % REF)
        return 0
    for d in (INC, LIB, OBJ):
        os.makedirs(d, exist_ok=True)
    write_if_changed(os.path.join(INC, "OgreBuildSettings.h"), BUILD_SETTINGS)
    write_if_changed(os.path.join(INC, "OgreComponents.h"), COMPONENTS)
    write_if_changed(os.path.join(INC, "OgreExports.h"),
                     export_header("OGRE_EXPORTS_H", "_OgreExport", "_OgrePrivate"))
    write_if_changed(os.path.join(INC, "OgreTinyExports.h"),
                     export_header("OGRE_TINY_EXPORTS_H", "_OgreTinyExport"))
    write_if_changed(os.path.join(INC, "OgreSTBICodecExports.h"),
                     export_header("OGRE_STBI_EXPORTS_H", "_OgreSTBICodecExport"))

    main_src = sorted(glob.glob(os.path.join(REF, "OgreMain/src/*.cpp")))
    drop = {"OgreFileSystemLayerNoOp.cpp", "OgreDDSCodec.cpp", "OgrePVRTCCodec.cpp",
            "OgreETCCodec.cpp", "OgreASTCCodec.cpp"}
    main_src = [s for s in main_src if os.path.basename(s) not in drop]
    main_src.append(os.path.join(REF, "OgreMain/src/GLX/OgreFileSystemLayer.cpp"))
    zip_c = os.path.join(REF, "OgreMain/src/zip/zip.c")
    tiny_src = sorted(glob.glob(os.path.join(REF, "RenderSystems/Tiny/src/*.cpp")))
    stbi_src = [os.path.join(REF, "PlugIns/STBICodec/src/OgreSTBICodec.cpp")]

    common = ("-O3 -DNDEBUG -msse -std=c++14 -fPIC -fvisibility=hidden "
              "-fvisibility-inlines-hidden -w -pthread")
    inc_main = " ".join("-I" + p for p in [
        INC, REF + "/OgreMain/include", REF + "/OgreMain/include/Threading",
        REF + "/OgreMain/src", REF + "/OgreMain/src/zip"])
    n = []
    n.append("cxx = %s\ncc = %s\n" % (CXX, CC))
    n.append("rule cxx\n  command = $cxx $flags -MMD -MF $out.d -c $in -o $out\n"
             "  depfile = $out.d\n  deps = gcc\n  description = CXX $out\n")
    n.append("rule cc\n  command = $cc $flags -c $in -o $out\n  description = CC $out\n")
    n.append("rule link\n  command = $cxx -shared -o $out $in $libs\n  description = LINK $out\n")

    def objs(srcs, tag, flags):
        out = []
        for s in srcs:
            o = os.path.join(OBJ, tag, os.path.basename(s) + ".o")
            n.append("build %s: cxx %s\n  flags = %s\n" % (o, s, flags))
            out.append(o)
        return out

    o_main = objs(main_src, "OgreMain", "%s -DOgreMain_EXPORTS %s" % (common, inc_main))
    zo = os.path.join(OBJ, "OgreMain", "zip.c.o")
    n.append("build %s: cc %s\n  flags = -O2 -fPIC -w -DMINIZ_NO_STDIO -I%s/OgreMain/src/zip\n"
             % (zo, zip_c, REF))
    o_main.append(zo)
    libmain = os.path.join(LIB, "libOgreMain.so")
    n.append("build %s: link %s\n  libs = -pthread -ldl -Wl,-soname,libOgreMain.so\n"
             % (libmain, " ".join(o_main)))

    rpath = "-Wl,-rpath,'$$ORIGIN' -L%s -lOgreMain" % LIB
    o_tiny = objs(tiny_src, "Tiny", "%s -DRenderSystem_Tiny_EXPORTS %s -I%s/RenderSystems/Tiny/include"
                  % (common, inc_main, REF))
    n.append("build %s: link %s | %s\n  libs = %s\n"
             % (os.path.join(LIB, "RenderSystem_Tiny.so"), " ".join(o_tiny), libmain, rpath))
    o_stbi = objs(stbi_src, "STBI", "%s -DCodec_STBI_EXPORTS %s -I%s/PlugIns/STBICodec/include"
                  % (common, inc_main, REF))
    n.append("build %s: link %s | %s\n  libs = %s\n"
             % (os.path.join(LIB, "Codec_STBI.so"), " ".join(o_stbi), libmain, rpath))
    write_if_changed(os.path.join(OBJ, "build.ninja"), "\n".join(n))

    r = subprocess.run(["ninja", "-C", OBJ, "-j", str(os.cpu_count() or 4)])
    if r.returncode:
        return r.returncode

    # media for the named configs (input data, not source): ogrehead / Sinbad / jaiqua
    media = os.path.join(OUT, "media")
    sm = os.path.join(REF, "Samples/Media")
    for sub in ("models", "scripts", "textures", "packs"):
        os.makedirs(os.path.join(media, sub), exist_ok=True)
    for f in ("ogrehead.mesh", "jaiqua.mesh", "jaiqua.skeleton"):
        shutil.copy2(os.path.join(sm, "models", f), os.path.join(media, "models", f))
    shutil.copy2(os.path.join(sm, "packs/Sinbad.zip"), os.path.join(media, "packs/Sinbad.zip"))
    for f in ("Ogre.material", "jaiqua.material"):
        p = os.path.join(sm, "materials/scripts", f)
        if os.path.exists(p):
            shutil.copy2(p, os.path.join(media, "scripts", f))
    # jaiqua's material lives in the 1 700-line Examples.material (full of unrelated dependencies);
    # the test scenes only need its one texture unit.  Without a material Tiny falls back to
    # BaseWhite and then reads its never-initialised DefaultShader::image pointer
    # (TinyRenderSystem ctor, OgreTinyRenderSystem.cpp:27-46) — undefined behaviour we do not test.
    with open(os.path.join(media, "scripts", "jaiqua.material"), "w") as f:
        f.write("material jaiqua\n{\n\ttechnique\n\t{\n\t\tpass\n\t\t{\n\t\t\ttexture_unit\n\t\t\t{\n"
                "\t\t\t\ttexture blue_jaiqua.jpg\n\t\t\t\ttex_address_mode clamp\n\t\t\t}\n\t\t}\n\t}\n}\n")
    for f in ("GreenSkin.jpg", "spheremap.png", "tusk.jpg", "WeirdEye.png", "dirt01.jpg",
              "blue_jaiqua.jpg"):
        p = os.path.join(sm, "materials/textures", f)
        if os.path.exists(p):
            shutil.copy2(p, os.path.join(media, "textures", f))
    print("build_ref: ok ->", LIB)
    return 0


if __name__ == "__main__":
    sys.exit(main())
