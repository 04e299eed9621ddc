#!/usr/bin/env python3
"""A crowd of animated entities through a render-system plugin in THIS process (one Ogre::Root per process): N entities of
one skeletal mesh, each re-skinned every frame through Root::renderOneFrame.  TEST / BENCH INFRASTRUCTURE.

  python tests/harness/run_crowd.py --plugin cuda|tiny --mesh Sinbad.mesh --count 1000 --size 640x360 [--frames 5]
         [--batch 0|1] [--out x.npz]
Prints one JSON line: seconds per frame, skinned Mvert/s; --out also saves the frame rendered at a fixed pose.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFLIB = os.path.join(ROOT, "oracle", "_ref", "lib")
MEDIA = os.path.join(ROOT, "oracle", "_ref", "media")
PLUGINS = {
    "cuda": (os.environ.get("OGRECU_PLUGIN", os.path.join(ROOT, "ogre_b200", "plugin", "RenderSystem_CUDA.so")), "CUDA"),
    "tiny": (os.path.join(REFLIB, "RenderSystem_Tiny.so"), "Tiny"),
}


def tsc_ghz():
    """TSC ticks per nanosecond, measured against the wall clock for 0.2 s"""
    import time
    try:
        src = "#include <x86intrin.h>\nunsigned long long t(void){return __rdtsc();}"
        import subprocess, tempfile
        d = tempfile.mkdtemp()
        open(d + "/t.c", "w").write(src)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", d + "/t.so", d + "/t.c"])
        t = C.CDLL(d + "/t.so"); t.t.restype = C.c_uint64
        a0, w0 = t.t(), time.perf_counter()
        time.sleep(0.2)
        return (t.t() - a0) / (time.perf_counter() - w0) / 1e9
    except Exception:
        return 2.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plugin", required=True, choices=sorted(PLUGINS))
    ap.add_argument("--mesh", default="Sinbad.mesh")
    ap.add_argument("--count", type=int, default=100)
    ap.add_argument("--spacing", type=float, default=12.0)
    ap.add_argument("--size", default="640x360")
    ap.add_argument("--frames", type=int, default=5)
    ap.add_argument("--batch", type=int, default=1, help="CUDA plugin: its 'OptimisedUtil Batching' option")
    ap.add_argument("--out", default="")
    ap.add_argument("--host-timing", type=int, default=0, help="plugin built by tools/plugin_host_timing.sh: print its per-bucket host time")
    a = ap.parse_args()
    w, h = [int(v) for v in a.size.split("x")]
    so, name = PLUGINS[a.plugin]
    L = C.CDLL(os.path.join(HERE, "libplugin_harness.so"), mode=C.RTLD_GLOBAL)
    L.ph_error.restype = C.c_char_p
    L.ph_crowd_open.argtypes = [C.c_char_p, C.c_uint32, C.c_float]
    L.ph_crowd_frame_time.restype = C.c_double
    L.ph_crowd_frame_time.argtypes = [C.c_int, C.c_float]
    L.ph_crowd_vertices.restype = C.c_size_t
    L.ph_crowd_render_at.argtypes = [C.c_float, C.c_void_p]
    L.ph_set_rs_option.argtypes = [C.c_char_p, C.c_char_p]
    L.ph_triangles.restype = C.c_size_t
    L.ph_batches.restype = C.c_size_t

    def ok(rc, what):
        if rc:
            raise SystemExit("%s failed (%d): %s" % (what, rc, L.ph_error().decode()))

    ok(L.ph_open(so.encode(), name.encode(), os.path.join(REFLIB, "Codec_STBI.so").encode(), MEDIA.encode(), w, h), "ph_open")
    if a.plugin == "cuda":
        ok(L.ph_set_rs_option(b"OptimisedUtil Batching", b"Yes" if a.batch else b"No"), "ph_set_rs_option")
    ok(L.ph_crowd_open(a.mesh.encode(), a.count, a.spacing), "ph_crowd_open")
    nverts = L.ph_crowd_vertices()
    timing = None
    if a.host_timing:
        P = C.CDLL(so)
        acc, calls = (C.c_uint64 * 8)(), (C.c_uint64 * 8)()
        L.ph_crowd_frame_time(2, 0.033)
        P.ogrecu_plugin_host_timing(acc, calls)
    dt = L.ph_crowd_frame_time(a.frames, 0.033)
    if a.host_timing:
        P.ogrecu_plugin_host_timing(acc, calls)
        # TSC rate from the wall clock of the same interval (frames + the 2 warm-up frames ph_crowd_frame_time runs first)
        names = ["_render", "softwareVertexSkinning", "other OptimisedUtil", "lock", "unlock", "flushPendingDraws", "_endFrame"]
        ghz = float(os.environ.get("OGRECU_TSC_GHZ", "0")) or tsc_ghz()
        nf = a.frames + 2
        timing = {n: {"ms_per_frame": acc[i] / ghz / 1e6 / nf, "calls_per_frame": calls[i] / nf} for i, n in enumerate(names)}
    frame = np.zeros((h, w, 3), np.uint8)
    ok(L.ph_crowd_render_at(1.234, frame.ctypes.data), "ph_crowd_render_at")
    line = {"plugin": a.plugin, "batch": a.batch if a.plugin == "cuda" else None, "mesh": a.mesh, "entities": a.count, "size": a.size,
            "vertices_per_frame": int(nverts), "s_per_frame": dt, "Mvert_per_s": nverts / dt / 1e6, "triangles": int(L.ph_triangles()),
            "batches": int(L.ph_batches()), "covered_fraction": float((frame != np.array([25, 51, 76], np.uint8)).any(axis=2).mean())}
    if timing:
        line["plugin_host_ms_per_frame"] = timing
    print(json.dumps(line))
    if a.out:
        np.savez_compressed(a.out, frame=frame)
    L.ph_close()


if __name__ == "__main__":
    main()
