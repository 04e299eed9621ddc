#!/usr/bin/env python3
"""Run one OGRE scene through a render-system plugin in THIS process (one Ogre::Root per process)
and save what came back.  TEST INFRASTRUCTURE: used by tests/test_plugin_*.py via subprocess.

  python tests/harness/run_scene.py --plugin cuda|tiny --mesh ogrehead.mesh --size 320x180 --out x.npz
         [--cam-z 140] [--rtt] [--quad] [--skel 0.3,1.7] [--time N]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFLIB = os.path.join(ROOT, "oracle", "_ref", "lib")
MEDIA = os.path.join(ROOT, "oracle", "_ref", "media")
PLUGINS = {
    "cuda": (os.path.join(ROOT, "ogre_b200", "plugin", "RenderSystem_CUDA.so"), "CUDA"),
    "tiny": (os.path.join(REFLIB, "RenderSystem_Tiny.so"), "Tiny"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plugin", required=True, choices=sorted(PLUGINS))
    ap.add_argument("--mesh", default="ogrehead.mesh")
    ap.add_argument("--size", default="320x180")
    ap.add_argument("--cam-z", type=float, default=140.0)
    ap.add_argument("--out", required=True)
    ap.add_argument("--rtt", action="store_true")
    ap.add_argument("--quad", action="store_true")
    ap.add_argument("--skel", default="")
    ap.add_argument("--time", type=int, default=0)
    ap.add_argument("--order", type=int, default=-1, help="pin the sub-entities' submission order: 0 ascending, 1 descending (-1: OGRE's)")
    ap.add_argument("--anim-time", type=int, default=0, help="N frames of the animated entity (re-skinned every frame): seconds per frame")
    ap.add_argument("--formats", default="", help="comma-separated Ogre::PixelFormat names: read-back in that format vs host PixelUtil")
    ap.add_argument("--pass-state", default="", help="cull,depthfunc,alphafunc,alpharef for every sub-entity's pass (0/-1/-1/0 = keep); "
                                                     "honoured by the CUDA plugin only when PH_BEYOND_TINY is set in the environment")
    ap.add_argument("--pose-frames", action="store_true", help="CUDA plugin: the pose-animated frames of tests/golden/pose_blend.npz "
                                                               "through the engine on our buffers and through CudaPoseAnimation")
    ap.add_argument("--scaled", action="store_true", help="with --rtt: scaled read-backs of the render texture vs the real Image::scale")
    ap.add_argument("--shadow", type=int, default=0, help="N shadow-volume cases: real generateShadowVolume vs CudaShadowVolume")
    ap.add_argument("--palettes", type=int, default=0, help="bone palettes at N time positions: engine vs CudaSkeletonPalettes")
    a = ap.parse_args()
    w, h = [int(v) for v in a.size.split("x")]
    so, name = PLUGINS[a.plugin]
    L = C.CDLL(os.path.join(HERE, "libplugin_harness.so"), mode=C.RTLD_GLOBAL)
    L.ph_error.restype = C.c_char_p
    L.ph_render_system_name.restype = C.c_char_p
    L.ph_window_native_target.restype = C.c_void_p
    L.ph_triangles.restype = C.c_size_t
    L.ph_batches.restype = C.c_size_t
    L.ph_skel_num_subs.restype = C.c_size_t
    L.ph_skel_sub_size.restype = C.c_size_t
    L.ph_skel_sub_size.argtypes = [C.c_size_t]
    L.ph_skel_sub_data.restype = C.c_void_p
    L.ph_skel_sub_data.argtypes = [C.c_size_t]
    L.ph_frame_time.restype = C.c_double
    L.ph_scene_open.argtypes = [C.c_char_p, C.c_float]
    L.ph_skel_pose.argtypes = [C.c_float]
    L.ph_read_colour.argtypes = [C.c_void_p, C.c_int]
    L.ph_read_colour_box.argtypes = [C.c_uint32] * 4 + [C.c_void_p]
    L.ph_render_to_texture.argtypes = [C.c_void_p]

    def ok(rc, what):
        if rc:
            raise SystemExit("%s failed (%d): %s" % (what, rc, L.ph_error().decode()))

    L.ph_get_mxcsr.restype = C.c_uint
    L.ph_set_mxcsr.argtypes = [C.c_uint]
    mxcsr_in = L.ph_get_mxcsr()
    if os.environ.get("PH_MXCSR"):                    # diagnostic: force the SSE control register (0x1f80 = IEEE default)
        L.ph_set_mxcsr(int(os.environ["PH_MXCSR"], 0))
    ok(L.ph_open(so.encode(), name.encode(), os.path.join(REFLIB, "Codec_STBI.so").encode(), MEDIA.encode(), w, h), "ph_open")
    out = {"mxcsr_at_start": np.array(mxcsr_in), "mxcsr_after_open": np.array(L.ph_get_mxcsr()),
           "render_system": np.array(L.ph_render_system_name().decode()),
           "optimised_util_replaced": np.array(L.ph_optimised_util_replaced())}
    ok(L.ph_scene_open(a.mesh.encode(), a.cam_z), "ph_scene_open")
    if a.order >= 0:
        ok(L.ph_pin_submission_order(a.order), "ph_pin_submission_order")
    if a.skel:
        for k, t in enumerate(float(v) for v in a.skel.split(",")):
            ok(L.ph_skel_pose(t), "ph_skel_pose")
            for i in range(L.ph_skel_num_subs()):
                n = L.ph_skel_sub_size(i)
                out["skinned_t%d_s%d" % (k, i)] = np.ctypeslib.as_array(C.cast(L.ph_skel_sub_data(i), C.POINTER(C.c_uint8)), shape=(n,)).copy()
    if a.palettes:
        L.ph_skel_num_bones.restype = C.c_size_t
        L.ph_skel_palettes.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        nb = L.ph_skel_num_bones()
        times = (np.arange(a.palettes, dtype=np.float32) * np.float32(0.0371)).astype(np.float32)
        eng = np.zeros((a.palettes, nb, 16), np.float32); dev = np.zeros_like(eng)
        ok(L.ph_skel_palettes(a.palettes, times.ctypes.data, eng.ctypes.data, dev.ctypes.data), "ph_skel_palettes")
        out["palettes_engine"], out["palettes_device"], out["palette_times"] = eng, dev, times
    if a.shadow:
        L.ph_shadow_num_groups.restype = C.c_size_t
        L.ph_shadow_volume.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32] + [C.c_void_p] * 5
        ng = L.ph_shadow_num_groups()
        rng = np.random.default_rng(12)
        cap = 150000
        bad = []
        for k in range(a.shadow):
            if k % 3 == 0:
                v = rng.normal(size=3); light = np.array([*(v / np.linalg.norm(v)), 0], np.float32)
            else:
                light = np.array([*(rng.uniform(-1, 1, 3) * rng.choice([5, 40, 300])), 1], np.float32)
            half = rng.uniform(3, 40, 3); ctr = rng.uniform(-5, 5, 3)
            bounds = np.array([*(ctr - half), *(ctr + half)], np.float32)
            flags, sep, used = k % 8, (k // 2) % 2, [0, 1000, cap - 500][k % 3]
            eng = np.zeros(cap, np.uint16); dev = np.zeros(cap, np.uint16)
            re_, rd = np.zeros((ng, 4), np.uint32), np.zeros((ng, 4), np.uint32)
            used2 = np.zeros(2, np.uint32)
            ok(L.ph_shadow_volume(light.ctypes.data, bounds.ctypes.data, flags, sep, cap, used, eng.ctypes.data, dev.ctypes.data,
                                  re_.ctypes.data, rd.ctypes.data, used2.ctypes.data), "ph_shadow_volume")
            cols = 4 if (sep and flags & 1) else 2
            if not (np.array_equal(eng, dev) and used2[0] == used2[1] and np.array_equal(re_[:, :cols], rd[:, :cols])):
                bad.append(k)
            if k == 0:
                out["shadow_indices_emitted"] = np.array(int((eng != 0xCDCD).sum()))
        out["shadow_cases"], out["shadow_mismatches"] = np.array(a.shadow), np.array(bad, np.int64)
    if a.pose_frames:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from datagen import pose_entity_case
        frames = np.load(os.path.join(ROOT, "tests", "golden", "pose_blend.npz"))["frames"]
        L.ph_pose_entity.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_void_p]
        for k, (seed, nv, nposes, normals) in enumerate(frames):
            base, offs, idx, off, nrm, infl = pose_entity_case(int(seed), int(nv), int(nposes))
            eng = np.zeros((int(nv), 6), np.float32); dev = np.zeros_like(eng)
            ok(L.ph_pose_entity(int(nv), base.ctypes.data, int(nposes), offs.ctypes.data, idx.ctypes.data, off.ctypes.data, nrm.ctypes.data,
                                infl.ctypes.data, int(normals), eng.ctypes.data, dev.ctypes.data), "ph_pose_entity")
            out["pose_engine_%d" % k], out["pose_device_%d" % k] = eng, dev
    if a.pass_state:
        ok(L.ph_set_pass_state(*[int(v) for v in a.pass_state.split(",")]), "ph_set_pass_state")
    ok(L.ph_render(), "ph_render")
    colour = np.zeros((h, w, 3), np.uint8)
    ok(L.ph_read_colour(colour.ctypes.data, 0), "ph_read_colour")
    out["colour"] = colour
    rgba = np.zeros((h, w, 4), np.uint8)
    ok(L.ph_read_colour(rgba.ctypes.data, 1), "ph_read_colour rgba")
    out["colour_rgba"] = rgba
    if a.formats:
        L.ph_read_colour_format_check.restype = C.c_long
        L.ph_read_colour_format_check.argtypes = [C.c_char_p]
        out["format_names"] = np.array(a.formats.split(","))
        out["format_mismatches"] = np.array([L.ph_read_colour_format_check(v.encode()) for v in a.formats.split(",")])
    box = np.zeros((h // 2 - 3, w // 2 - 5, 3), np.uint8)
    ok(L.ph_read_colour_box(5, 3, w // 2, h // 2, box.ctypes.data), "ph_read_colour_box")
    out["colour_box"] = box
    out["stats"] = np.array([L.ph_triangles(), L.ph_batches()])
    rt = L.ph_window_native_target()
    if rt and a.plugin == "cuda":
        sys.path.insert(0, ROOT)
        from ogre_b200 import capi
        depth = np.zeros((h, w), np.float32)
        capi.check(capi.lib().ogrecu_rt_read_depth(rt, 0, 0, w, h, depth.ctypes.data, w * 4))
        out["depth"] = depth
    if a.rtt:
        rtt = np.zeros((h, w, 3), np.uint8)
        ok(L.ph_render_to_texture(rtt.ctypes.data), "ph_render_to_texture")
        out["rtt"] = rtt
        if a.scaled:
            L.ph_rtt_scaled_check.restype = C.c_long
            L.ph_rtt_scaled_check.argtypes = [C.c_uint32] * 7 + [C.c_char_p]
            cases = []
            for fmt in ("PF_BYTE_RGB", "PF_A8R8G8B8", "PF_BYTE_RGBA", "PF_X8B8G8R8", "PF_R5G6B5", "PF_L8", "PF_FLOAT32_RGB"):
                for (l, t, r, b, dw, dh, pad) in ((0, 0, w, h, w // 2, h // 2, 0), (0, 0, w, h, w * 2 + 1, h + 7, 5), (3, 5, w - 7, h - 2, 97, 33, 2),
                                                  (10, 10, 11, 11, 8, 8, 0), (0, 0, w, h, 1, 1, 0), (0, 0, w, 1, w, 3, 1)):
                    cases.append((fmt, l, t, r, b, dw, dh, pad, L.ph_rtt_scaled_check(l, t, r, b, dw, dh, pad, fmt.encode())))
            out["scaled_cases"] = np.array([c[0] + " %d,%d,%d,%d->%dx%d+%d" % c[1:8] for c in cases])
            out["scaled_mismatches"] = np.array([c[8] for c in cases])
        if a.quad:
            ok(L.ph_show_rtt_on_quad(), "ph_show_rtt_on_quad")
            q = np.zeros((h, w, 3), np.uint8)
            ok(L.ph_read_colour(q.ctypes.data, 0), "ph_read_colour quad")
            out["quad"] = q
    if a.time:
        out["frame_seconds"] = np.array(L.ph_frame_time(a.time))
    if a.anim_time:
        L.ph_anim_frame_time.restype = C.c_double
        L.ph_anim_frame_time.argtypes = [C.c_int, C.c_float]
        out["anim_frame_seconds"] = np.array(L.ph_anim_frame_time(a.anim_time, 0.013))
        print("anim_frame_ms %.4f" % (float(out["anim_frame_seconds"]) * 1e3))
    np.savez_compressed(a.out, **out)
    L.ph_close()
    print("ok", a.plugin, str(out["render_system"]), "tris", out["stats"][0], "batches", out["stats"][1])


if __name__ == "__main__":
    main()
