#!/usr/bin/env python3
"""BASELINE configs 4 and 5 through the C ABI's DEVICE GROUP (ogrecu_init_devices): ONE process drives N B200s — what a C++ OGRE
application that loads RenderSystem_CUDA.so with the "Devices" option gets.  No torch.distributed, no NCCL: the geometry lives in
ogrecu_bufs (replicated to the peers over NVLink on first use), config 5's frame is a SPLIT target (interleaved bands, every
device rasterises its bands straight into the root's planes), config 4's views are targets PLACED round-robin on the devices.

  python tools/bench_group.py --devices 0,1[,2,...] [--config c4,c5] [--views 1024] [--c5-tris 20000000] [--band-tiles 4]

Prints one JSON line per config; every frame is checked against the same frame rendered by a single-device context.
Timing: CUDA events on the root's stream (which waits for the peers at the end of every flush), after warm-up, synchronised on
both sides."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ogre_b200  # noqa: E402
from ogre_b200 import DrawDesc, capi, check, lib  # noqa: E402
from datagen import viewport_matrix  # noqa: E402
import bench_configs as bc  # noqa: E402


class RootTimer:
    """CUDA events on the root context's own stream through cuda-python-free ctypes: cudaEvent* from libcudart"""

    def __init__(self, ctx):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.stream = torch.cuda.Stream(device=ctx.device)
        ctx.set_stream(self.stream.cuda_stream)

    def time(self, fn, iters=5, warm=2):
        torch = self.torch
        for _ in range(warm):
            fn()
        self.ctx.sync()
        with torch.cuda.device(self.ctx.device):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(self.stream)
            for _ in range(iters):
                fn()
            b.record(self.stream)
            self.ctx.sync()
            torch.cuda.synchronize(self.ctx.device)
            return a.elapsed_time(b) / iters * 1e-3


def buffer_scene(ctx, verts, idx, tex, w, h):
    """the scene of tools/bench_configs.py in HardwareBuffers (ogrecu_buf): the group's peers read replicas"""
    vb = ogre_b200.HardwareBuffer(ctx, verts.nbytes); vb.write(verts)
    ib = ogre_b200.HardwareBuffer(ctx, idx.nbytes); ib.write(idx)
    texture = ogre_b200.Texture(ctx, tex)
    d = DrawDesc()
    d.positions, d.pos_stride = vb.device_ptr, 32
    d.normals, d.normal_stride = vb.device_ptr + 12, 32
    d.uvs, d.uv_stride = vb.device_ptr + 24, 32
    d.indices, d.index_type, d.draw_count = ib.device_ptr, 4, len(idx)
    d.vertex_count = len(verts)
    d.prim_type = ogre_b200.PRIM_TRIANGLE_LIST
    d.flags = ogre_b200.DRAW_DEPTH_TEST | ogre_b200.DRAW_DEPTH_WRITE | ogre_b200.DRAW_LIGHTING
    d.texture = texture.handle
    d.tex_matrix[:] = np.eye(4, dtype=np.float32).reshape(16)
    d.viewport[:] = viewport_matrix(w, h).reshape(16)
    d.ambient[:] = (0.2, 0.2, 0.25, 1.0)
    return d, (vb, ib, texture)


def free_scene(keep):
    vb, ib, tex = keep
    tex.destroy(); vb.destroy(); ib.destroy()


def views_run(desc, rts, mvps, mvits, lights, chunk):
    L = lib()
    n = len(rts)
    handles = [r.handle for r in rts]

    def run():
        for r in rts:
            r.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            hs = (C.c_void_p * (e - s))(*handles[s:e])
            check(L.ogrecu_draw_views(hs, e - s, C.byref(desc), mvps[s:e].ctypes.data_as(capi.c_float_p),
                                      mvits[s:e].ctypes.data_as(capi.c_float_p), lights[s:e].ctypes.data_as(capi.c_float_p)))
    return run


def c5(devices, tris, band_tiles):
    w, h = 3840, 2160
    n = int(round((tris / 2) ** 0.5))
    verts, idx = bc.terrain(n)
    ntri = len(idx) // 3
    tex = bc.checker_texture(1024)
    mvps, mvits, lights = bc.camera_ring(1, w, h)
    out = {"config": "c5 one 3840x2160 frame, %d triangles, textured, lit: C-ABI device group of %d GPUs in one process, split target "
                     "(this opaque, depth-tested content takes the sort-last form: every device draws its slice of the triangles into private "
                     "planes, each device merges its share of the pixels into the root's planes through peer memory; other content: "
                     "interleaved bands of %d tile rows rasterised straight into the root's planes)" % (ntri, len(devices), band_tiles), "n_gpus": len(devices), "triangles": ntri}
    frames = {}
    for name, devs in (("single", devices[:1]), ("group", devices)):
        ctx = ogre_b200.Context(devices=devs) if len(devs) > 1 else ogre_b200.Context(devs[0])
        if len(devs) > 1:
            ctx.set_config_option("split_band_tiles", band_tiles)
        T = RootTimer(ctx)
        desc, keep = buffer_scene(ctx, verts, idx, tex, w, h)
        rt = ogre_b200.RenderTarget(ctx, w, h, placement=ogre_b200.PLACE_SPLIT)
        desc.mvp[:] = mvps[0].reshape(16); desc.mvit[:] = mvits[0].reshape(16); desc.light_dir[:] = lights[0]

        def run():
            rt.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
            rt.draw(desc)
            rt.flush()
        dt = T.time(run, iters=5, warm=2)
        frames[name] = rt.read_colour()
        host = ogre_b200.pinned_empty((h, w, 3))
        import time
        run(); rt.read_colour(out=host)
        t0 = time.perf_counter()
        for _ in range(3):
            run(); rt.read_colour(out=host)
        dt_rb = (time.perf_counter() - t0) / 3
        out[name] = {"ms_per_frame": dt * 1e3, "frames_per_s": 1 / dt, "Mtri_per_s": ntri / dt / 1e6, "ms_per_frame_with_readback": dt_rb * 1e3,
                     "stats": rt.stats()}
        rt.destroy(); free_scene(keep)
        ctx.set_stream(None); ctx.close()
    out["identical_to_single_gpu_frame"] = bool(np.array_equal(frames["single"], frames["group"]))
    out["speedup"] = out["single"]["ms_per_frame"] / out["group"]["ms_per_frame"]
    return out


def c4(devices, views, chunk_per_device=32):
    w, h = 640, 480
    verts, idx = bc.terrain(707)
    ntri = len(idx) // 3
    tex = bc.checker_texture(1024)
    mvps, mvits, lights = bc.camera_ring(views, w, h)
    G = len(devices)
    out = {"config": "c4 batched offscreen views: %d views 640x480, %d triangles, 1024^2 texture, lit: C-ABI device group of %d GPUs in one "
                     "process, target i placed on device i mod G, %d views per device and ogrecu_draw_views call" % (views, ntri, G, chunk_per_device),
           "n_gpus": G, "views": views}
    check_views = sorted(set([0, 1, views // 2, views - 1]))
    images = {}
    for name, devs in (("single", devices[:1]), ("group", devices)):
        ctx = ogre_b200.Context(devices=devs) if len(devs) > 1 else ogre_b200.Context(devs[0])
        T = RootTimer(ctx)
        desc, keep = buffer_scene(ctx, verts, idx, tex, w, h)
        g = len(devs)
        rts = [ogre_b200.RenderTarget(ctx, w, h, placement=k % g) for k in range(views)]
        run = views_run(desc, rts, mvps, mvits, lights, chunk_per_device * g)
        dt = T.time(run, iters=2, warm=1)
        images[name] = [rts[v].read_colour() for v in check_views]
        out[name] = {"ms_per_view": dt / views * 1e3, "views_per_s": views / dt, "Mtri_per_s": views * ntri / dt / 1e6}
        for r in rts:
            r.destroy()
        free_scene(keep)
        ctx.set_stream(None); ctx.close()
    out["views_identical_to_single_gpu"] = bool(all(np.array_equal(a, b) for a, b in zip(images["single"], images["group"])))
    out["speedup"] = out["group"]["views_per_s"] / out["single"]["views_per_s"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="0,1")
    ap.add_argument("--config", default="c4,c5")
    ap.add_argument("--views", type=int, default=1024)
    ap.add_argument("--c5-tris", type=int, default=20000000)
    ap.add_argument("--band-tiles", type=int, default=4)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    devices = [int(v) for v in a.devices.split(",")]
    lines = []
    for cfg in a.config.split(","):
        r = c5(devices, a.c5_tris, a.band_tiles) if cfg == "c5" else c4(devices, a.views)
        print(json.dumps(r), flush=True)
        lines.append(r)
    if a.out:
        with open(a.out, "w") as f:
            for r in lines:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
