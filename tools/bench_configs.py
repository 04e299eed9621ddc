#!/usr/bin/env python3
"""Throughput of the BASELINE.json configs that are not bench.py's headline (config 2): development /
reporting tool, prints one JSON line per config and (optionally) writes them to a file for profiles/.

  python tools/bench_configs.py --config c1,c3,c4,c5 [--views 1024] [--c5-tris 20000000] [--cpu] [--out file]

  c1  ogrehead 1280x720: the captured _render calls of the golden fixture replayed through the C ABI
  c3  stencil-shadow streams: calculateFaceNormals + 8x calculateLightFacing + 8x extrudeVertices, 2 M triangles
  c4  batched offscreen views: N 640x480 textured, lit views of a 1 M-triangle scene (ogrecu_draw_views)
  c5  one 3840x2160 frame of a 20 M-triangle textured scene (single GPU here; strips in bench of §8e)
CPU numbers (--cpu) come from the unmodified reference (oracle/_ref) when present, else the oracle port,
on a bounded sample; they are baselines, not targets.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ogre_b200  # noqa: E402
from ogre_b200 import DrawDesc, capi, check, lib  # noqa: E402
from datagen import DrawCase, edge_triangles, look_at, perspective, viewport_matrix  # noqa: E402

RANK = int(os.environ.get("RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
PEAK = 6548.2
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


class Timer:
    """CUDA-event timing on the context's stream through torch (the stream is torch's)."""

    def __init__(self, ctx):
        import torch
        self.torch = torch
        self.stream = torch.cuda.Stream()
        ctx.set_stream(self.stream.cuda_stream)
        self.ctx = ctx

    def barrier(self):
        if WORLD > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def time(self, fn, iters=10, warm=3):
        """Seconds per call: CUDA events on the launch stream, barrier + synchronize on both sides, max over ranks."""
        for _ in range(warm):
            fn()
        self.barrier()
        a, b = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        a.record(self.stream)
        for _ in range(iters):
            fn()
        b.record(self.stream)
        self.barrier()
        ms = a.elapsed_time(b)
        if WORLD > 1:
            import torch.distributed as dist
            t = self.torch.tensor([ms], dtype=self.torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / iters * 1e-3


def terrain(n, height=0.6, seed=3):
    """(n+1)^2-vertex height field over [-10,10]^2 (y up), 2*n*n triangles wound to face +y viewers;
    vertex = pos f3, normal f3, uv f2 (stride 32, ogrehead's layout)."""
    g = np.linspace(-10, 10, n + 1)
    x, z = np.meshgrid(g, g, indexing="xy")
    y = height * (np.sin(0.9 * x) * np.cos(0.7 * z) + 0.3 * np.sin(2.3 * x + 1.7 * z))
    dydx = height * (0.9 * np.cos(0.9 * x) * np.cos(0.7 * z) + 0.69 * np.cos(2.3 * x + 1.7 * z))
    dydz = height * (-0.7 * np.sin(0.9 * x) * np.sin(0.7 * z) + 0.51 * np.cos(2.3 * x + 1.7 * z))
    nrm = np.stack([-dydx, np.ones_like(x), -dydz], axis=-1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    uv = np.stack([(x + 10) / 20 * 8, (z + 10) / 20 * 8], axis=-1)
    verts = np.concatenate([np.stack([x, y, z], axis=-1), nrm, uv], axis=-1).reshape(-1, 8).astype(np.float32)
    i = np.arange(n)[None, :] + (n + 1) * np.arange(n)[:, None]
    a, b, c, d = i, i + 1, i + n + 1, i + n + 2
    # Tiny culls cross(p2-p0, p2-p1) > 0 in flipped screen space: this winding survives when seen from above
    tris = np.stack([np.stack([a, c, b], -1), np.stack([b, c, d], -1)], axis=2).reshape(-1, 3)
    return verts, tris.astype(np.uint32).reshape(-1)


def camera_ring(nviews, w, h, radius=22.0):
    proj = perspective(45, w / h, 1.0, 100.0)
    proj[1, :] *= -1          # OgreMain flips the projection for targets with requiresTextureFlipping (Tiny windows)
    mvps, mvits, lights = [], [], []
    for k in range(nviews):
        th = 2 * np.pi * k / nviews
        ph = np.radians(25 + 40 * ((k * 7) % nviews) / max(nviews, 1))
        eye = radius * np.array([np.cos(th) * np.cos(ph), np.sin(ph), np.sin(th) * np.cos(ph)])
        view = look_at(eye, (0, 0, 0))
        mvps.append((proj @ view).astype(np.float32))
        mvits.append(np.linalg.inv(view).T.astype(np.float32))
        l = view[:3, :3] @ np.array([0.4, 0.8, 0.3]); lights.append((l / np.linalg.norm(l)).astype(np.float32))
    # np.stack keeps the F-order of the transposed inverses: force C layout before handing out raw pointers
    return (np.ascontiguousarray(np.stack(mvps)), np.ascontiguousarray(np.stack(mvits)), np.ascontiguousarray(np.stack(lights)))


def checker_texture(size):
    y, x = np.mgrid[0:size, 0:size]
    t = np.zeros((size, size, 4), np.uint8)
    t[..., 0] = 80 + 100 * ((x // 32 + y // 32) % 2); t[..., 1] = 60 + (x * 3) % 160; t[..., 2] = 50 + (y * 5) % 180; t[..., 3] = 255
    return t


def upload_scene(ctx, verts, idx, tex, w, h):
    vb, ib = ctx.to_device(verts), ctx.to_device(idx)
    texture = ogre_b200.Texture(ctx, tex)
    d = DrawDesc()
    d.positions, d.pos_stride = vb.ptr, 32
    d.normals, d.normal_stride = vb.ptr + 12, 32
    d.uvs, d.uv_stride = vb.ptr + 24, 32
    d.indices, d.index_type, d.draw_count = ib.ptr, 4, len(idx)
    d.vertex_count = len(verts)
    d.prim_type = ogre_b200.PRIM_TRIANGLE_LIST
    d.flags = ogre_b200.DRAW_DEPTH_TEST | ogre_b200.DRAW_DEPTH_WRITE | ogre_b200.DRAW_LIGHTING
    d.texture = texture.handle
    d.tex_matrix[:] = np.eye(4, dtype=np.float32).reshape(16)
    d.viewport[:] = viewport_matrix(w, h).reshape(16)
    d.ambient[:] = (0.2, 0.2, 0.25, 1.0)
    return d, (vb, ib, texture)


def c1(ctx, T, args):
    from goldenlib import load_scene
    from rasterlib import GpuScene
    cases, colour, depth, (w, h) = load_scene("ogrehead_1280x720.npz")
    scene = GpuScene(ctx, cases)
    rt = ogre_b200.RenderTarget(ctx, w, h)

    def frame():
        scene.render(rt, (0.1, 0.2, 0.3, 1.0), 1.0)
    dt = T.time(frame, iters=50)
    assert np.array_equal(rt.read_colour(), colour)
    host = ogre_b200.pinned_empty((h, w, 3))             # application read-back buffer (pinned: direct DMA)
    frame(); rt.read_colour(out=host)
    t0 = time.perf_counter()
    for _ in range(20):
        frame(); rt.read_colour(out=host)
    dt_rb = (time.perf_counter() - t0) / 20
    assert np.array_equal(host, colour)
    # the same frame read back as 32-bit BGRA (PF_A8R8G8B8): converted on the device vs RGB read-back + host shuffle
    host32 = ogre_b200.pinned_empty((h, w), np.uint32)
    frame(); rt.read_colour_packed32(16, 8, 0, 0xFF000000, out=host32)
    t0 = time.perf_counter()
    for _ in range(20):
        frame(); rt.read_colour_packed32(16, 8, 0, 0xFF000000, out=host32)
    dt_rb32 = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(5):
        frame(); rt.read_colour(out=host)
        conv = (host[..., 0].astype(np.uint32) << 16) | (host[..., 1].astype(np.uint32) << 8) | host[..., 2] | 0xFF000000
    dt_rb32_host = (time.perf_counter() - t0) / 5
    assert np.array_equal(conv, host32)
    # the frame read back at half size (RenderTexture::copyContentsToMemory into a 640x360 box -> Image::scale): resampled on the
    # device (ogrecu_rt_read_colour_scaled) vs full read-back + the restated Image::scale on one host core (C, -O2)
    import oracle
    half = rt.read_colour_scaled(w // 2, h // 2)
    t0 = time.perf_counter()
    for _ in range(20):
        frame(); half = rt.read_colour_scaled(w // 2, h // 2)
    dt_half = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(5):
        frame(); rt.read_colour(out=host)
        half_host = oracle.scale_bilinear(host, w // 2, h // 2)
    dt_half_host = (time.perf_counter() - t0) / 5
    assert np.array_equal(half, half_host)
    out = {"config": "c1 ogrehead 1280x720 (4336 tris, 4 draws) replayed through the C ABI", "ms_per_frame_device": dt * 1e3,
           "ms_per_frame_with_half_size_readback_device_scale": dt_half * 1e3,
           "ms_per_frame_with_half_size_readback_host_scale_1core": dt_half_host * 1e3,
           "frames_per_s_device": 1 / dt, "ms_per_frame_with_readback": dt_rb * 1e3, "frames_per_s_with_readback": 1 / dt_rb,
           "Mtri_per_s_device": 4336 / dt / 1e6, "image_identical_to_tiny": True,
           "ms_per_frame_with_bgra_readback_device_conversion": dt_rb32 * 1e3,
           "ms_per_frame_with_bgra_readback_host_conversion_numpy": dt_rb32_host * 1e3}
    if args.cpu:
        from oracle import ref
        if ref.available():
            l = ref.lib()
            assert l.ref_scene_open(b"ogrehead.mesh", w, h, 140.0) == 0
            out["cpu_tiny_ms_per_frame"] = l.ref_scene_time(5) * 1e3
            out["cpu"] = "real TinyRenderSystem through Root::renderOneFrame, 1 thread"
    return out


def manifold_edges(tri_idx, nv):
    """EdgeData of a closed manifold triangle list (vectorised EdgeListBuilder stand-in for the synthetic mesh): every
    directed half-edge (v0, v1) of triangle t pairs with the opposite half-edge (v1, v0) of its neighbour; the edge is
    stored once, by the half-edge that comes first in triangle order, as {tri0, tri1, v0, v1, degenerate=0}."""
    nt = len(tri_idx)
    v0 = tri_idx.reshape(-1).astype(np.int64)
    v1 = np.roll(tri_idx, -1, axis=1).reshape(-1).astype(np.int64)
    t = np.repeat(np.arange(nt, dtype=np.int64), 3)
    key, rkey = v0 * nv + v1, v1 * nv + v0
    order = np.argsort(key)
    pos = np.searchsorted(key[order], rkey)
    mate = order[pos]                                   # half-edge index of the opposite direction
    assert np.array_equal(key[mate], rkey), "mesh is not a closed manifold"
    keep = np.arange(len(key)) < mate                   # the earlier half-edge owns the edge
    e = np.stack([t[keep], t[mate[keep]], v0[keep], v1[keep], np.zeros(keep.sum(), np.int64)], 1).astype(np.uint32)
    groups = np.array([[0, len(e), 0, nt, nv]], np.uint32)
    return {"groups": groups, "edges": e, "tris": tri_idx.astype(np.uint32)}


def c3(ctx, T, args):
    nu = int(round((args.c3_tris / 2) ** 0.5))
    nv, nt = nu * nu, 2 * nu * nu
    # closed torus grid: V = nu*nu, T = 2*nu*nu
    u, v = np.meshgrid(np.linspace(0, 2 * np.pi, nu, endpoint=False), np.linspace(0, 2 * np.pi, nu, endpoint=False), indexing="ij")
    pos = np.stack([(3 + np.cos(v)) * np.cos(u), np.sin(v), (3 + np.cos(v)) * np.sin(u)], -1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(nu), np.arange(nu), indexing="ij")
    a = i * nu + j; b = ((i + 1) % nu) * nu + j; c = i * nu + (j + 1) % nu; d = ((i + 1) % nu) * nu + (j + 1) % nu
    tri_idx = np.stack([np.stack([a, b, c], -1), np.stack([b, d, c], -1)], 2).reshape(-1, 3)
    tris = np.zeros((nt, 8), np.uint32)
    tris[:, 2:5] = tri_idx; tris[:, 5:8] = tri_idx
    lights = np.array([[12 * np.cos(k * np.pi / 4), 6.0, 12 * np.sin(k * np.pi / 4), 1.0] for k in range(8)], np.float32)
    dpos, dtri = ctx.to_device(pos), ctx.to_device(tris)
    dfn = ctx.alloc(nt * 16); dfac = ctx.alloc(8 * nt); dext = ctx.alloc(8 * nv * 12)
    L = lib()
    lp = lights.ctypes.data_as(capi.c_float_p)

    def fn():
        check(L.ogrecu_face_normals_dev(ctx.handle, dpos.ptr, dtri.ptr, dfn.ptr, nt))

    def facing():
        check(L.ogrecu_light_facing_multi_dev(ctx.handle, lp, 8, dfn.ptr, dfac.ptr, nt, nt))

    def extrude():
        for k in range(8):
            check(L.ogrecu_extrude_dev(ctx.handle, C.cast(lights[k].ctypes.data, capi.c_float_p), 40.0, dpos.ptr, dext.ptr + k * nv * 12, nv))

    def extrude_fused():
        check(L.ogrecu_extrude_multi_dev(ctx.handle, lp, 8, 40.0, dpos.ptr, dext.ptr, nv * 3, nv))

    def frame():
        fn(); facing(); extrude_fused()
    t_fn, t_fac, t_ext, t_all = T.time(fn), T.time(facing), T.time(extrude), T.time(frame)
    t_extf = T.time(extrude_fused)
    b_extf = nv * 12 + 8 * nv * 12
    # silhouette extraction (ShadowCaster::generateShadowVolume) for the 8 lights, 32-bit indices (V + V > 65 536)
    el = ogre_b200.EdgeList(ctx, manifold_edges(tri_idx, nv))
    didx = ctx.alloc(8 * 12 * nt * 4)
    sil = {"emitted": 0}

    def silhouettes():
        used = 0
        for k in range(8):
            _, used, n = el.shadow_volume(dfac.ptr + k * nt, False, 0, 3, False, 4, didx.ptr, 8 * 12 * nt, used)
        sil["emitted"] = used
    t0 = time.perf_counter(); silhouettes(); ctx.sync()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        silhouettes()
    ctx.sync()
    t_sil = (time.perf_counter() - t0) / reps           # synchronous API (counts come back to the host): wall clock
    # the same for the 8 lights in one submission without a host round trip, and the whole frame (normals -> facing -> extrusion ->
    # silhouettes) back to back on the stream
    region = 12 * nt
    dirs, mcs, fls = (C.c_int * 8)(*[0] * 8), (C.c_int * 8)(*[0] * 8), (C.c_uint32 * 8)(*[3] * 8)
    rng_host = ogre_b200.pinned_empty((8, len(el.groups), 4), np.uint32); res_host = ogre_b200.pinned_empty((8, 2), np.uint32)

    def silhouettes_multi():
        check(L.ogrecu_shadow_volumes_multi_dev(el.handle, 8, dfac.ptr, nt, dirs, mcs, fls, 0, 4, didx.ptr, region, rng_host.ctypes.data,
                                                res_host.ctypes.data))

    def frame_with_silhouettes():
        fn(); facing(); extrude_fused(); silhouettes_multi()
    t_silm = T.time(silhouettes_multi)
    t_all_sil = T.time(frame_with_silhouettes)
    ctx.sync()
    assert int(res_host[:, 0].sum()) == sil["emitted"] and not res_host[:, 1].any()
    b_fn = nt * 48 + nv * 12
    b_fac = nt * 16 + 8 * nt           # normals read ONCE for the 8 lights (fused), 8 facing bytes written
    b_ext = 8 * nv * 24
    out = {"config": "c3 shadow-volume streams: %d triangles, %d vertices, 8 point lights" % (nt, nv),
           "ms_per_frame": t_all * 1e3, "frames_per_s": 1 / t_all,
           "face_normals": {"ms": t_fn * 1e3, "Mtri_per_s": nt / t_fn / 1e6, "GBps": b_fn / t_fn / 1e9, "frac_of_hbm_peak": b_fn / t_fn / 1e9 / PEAK},
           "light_facing_8_fused": {"ms": t_fac * 1e3, "Mface_light_per_s": 8 * nt / t_fac / 1e6, "GBps": b_fac / t_fac / 1e9, "frac_of_hbm_peak": b_fac / t_fac / 1e9 / PEAK},
           "extrude_8": {"ms": t_ext * 1e3, "Mvert_per_s": 8 * nv / t_ext / 1e6, "GBps": b_ext / t_ext / 1e9, "frac_of_hbm_peak": b_ext / t_ext / 1e9 / PEAK},
           "extrude_8_fused": {"ms": t_extf * 1e3, "Mvert_per_s": 8 * nv / t_extf / 1e6, "GBps": b_extf / t_extf / 1e9,
                               "frac_of_hbm_peak": b_extf / t_extf / 1e9 / PEAK, "note": "one launch, positions read once (frame uses this)"},
           "shadow_volume_indices_8": {"ms": t_sil * 1e3, "edges": int(len(el.groups) and el.groups[0][1]), "indices_emitted": int(sil["emitted"]),
                                       "Medge_per_s": 8 * int(el.groups[0][1]) / t_sil / 1e6,
                                       "note": "generateShadowVolume for 8 lights, dark + light caps, 32-bit indices; synchronous calls, wall clock"},
           "shadow_volume_indices_8_one_submission": {"ms": t_silm * 1e3, "Medge_per_s": 8 * int(el.groups[0][1]) / t_silm / 1e6,
                                                      "note": "ogrecu_shadow_volumes_multi_dev: 3 launches for the 8 lights, no host round trip; device time"},
           "ms_per_frame_with_silhouettes": t_all_sil * 1e3}
    if args.cpu:
        from oracle import ref
        if ref.available():
            l = ref.lib()
            fnh = np.zeros((nt, 4), np.float32); fac = np.zeros(nt, np.int8); ext = np.zeros((nv, 3), np.float32)
            t0 = time.perf_counter(); l.ref_face_normals(0, pos.ctypes.data, tris.ctypes.data, fnh.ctypes.data, nt); c_fn = time.perf_counter() - t0
            t0 = time.perf_counter()
            for k in range(8):
                l.ref_light_facing(0, lights[k].ctypes.data, fnh.ctypes.data, fac.ctypes.data, nt)
            c_fac = time.perf_counter() - t0
            t0 = time.perf_counter()
            for k in range(8):
                l.ref_extrude(0, lights[k].ctypes.data, 40.0, pos.ctypes.data, ext.ctypes.data, nv)
            c_ext = time.perf_counter() - t0
            out["cpu_sse_1_thread_ms"] = {"face_normals": c_fn * 1e3, "light_facing_8": c_fac * 1e3, "extrude_8": c_ext * 1e3,
                                          "frame": (c_fn + c_fac + c_ext) * 1e3}
    return out


def raster_views(ctx, T, desc, rts, mvps, mvits, lights, chunk, descs=None):
    """`descs`: optional per-chunk draw descriptors (chunk k uses descs[k % len(descs)]) for targets that live in more than
    one context (two contexts per GPU: their pipeline passes run on different streams and overlap)."""
    L = lib()
    n = len(rts)
    handles = (C.c_void_p * n)(*[r.handle for r in rts])

    def run():
        for r in rts:
            r.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
        for k, s in enumerate(range(0, n, chunk)):
            e = min(n, s + chunk)
            hs = (C.c_void_p * (e - s))(*handles[s:e])
            d = descs[k % len(descs)] if descs else desc
            check(L.ogrecu_draw_views(hs, e - s, C.byref(d), mvps[s:e].ctypes.data_as(capi.c_float_p),
                                      mvits[s:e].ctypes.data_as(capi.c_float_p), lights[s:e].ctypes.data_as(capi.c_float_p)))
    return run


def cpu_view(case, w, h, gpu_image, unit):
    """One frame of `case` on one host core: through the unmodified TinyRenderSystem::_render (oracle/_ref) when it is there,
    else the oracle port; also says whether the GPU image equals it."""
    from oracle import ref
    t0 = time.perf_counter()
    if ref.available():
        rc = ref.tiny_render([case], w, h)[0]
        kind = "reference"
    else:
        from rasterlib import render_oracle
        rc = render_oracle([case], w, h)[0]
        kind = "port"
    dt = time.perf_counter() - t0
    return {"cpu_s_per_%s_1_thread" % unit: dt, "cpu_kind": kind, "identical_to_cpu": bool(np.array_equal(rc, gpu_image))}


def c4(ctx, T, args):
    w, h = 640, 480
    n = int(round((args.c4_tris / 2) ** 0.5))
    verts, idx = terrain(n)
    ntri = len(idx) // 3
    desc, keep = upload_scene(ctx, verts, idx, checker_texture(1024), w, h)
    total_views = args.views
    mvps, mvits, lights = camera_ring(total_views, w, h)
    if WORLD > 1:                                        # view i -> rank i mod G, geometry replicated, no collective
        from ogre_b200.shard import view_indices
        mine = view_indices(total_views, RANK, WORLD)
        mvps, mvits, lights = (np.ascontiguousarray(a[mine]) for a in (mvps, mvits, lights))
    nviews = len(mvps)
    descs = None
    if args.contexts > 1:
        # application-level overlap: the targets of chunk k live in context k % contexts (own stream, own pipeline scratch);
        # geometry is shared (raw device pointers), the texture is uploaded once per context
        import torch
        ctxs, descs = [ctx], [desc]
        for _ in range(args.contexts - 1):
            c2 = ogre_b200.Context(LOCAL)
            st2 = torch.cuda.Stream()
            c2.set_stream(st2.cuda_stream)
            tex2 = ogre_b200.Texture(c2, checker_texture(1024))
            d2 = DrawDesc.from_buffer_copy(desc)
            d2.texture = tex2.handle
            keep = keep + (c2, st2, tex2)
            ctxs.append(c2); descs.append(d2)
        rts = [ogre_b200.RenderTarget(ctxs[(k // args.chunk) % len(ctxs)], w, h) for k in range(nviews)]
        run_inner = raster_views(ctx, T, desc, rts, mvps, mvits, lights, args.chunk, descs)
        side_streams = [x for x in keep if isinstance(x, torch.cuda.Stream)]

        def run():
            ev = torch.cuda.Event(); ev.record(T.stream)
            for st in side_streams:
                st.wait_event(ev)
            run_inner()
            for st in side_streams:
                T.stream.wait_stream(st)
    else:
        rts = [ogre_b200.RenderTarget(ctx, w, h) for _ in range(nviews)]
        run = raster_views(ctx, T, desc, rts, mvps, mvits, lights, args.chunk)
    dt = T.time(run, iters=2, warm=1)
    stats = rts[0].stats()
    host = ogre_b200.pinned_empty((nviews, h, w, 3))     # one pinned read-back buffer for all views
    imgs = [r.read_colour() for r in rts]                # pageable destination (staged), also warms the path
    t0 = time.perf_counter()
    run()
    for k, r in enumerate(rts):
        r.read_colour(out=host[k])
    dt_rb = time.perf_counter() - t0
    assert all(np.array_equal(host[k], imgs[k]) for k in range(0, nviews, max(1, nviews // 8)))
    covered = float(np.mean([(im != np.array([25, 51, 76], np.uint8)).any(axis=2).mean() for im in imgs[:8]]))
    if WORLD > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([dt_rb], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_rb = float(t.item())
    out = {"config": "c4 batched offscreen views: %d views 640x480, %d triangles, 1024^2 texture, lit" % (total_views, ntri),
           "n_gpus": WORLD, "views_per_gpu": nviews, "sharding": "view i -> GPU i mod G, geometry replicated, no collective",
           "views_per_s_device": total_views / dt, "Mtri_per_s_device": total_views * ntri / dt / 1e6, "ms_per_view_device": dt / total_views * 1e3,
           "views_per_s_with_readback": total_views / dt_rb, "covered_fraction": covered, "stats_view0": stats,
           "compulsory_GBps_per_gpu": nviews * (len(idx) * 4 + len(verts) * 32 + w * h * 7) / dt / 1e9}
    if args.cpu:
        case = DrawCase(verts, idx, index_type=4, texture=checker_texture(1024), mvp=mvps[0], mvit=mvits[0], viewport=viewport_matrix(w, h),
                        light_dir=(1, 0, 0), ambient=(0.2, 0.2, 0.25, 1.0))
        case.light_dir = lights[0]
        out.update(cpu_view(case, w, h, imgs[0], "view"))
    for r in rts:
        r.destroy()
    return out


def c5(ctx, T, args):
    w, h = 3840, 2160
    n = int(round((args.c5_tris / 2) ** 0.5))
    verts, idx = terrain(n)
    ntri = len(idx) // 3
    desc, keep = upload_scene(ctx, verts, idx, checker_texture(1024), w, h)
    mvps, mvits, lights = camera_ring(1, w, h)
    rt = ogre_b200.RenderTarget(ctx, w, h)
    run = raster_views(ctx, T, desc, [rt], mvps, mvits, lights, 1)
    if WORLD > 1:
        args._c5_desc, args._c5_cam = desc, (mvps, mvits, lights)
        return c5_split(ctx, T, args, rt, run, ntri, len(idx), len(verts))
    dt = T.time(run, iters=3, warm=1)
    stats = rt.stats()
    host = ogre_b200.pinned_empty((h, w, 3))
    run(); img = rt.read_colour()
    t0 = time.perf_counter()
    for _ in range(3):
        run(); rt.read_colour(out=host)
    dt_rb = (time.perf_counter() - t0) / 3
    assert np.array_equal(host, img)
    out = {"config": "c5 one 3840x2160 frame, %d triangles, textured, lit (1 GPU, whole frame)" % ntri,
           "ms_per_frame_device": dt * 1e3, "frames_per_s_device": 1 / dt, "Mtri_per_s_device": ntri / dt / 1e6,
           "ms_per_frame_with_readback": dt_rb * 1e3, "stats": stats,
           "compulsory_GBps": (len(idx) * 4 + len(verts) * 32 + w * h * 7) / dt / 1e9,
           "covered_fraction": float((img != np.array([25, 51, 76], np.uint8)).any(axis=2).mean())}
    if args.cpu:
        case = DrawCase(verts, idx, index_type=4, texture=checker_texture(1024), mvp=mvps[0], mvit=mvits[0], viewport=viewport_matrix(w, h),
                        light_dir=(1, 0, 0), ambient=(0.2, 0.2, 0.25, 1.0))
        case.light_dir = lights[0]
        out.update(cpu_view(case, w, h, img, "frame"))
    rt.destroy()
    return out


def c5_split(ctx, T, args, rt, run, ntri, nidx, nverts):
    """Split-frame form of c5: every rank holds the geometry, rasterises its strip of screen-tile rows
    (ogrecu_rt_set_strip) and the disjoint strips are gathered to rank 0 over NVLink (NCCL gather) at readback only.
    Checked on rank 0 against the same frame rendered whole on one GPU."""
    import torch
    import torch.distributed as dist
    from ogre_b200.shard import DevicePlane, composite_sort_last, gather_bands, gather_strips, strip_rows, triangle_range
    w, h = rt.width, rt.height
    p = C.c_void_p()
    check(lib().ogrecu_rt_colour_device_ptr(rt.handle, C.byref(p)))
    plane = DevicePlane(p.value, (h, w, 3)).tensor(torch.device("cuda", LOCAL))
    y0, y1 = strip_rows(h, RANK, WORLD)
    rt.set_strip(y0, y1)
    result = [None]

    def render_and_composite():
        run()
        with torch.cuda.stream(T.stream):
            result[0] = gather_strips(plane[y0:y1], y0, y1, h, dst=0)
    dt_render = T.time(run, iters=3, warm=1)
    strip_stats = rt.stats()
    dt = T.time(render_and_composite, iters=3, warm=1)
    ok = None
    if RANK == 0:
        frame = result[0].cpu().numpy()
        rt.set_strip(0, h)
        run()
        whole = rt.read_colour()
        ok = bool(np.array_equal(frame, whole))
    # the load-balanced split: interleaved bands of one tile row each (band b -> rank b mod G)
    rt.set_bands(16, RANK, WORLD)

    def render_and_composite_bands():
        run()
        with torch.cuda.stream(T.stream):
            result[0] = gather_bands(plane, h, 16, dst=0)
    dt_render_b = T.time(run, iters=3, warm=1)
    band_stats = rt.stats()
    dt_b = T.time(render_and_composite_bands, iters=3, warm=1)
    ok_b = None
    if RANK == 0:
        ok_b = bool(np.array_equal(result[0].cpu().numpy(), whole))
    T.barrier()
    interleaved = {"ms_per_frame_render_only": dt_render_b * 1e3, "ms_per_frame_with_composite": dt_b * 1e3, "frames_per_s": 1 / dt_b,
                   "Mtri_per_s": ntri / dt_b / 1e6, "identical_to_single_gpu_frame": ok_b, "stats_rank0": band_stats,
                   "split": "interleaved bands of one tile row (ogrecu_rt_set_bands), band b -> GPU b mod G"}
    # sort-last: this rank draws only ITS slice of the triangles into the full-size target; planes merged by depth
    rt.set_strip(0, h)
    pz = C.c_void_p()
    check(lib().ogrecu_rt_depth_device_ptr(rt.handle, C.byref(pz)))
    zplane = DevicePlane(pz.value, (h, w), "<f4").tensor(torch.device("cuda", LOCAL))
    lo, hi = triangle_range(ntri, RANK, WORLD)
    dslice = DrawDesc.from_buffer_copy(args._c5_desc)
    dslice.indices = args._c5_desc.indices + lo * 12
    dslice.draw_count = (hi - lo) * 3
    hs = (C.c_void_p * 1)(rt.handle)
    mv, mi, li = args._c5_cam

    def render_slice():
        rt.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
        check(lib().ogrecu_draw_views(hs, 1, C.byref(dslice), mv.ctypes.data_as(capi.c_float_p), mi.ctypes.data_as(capi.c_float_p),
                                      li.ctypes.data_as(capi.c_float_p)))

    def render_and_composite_sl():
        render_slice()
        check(lib().ogrecu_flush(rt.handle))
        with torch.cuda.stream(T.stream):
            result[0] = composite_sort_last(ctx, plane, zplane, h, (0.1, 0.2, 0.3, 1.0), 1.0, dst=0)

    def render_slice_flush():
        render_slice()
        check(lib().ogrecu_flush(rt.handle))
    dt_render_sl = T.time(render_slice_flush, iters=5, warm=2)
    sl_stats = rt.stats()
    dt_sl = T.time(render_and_composite_sl, iters=10, warm=5)          # the all-to-all connects its peers lazily: warm it up

    def composite_only():
        with torch.cuda.stream(T.stream):
            result[0] = composite_sort_last(ctx, plane, zplane, h, (0.1, 0.2, 0.3, 1.0), 1.0, dst=0)
    dt_comp = T.time(composite_only, iters=10, warm=2)
    render_slice_flush()
    with torch.cuda.stream(T.stream):
        result[0] = composite_sort_last(ctx, plane, zplane, h, (0.1, 0.2, 0.3, 1.0), 1.0, dst=0)
    T.barrier()
    ok_sl = None
    if RANK == 0:
        ok_sl = bool(np.array_equal(result[0][0].cpu().numpy(), whole))
    T.barrier()
    sort_last = {"ms_per_frame_render_only": dt_render_sl * 1e3, "ms_per_frame_with_composite": dt_sl * 1e3, "ms_composite_only": dt_comp * 1e3,
                 "frames_per_s": 1 / dt_sl,
                 "Mtri_per_s": ntri / dt_sl / 1e6, "identical_to_single_gpu_frame": ok_sl, "stats_rank0": sl_stats,
                 "exchange_bytes_per_rank": int(w * h * 7 * (WORLD - 1) / WORLD),
                 "split": "sort-last: rank r draws triangles [r T/G, (r+1) T/G), all-to-all of strips, per-pixel depth merge "
                          "(ogrecu_composite_depth_dev), gather to rank 0"}
    return {"config": "c5 one 3840x2160 frame, %d triangles, textured, lit: split-frame over %d GPUs (contiguous strips of tile rows), "
                      "NCCL gather of the strips to rank 0 at readback" % (ntri, WORLD), "n_gpus": WORLD,
            "ms_per_frame_render_only": dt_render * 1e3, "ms_per_frame_with_composite": dt * 1e3, "frames_per_s": 1 / dt,
            "Mtri_per_s": ntri / dt / 1e6, "composite_bytes": int(w * h * 3 * (WORLD - 1) / WORLD),
            "identical_to_single_gpu_frame": ok, "stats_rank0_strip": strip_stats, "interleaved": interleaved, "sort_last": sort_last}


def c2j(ctx, T, args):
    """config 2, second mesh: jaiqua (1 sub-mesh, 2 460 vertices, 38 bones of which 23 used, 2 weights) x 10 000 animated
    instances, the real mesh streams and poses of tests/golden/jaiqua_skin.npz; same step as bench.py's (morph + skin, fused
    and separate).  bench.py itself stays on Sinbad (the workload its contract line names)."""
    import torch
    import bench
    from ogre_b200 import SkinJob
    bench.FIXTURE = os.path.join(ROOT, "tests", "golden", "jaiqua_skin.npz")
    subs, palettes = bench.load_sinbad()
    assert len(subs) == 1
    s = subs[0]
    n, ninst, nb = s["nverts"], 10000, palettes.shape[1]
    pal_h, t_h = bench.instance_inputs(ninst, palettes, RANK)
    dev = torch.device("cuda", LOCAL)
    L = lib()
    with torch.cuda.stream(T.stream):
        d_pal, d_t = torch.from_numpy(pal_h).to(dev), torch.from_numpy(t_h).to(dev)
        key1, key2 = torch.from_numpy(s["key1"]).to(dev), torch.from_numpy(s["key2"]).to(dev)
        blend = torch.from_numpy(s["blend"]).to(dev)
        bmap = torch.from_numpy(s["bone_map"].view(np.int16)).to(dev)
        tmp, dst = torch.empty(ninst, n, 6, device=dev), torch.empty(ninst, n, 6, device=dev)
    j = SkinJob()
    j.src_pos, j.src_norm = tmp.data_ptr(), tmp.data_ptr() + 12
    j.dst_pos, j.dst_norm = dst.data_ptr(), dst.data_ptr() + 12
    j.blend_weight, j.blend_index = blend.data_ptr() + 4, blend.data_ptr()
    j.palette, j.bone_map = d_pal.data_ptr(), bmap.data_ptr()
    j.src_pos_stride = j.dst_pos_stride = j.src_norm_stride = j.dst_norm_stride = 24
    j.blend_weight_stride = j.blend_index_stride = s["blend_stride"]
    j.num_weights, j.num_matrices, j.num_vertices = s["nw"], len(s["bone_map"]), n
    jf = SkinJob.from_buffer_copy(j)
    jf.src_pos, jf.src_norm = key1.data_ptr(), key1.data_ptr() + 12
    T.stream.synchronize()

    def morph():
        check(L.ogrecu_morph_instanced_dev(ctx.handle, d_t.data_ptr(), key1.data_ptr(), key2.data_ptr(), tmp.data_ptr(), 24, 24, 24, n, 1,
                                           ninst, n * 24))

    def skin():
        check(L.ogrecu_skin_instanced_dev(ctx.handle, C.byref(j), ninst, nb * 16, n * 24, n * 24))

    def fused():
        check(L.ogrecu_anim_instanced_dev(ctx.handle, C.byref(jf), key2.data_ptr(), d_t.data_ptr(), None, 0, ninst, nb * 16, n * 24))

    def separate():
        morph(); skin()
    # the same step through the lane kernel (round 2: plan + ogrecu_anim_frame_dev, what bench.py times for Sinbad)
    from ogre_b200 import AnimJob
    aj = (AnimJob * 1)()
    aj[0].skin = jf
    aj[0].key2 = key2.data_ptr()
    aj[0].dst_instance_stride_bytes = n * 24
    plan = C.c_void_p()
    check(L.ogrecu_skin_plan_create(ctx.handle, C.byref(aj[0].skin), C.byref(plan)))
    plans = (C.c_void_p * 1)(plan)

    def lane():
        check(L.ogrecu_anim_frame_dev(ctx.handle, aj, plans, 1, d_t.data_ptr(), ninst, nb * 16))
    dt_f, dt_s, dt_skin = T.time(fused, iters=50), T.time(separate, iters=50), T.time(skin, iters=50)
    dt_lane = T.time(lane, iters=50)
    lane(); T.stream.synchronize(); lane_bits = dst[4711].cpu().numpy().copy()
    # parity spot check: instance 4711 of the fused step against the separate one (bit patterns)
    fused(); T.stream.synchronize(); a = dst[4711].cpu().numpy().copy()
    separate(); T.stream.synchronize(); b = dst[4711].cpu().numpy()
    tot = ninst * n * WORLD
    pal_bytes = ninst * len(s["bone_map"]) * 48
    return {"config": "c2 (jaiqua) skeletal skinning: 1 sub-mesh, %d vertices, %d bones (%d used), %d weights x %d instances per GPU; "
                      "morph + skin" % (n, nb, len(s["bone_map"]), s["nw"], ninst), "n_gpus": WORLD,
            "fused_ms": dt_f * 1e3, "fused_Mvert_per_s": tot / dt_f / 1e6, "separate_ms": dt_s * 1e3, "separate_Mvert_per_s": tot / dt_s / 1e6,
            "skin_only_ms": dt_skin * 1e3, "skin_only_Mvert_per_s": tot / dt_skin / 1e6,
            "lane_ms": dt_lane * 1e3, "lane_Mvert_per_s": tot / dt_lane / 1e6, "lane_frac_of_hbm_peak": (ninst * n * 24 + pal_bytes) / dt_lane / 1e9 / PEAK,
            "lane_equals_separate_bits": bool(np.array_equal(lane_bits.view(np.uint32), b.view(np.uint32))),
            "fused_frac_of_hbm_peak": (ninst * n * 24 + pal_bytes) / dt_f / 1e9 / PEAK,
            "skin_only_frac_of_hbm_peak": (ninst * n * 48 + pal_bytes) / dt_skin / 1e9 / PEAK,
            "fused_equals_separate_bits": bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c1,c3,c4,c5")
    ap.add_argument("--views", type=int, default=1024)
    ap.add_argument("--chunk", type=int, default=32, help="views per ogrecu_draw_views pass")
    ap.add_argument("--contexts", type=int, default=1, help="c4: contexts (streams) per GPU whose pipeline passes overlap")
    ap.add_argument("--c3-tris", type=int, default=2000000)
    ap.add_argument("--c4-tris", type=int, default=1000000)
    ap.add_argument("--c5-tris", type=int, default=20000000)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    torch.cuda.set_device(LOCAL)
    if WORLD > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
    ctx = ogre_b200.Context(LOCAL)
    T = Timer(ctx)
    lines = []
    for name in args.config.split(","):
        t0 = time.time()
        res = {"c1": c1, "c2j": c2j, "c3": c3, "c4": c4, "c5": c5}[name](ctx, T, args)
        res["wall_s"] = time.time() - t0
        if RANK == 0:
            print(json.dumps(res), flush=True)
        lines.append(res)
    if args.out and RANK == 0:
        with open(args.out, "w") as f:
            for r in lines:
                f.write(json.dumps(r) + "\n")
    ctx.set_stream(None)
    if WORLD > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
