"""Kernel-level timing of the OptimisedUtil streams on one GPU (CUDA events on the launch
stream, inputs larger than L2).  Development tool; bench.py is the contract."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ogre_b200
from ogre_b200 import SkinJob, check, lib

PEAK = 6537.0
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, stream, iters=20, warm=3):
    for _ in range(warm):
        fn()
    stream.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(stream)
        fn()
        b.record(stream)
    stream.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    torch.cuda.set_device(0)
    ctx = ogre_b200.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    dev = torch.device("cuda:0")
    out = {}
    with torch.cuda.stream(stream):
        # ---- skinning, synthetic stream: 2^24 vertices, 4 weights, 64 bones, 68 B/vertex
        n, nw, nb = 1 << 24, 4, 64
        g = torch.Generator(device=dev).manual_seed(1234)
        src = torch.empty(n, 6, device=dev).uniform_(-10, 10, generator=g)
        w = torch.rand(n, nw, device=dev, generator=g) + 0.05
        w = w / w.sum(1, keepdim=True)
        idx = torch.randint(0, nb, (n, 4), device=dev, dtype=torch.uint8, generator=g)
        blend = torch.empty(n, 4 + 4 * nw, device=dev, dtype=torch.uint8)
        blend[:, :4] = idx
        blend[:, 4:] = w.contiguous().view(torch.uint8).reshape(n, 4 * nw)
        pal = torch.eye(4, device=dev).repeat(nb, 1, 1) + 0.01 * torch.randn(nb, 4, 4, device=dev, generator=g)
        pal[:, 3, :] = torch.tensor([0, 0, 0, 1.0], device=dev)
        pal = pal.contiguous()
        dst = torch.empty(n, 6, device=dev)
        for nwk in (4, 2, 1):
            j = SkinJob()
            j.src_pos, j.src_norm = src.data_ptr(), src.data_ptr() + 12
            j.dst_pos, j.dst_norm = dst.data_ptr(), dst.data_ptr() + 12
            j.blend_weight, j.blend_index = blend.data_ptr() + 4, blend.data_ptr()
            j.palette, j.bone_map = pal.data_ptr(), None
            j.src_pos_stride = j.dst_pos_stride = j.src_norm_stride = j.dst_norm_stride = 24
            j.blend_weight_stride = j.blend_index_stride = 20
            j.num_weights, j.num_matrices, j.num_vertices = nwk, nb, n
            fn = lambda: check(lib().ogrecu_skin_instanced_dev(ctx.handle, C.byref(j), 1, 0, 0, 0))
            med, best = timeit(fn, stream)
            bpv = 24 + 20 + 24
            out["skin_stream_nw%d" % nwk] = {"ms": med, "best_ms": best, "Gvert/s": n / med / 1e6, "GB/s": n * bpv / med / 1e6,
                                            "frac": n * bpv / med / 1e6 / PEAK}
        # ---- skinning, instanced: 5534-vertex mesh x 10000 instances, 69 bones, nw=4
        nv, ninst, nb = 5534, 10000, 69
        pals = torch.eye(4, device=dev).repeat(ninst * nb, 1, 1).contiguous()
        dsti = torch.empty(ninst, nv, 6, device=dev)
        j = SkinJob()
        j.src_pos, j.src_norm = src.data_ptr(), src.data_ptr() + 12
        j.dst_pos, j.dst_norm = dsti.data_ptr(), dsti.data_ptr() + 12
        j.blend_weight, j.blend_index = blend.data_ptr() + 4, blend.data_ptr()
        j.palette, j.bone_map = pals.data_ptr(), None
        j.src_pos_stride = j.dst_pos_stride = j.src_norm_stride = j.dst_norm_stride = 24
        j.blend_weight_stride = j.blend_index_stride = 20
        j.num_weights, j.num_matrices, j.num_vertices = 4, 64, nv
        fn = lambda: check(lib().ogrecu_skin_instanced_dev(ctx.handle, C.byref(j), ninst, nb * 16, 0, nv * 24))
        med, best = timeit(fn, stream)
        tot = nv * ninst
        out["skin_instanced"] = {"ms": med, "best_ms": best, "Gvert/s": tot / med / 1e6, "write GB/s": tot * 24 / med / 1e6,
                                 "frac_compulsory": (tot * 24 + ninst * nb * 64) / med / 1e6 / PEAK}
        # ---- morph
        a = torch.randn(n * 3, device=dev); b = torch.randn(n * 3, device=dev); d = torch.empty(n * 3, device=dev)
        fn = lambda: check(lib().ogrecu_morph_dev(ctx.handle, 0.37, a.data_ptr(), b.data_ptr(), d.data_ptr(), 12, 12, 12, n, 0))
        med, best = timeit(fn, stream)
        out["morph_pos"] = {"ms": med, "Gvert/s": n / med / 1e6, "GB/s": n * 36 / med / 1e6, "frac": n * 36 / med / 1e6 / PEAK}
        n2 = n // 2
        fn = lambda: check(lib().ogrecu_morph_dev(ctx.handle, 0.37, a.data_ptr(), b.data_ptr(), d.data_ptr(), 24, 24, 24, n2, 1))
        med, best = timeit(fn, stream)
        out["morph_posnorm"] = {"ms": med, "Gvert/s": n2 / med / 1e6, "GB/s": n2 * 72 / med / 1e6, "frac": n2 * 72 / med / 1e6 / PEAK}
        # ---- shadow volume streams: 2M triangles, 1M vertices
        nt, nvx = 2_000_000, 1_000_000
        pos = torch.randn(nvx, 3, device=dev)
        tris = torch.zeros(nt, 8, device=dev, dtype=torch.int32)
        # mesh-like locality: indices near i/2
        base = (torch.arange(nt, device=dev) // 2).clamp(max=nvx - 1100)
        tris[:, 2] = base
        tris[:, 3] = base + torch.randint(1, 1000, (nt,), device=dev, generator=g)
        tris[:, 4] = base + torch.randint(1, 1000, (nt,), device=dev, generator=g)
        fnrm = torch.empty(nt, 4, device=dev)
        fn = lambda: check(lib().ogrecu_face_normals_dev(ctx.handle, pos.data_ptr(), tris.data_ptr(), fnrm.data_ptr(), nt))
        med, best = timeit(fn, stream)
        out["face_normals"] = {"ms": med, "Mtri/s": nt / med / 1e3, "GB/s": (nt * 48 + nvx * 12) / med / 1e6}
        lights = (C.c_float * 32)(*[float(v) for v in np.random.default_rng(1).uniform(-30, 30, 32)])
        for k in range(8):
            lights[4 * k + 3] = 1.0
        fac = torch.empty(8, nt, device=dev, dtype=torch.int8)
        fn = lambda: check(lib().ogrecu_light_facing_multi_dev(ctx.handle, lights, 8, fnrm.data_ptr(), fac.data_ptr(), nt, nt))
        med, best = timeit(fn, stream)
        out["light_facing_x8"] = {"ms": med, "GB/s": nt * (16 + 8) / med / 1e6}
        ext = torch.empty(nvx, 3, device=dev)
        fn = lambda: check(lib().ogrecu_extrude_dev(ctx.handle, lights, 100.0, pos.data_ptr(), ext.data_ptr(), nvx))
        med, best = timeit(fn, stream)
        out["extrude_point"] = {"ms": med, "Mvert/s": nvx / med / 1e3, "GB/s": nvx * 24 / med / 1e6, "frac": nvx * 24 / med / 1e6 / PEAK}
    stream.synchronize()
    ctx.set_stream(None)
    for k, v in out.items():
        print(k, json.dumps({a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()}))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/microbench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
