#!/usr/bin/env python3
"""Development A/B of the two instanced morph+skin paths on the bench workload (Sinbad x N instances):
ogrecu_anim_instanced_multi_dev (tile/pair kernels, seven launches) vs ogrecu_anim_frame_dev (lane kernel from plans,
one launch).  Checks the outputs bit for bit and prints ms per step for each; per-sub-mesh times with --per-sub.
Usage (GPU box): python tools/lane_bench.py [--instances 10000] [--steps 20] [--per-sub]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--per-sub", action="store_true")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import torch
    import ogre_b200
    from ogre_b200 import AnimJob, SkinJob, check, lib
    dev = torch.device("cuda", 0)
    subs, palettes = bench.load_sinbad()
    ninst, nb = args.instances, palettes.shape[1]
    pal_h, t_h = bench.instance_inputs(ninst, palettes, 0)
    ctx = ogre_b200.Context(0)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    L = lib()
    with torch.cuda.stream(stream):
        d_pal, d_t = torch.from_numpy(pal_h).to(dev), torch.from_numpy(t_h).to(dev)
        ds = []
        for s in subs:
            n = s["nverts"]
            e = {"key1": torch.from_numpy(s["key1"]).to(dev), "key2": torch.from_numpy(s["key2"]).to(dev),
                 "blend": torch.from_numpy(s["blend"]).to(dev), "map": torch.from_numpy(s["bone_map"].view(np.int16)).to(dev),
                 "a": torch.zeros(ninst, n, 6, device=dev), "b": torch.zeros(ninst, n, 6, device=dev)}
            ds.append(e)
    stream.synchronize()

    def jobs_for(which, only=None):
        idx = range(len(ds)) if only is None else [only]
        arr = (AnimJob * len(idx))()
        for k, i in enumerate(idx):
            e, s = ds[i], subs[i]
            j = SkinJob()
            j.src_pos, j.src_norm = e["key1"].data_ptr(), e["key1"].data_ptr() + 12
            j.dst_pos, j.dst_norm = e[which].data_ptr(), e[which].data_ptr() + 12
            j.blend_weight, j.blend_index = e["blend"].data_ptr() + 4, e["blend"].data_ptr()
            j.palette, j.bone_map = d_pal.data_ptr(), e["map"].data_ptr()
            j.src_pos_stride = j.dst_pos_stride = j.src_norm_stride = j.dst_norm_stride = 24
            j.blend_weight_stride = j.blend_index_stride = s["blend_stride"]
            j.num_weights, j.num_matrices, j.num_vertices = s["nw"], len(s["bone_map"]), s["nverts"]
            arr[k].skin = j
            arr[k].key2 = e["key2"].data_ptr()
            arr[k].dst_instance_stride_bytes = s["nverts"] * 24
        return arr

    ja, jb = jobs_for("a"), jobs_for("b")
    plans = (C.c_void_p * len(ds))()
    chunks = []
    for k in range(len(ds)):
        p = C.c_void_p()
        check(L.ogrecu_skin_plan_create(ctx.handle, C.byref(jb[k].skin), C.byref(p)))
        plans[k] = p
        nc = C.c_uint32()
        check(L.ogrecu_skin_plan_info(p, C.byref(nc)))
        chunks.append(nc.value)

    def old():
        check(L.ogrecu_anim_instanced_multi_dev(ctx.handle, ja, len(ds), d_t.data_ptr(), ninst, nb * 16))

    def new():
        check(L.ogrecu_anim_frame_dev(ctx.handle, jb, plans, len(ds), d_t.data_ptr(), ninst, nb * 16))

    def timed(fn, steps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps

    old(); new(); torch.cuda.synchronize()
    same = all(torch.equal(e["a"].view(torch.int32), e["b"].view(torch.int32)) for e in ds)
    bad = [i for i, e in enumerate(ds) if not torch.equal(e["a"].view(torch.int32), e["b"].view(torch.int32))]
    nverts = sum(s["nverts"] for s in subs)
    ms_old, ms_new = timed(old, args.steps), timed(new, args.steps)
    prof = None
    if hasattr(C.CDLL(os.environ.get("OGRECU_LIB", "")) if os.environ.get("OGRECU_LIB") else object(), "ogrecu_debug_lane_profile"):
        # a -DLN_PROFILE build: cycles of CTA 0's producer phases and consumer warp 0 over ONE launch
        D = C.CDLL(os.environ["OGRECU_LIB"])
        new(); 
        out = (C.c_uint64 * 16)()
        D.ogrecu_debug_lane_profile(C.c_void_p(ctx.handle), out)
        names = ["producer prep", "producer waits for buffer", "producer issues copies", "producer records/keys", "producer patch+arrive",
                 "consumer loop", "consumer waits for unit", "consumer walks"]
        prof = {n: round(out[k] / 1965.0, 1) for k, n in enumerate(names)}          # microseconds at 1.965 GHz
    out = {"tag": args.tag, "lib": os.environ.get("OGRECU_LIB", "default"), "instances": ninst, "identical": same, "differing_sub_meshes": bad,
           "chunks": chunks, "ms_old": round(ms_old, 4), "ms_new": round(ms_new, 4),
           "gvert_old": round(ninst * nverts / ms_old / 1e6, 1), "gvert_new": round(ninst * nverts / ms_new / 1e6, 1),
           "frac_new": round(ninst * nverts * 24 / (ms_new * 1e-3) / 6537e9, 3), "phase_us_cta0": prof}
    if args.per_sub:
        per = []
        for i in range(len(ds)):
            j1 = jobs_for("b", i)
            p1 = (C.c_void_p * 1)(plans[i])
            j0 = jobs_for("a", i)
            f_new = lambda: check(L.ogrecu_anim_frame_dev(ctx.handle, j1, p1, 1, d_t.data_ptr(), ninst, nb * 16))
            f_old = lambda: check(L.ogrecu_anim_instanced_multi_dev(ctx.handle, j0, 1, d_t.data_ptr(), ninst, nb * 16))
            per.append({"sub": i, "nverts": subs[i]["nverts"], "nw": subs[i]["nw"], "ms_old": round(timed(f_old, args.steps), 4),
                        "ms_new": round(timed(f_new, args.steps), 4)})
        out["per_sub"] = per
    print(json.dumps(out))
    for k in range(len(ds)):
        check(L.ogrecu_skin_plan_destroy(plans[k]))
    ctx.close()


if __name__ == "__main__":
    main()
