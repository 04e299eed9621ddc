#!/usr/bin/env python3
"""bench.py — headline benchmark of the OGRE software-rendering hot path on B200.

Contract (driver):  python bench.py --gpus N --steps K --warmup W [--impl reference]
  N>1 is launched under torch.distributed.run (one rank per GPU, NCCL for barrier/max and the split-frame gather only —
  instances and views shard with NO data-path collective, so scaling is "weak").

Workload of the contract line (BASELINE.json configs[1]): Sinbad (7 sub-meshes, 5 534 vertices, 69 bones, 1/4/3/1/1/3/4
weights per vertex; the real mesh streams from tests/golden/sinbad_skin.npz, the real skeleton and its "Dance" animation
from tests/golden/sinbad_skeleton.npz) x 10 000 animated instances per GPU.  One STEP = one frame of vertex animation for
every instance, in the order Entity::updateAnimation runs it (OgreMain/src/OgreEntity.cpp:747-911):
  softwareVertexMorph    two shared keyframes -> the instance's animated source (pos+normal, stride 24)
  softwareVertexSkinning instance source + instance bone palette -> instance output (pos+normal)
as ONE call, ogrecu_anim_frame_dev (palette repack + the instance-pair-per-lane kernel, csrc/skin_lane.cu).
metric = skinned vertices per second (Mvert/s), whole job.

One JSON line (rank 0):
  value      device-timed (CUDA events on the launch stream), palettes and keyframes resident in HBM
  e2e        the same metric through the whole pipeline an OGRE application runs per frame, host inputs to host result:
             H2D each instance's ANIMATION STATE (12 B) and morph weight from pinned memory -> bone palettes on the device
             (ogrecu_skel_palettes_dev = SkeletonInstance::setAnimationState + _getBoneMatrices) -> morph + skin of every
             instance (the skinned buffers stay in HBM) -> the _render that consumes them (all instances drawn into a
             1920x1080 target by the rasteriser) -> D2H of the rendered frame.  The reference arm pays no render.
  roofline   the dominant kernel of the timed step (skin_lane_kernel: one launch covers all seven sub-meshes), timed alone
             with CUDA events, against the measured HBM peak
  cpu_baseline  the unmodified reference (oracle/_ref: real SkeletonInstance + OptimisedUtilSSE) on the host cores
  extra.render  BASELINE's render configs, same run: c1 (ogrehead 1280x720), c4 (1024 offscreen 640x480 views of a
             1 M-triangle scene, view i -> GPU i mod G), c5 (one 3840x2160 frame of 20 M triangles; split-frame strips +
             NCCL gather under torchrun) with frames/s, Mtri/s, roofline, the real Tiny as cpu_baseline and an e2e with
             the full colour read-back.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

FIXTURE = os.path.join(ROOT, "tests", "golden", "sinbad_skin.npz")
FALLBACK_HBM_GBS = 6650.0     # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
FRAME_W, FRAME_H = 1920, 1080  # the e2e step's render target


def load_sinbad():
    z = np.load(FIXTURE)
    subs = []
    for i in range(int(z["nsubs"])):
        m = [int(v) for v in z["s%d_meta" % i]]
        s = dict(zip(("src_stride", "norm_off", "blend_stride", "weight_off", "index_off", "nw", "nverts", "has_norm",
                      "dst_stride"), m))
        assert s["src_stride"] == 24 and s["norm_off"] == 12 and s["index_off"] == 0 and s["weight_off"] == 4
        key1 = z["s%d_src" % i].view(np.float32).reshape(-1, 6).copy()
        # second morph keyframe: a deterministic "breathing" deformation of the bind pose
        key2 = key1.copy()
        key2[:, :3] += 0.05 * np.sin(3.0 * key1[:, [1, 2, 0]]).astype(np.float32) * key1[:, 3:]
        n2 = key1[:, 3:] + 0.1 * np.cos(2.0 * key1[:, :3]).astype(np.float32)
        key2[:, 3:] = (n2 / np.linalg.norm(n2, axis=1, keepdims=True)).astype(np.float32)
        s["key1"], s["key2"] = key1, key2
        s["blend"] = z["s%d_blend" % i].copy()
        s["bone_map"] = z["s%d_bone_map" % i].astype(np.uint16)
        subs.append(s)
    return subs, z["palettes"].astype(np.float32)


def instance_inputs(ninst, palettes, rank):
    """Per-instance bone palettes [ninst, nbones, 16] (real poses, cycled with a per-instance
    root offset) and morph weights t[ninst]."""
    npose = palettes.shape[0]
    idx = (np.arange(ninst) * 7 + rank * 3) % npose
    pal = palettes[idx].copy()
    off = np.zeros((ninst, 1, 16), np.float32)
    off[:, 0, 3] = (np.arange(ninst) % 100) * 3.0        # instances stand on a 100 x 100 grid
    off[:, 0, 11] = (np.arange(ninst) // 100) * 3.0
    pal += off
    t = ((np.arange(ninst) * 0.01) % 1.0).astype(np.float32)
    return np.ascontiguousarray(pal), t


def instance_states(ninst, rank, dance_len):
    """Per-instance animation state: animation 0 ("Dance"), time position, weight 1."""
    anim = np.zeros(ninst, np.int32)
    tpos = np.mod((np.arange(ninst) * 0.01 + rank).astype(np.float32), np.float32(dance_len)).astype(np.float32)
    weight = np.ones(ninst, np.float32)
    return anim, tpos, weight


class Clocks:
    """nvidia-smi clock/throttle sampling during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def workload_config(ninst, nverts, gpus=1):
    """The `config` object: the same for both arms (the reference arm runs the same instances on the host cores)."""
    return {"workload": "config[1] skeletal skinning: Sinbad (7 sub-meshes, %d vertices, 69 bones, 1/4/3/1/1/3/4 weights) x %d animated "
                        "instances per GPU; step = softwareVertexMorph (2 shared keyframes -> instance source) + softwareVertexSkinning "
                        "(pos+normal shared-packed, per-instance palette)" % (nverts, ninst),
            "instances_per_gpu": ninst, "vertices_per_instance": nverts,
            "form": "fused", "parallelism": "instances sharded per GPU, no collective (dp%d)" % gpus,
            "l2": "every step writes 1.33 GB of skinned vertices per GPU (>> 126 MB L2): outputs larger than L2, no flush needed"}


# ------------------------------------------------------------------------------ reference arm

def reference_step_fn(subs, palettes, ninst, threads, rank=0):
    """Returns (fn, kind, cores): fn() runs one frame of `ninst` instances on the CPU with the unmodified reference
    (oracle/_ref) — SkeletonInstance::setAnimationState + _getBoneMatrices per instance (one thread: OGRE evaluates
    entities on the render thread and SkeletonInstance is not thread-safe), then softwareVertexMorph +
    softwareVertexSkinning of every sub-mesh over `threads` host threads — and returns (seconds, seconds in the palettes).
    Falls back to the oracle port (C restatement, palettes from the fixture poses) where oracle/_ref is absent."""
    from oracle import ref
    pal, t = instance_inputs(ninst, palettes, rank)
    nb = pal.shape[1]
    bufs = []
    for s in subs:
        n = s["nverts"]
        bufs.append((np.zeros((ninst, n, 6), np.float32), np.zeros((ninst, n, 6), np.float32)))
    if ref.available():
        l = ref.lib()
        # (ref_skel_num_bones counts the palette of the last evaluated pose: evaluate one)
        have_skel = l.ref_skel_open(b"Sinbad.mesh", 64, 64) == 0 and l.ref_skel_pose(0.0) == 0 and l.ref_skel_num_bones() == nb
        _, tpos, _ = instance_states(ninst, rank, 1e9)
        live = np.zeros((ninst, nb, 16), np.float32)

        def fn():
            tp = 0.0
            src_pal = pal
            if have_skel:
                tp = l.ref_bench_palettes(0, tpos.ctypes.data, ninst, live.ctypes.data)
                src_pal = live
            tot = tp
            for s, (tmp, dst) in zip(subs, bufs):
                bl = s["blend"].ctypes.data
                tot += l.ref_bench_anim_sub(s["key1"].ctypes.data, s["key2"].ctypes.data, t.ctypes.data, bl + 4, bl,
                                            s["blend_stride"], s["blend_stride"], s["nw"], s["nverts"], src_pal.ctypes.data, nb * 16,
                                            s["bone_map"].ctypes.data, len(s["bone_map"]), ninst, tmp.ctypes.data, dst.ctypes.data,
                                            threads, 1)
            return tot, tp
        return fn, "reference", threads
    # oracle port (C restatement), threaded over instances from Python (ctypes drops the GIL)
    import oracle
    ol = oracle.lib()
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(threads)

    def work(k):
        for s, (tmp, dst) in zip(subs, bufs):
            n = s["nverts"]
            bl = s["blend"].ctypes.data
            for i in range(k, ninst, threads):
                mats = np.ascontiguousarray(pal[i][s["bone_map"]])
                keep, pp = oracle.palette_ptrs(mats)
                src = tmp[i].ctypes.data
                ol.oracle_morph(float(t[i]), s["key1"].ctypes.data, s["key2"].ctypes.data, src, 24, 24, 24, n, 1)
                d = dst[i].ctypes.data
                ol.oracle_skin_sse(src, d, src + 12, d + 12, bl + 4, bl, pp, 24, 24, 24, 24, s["blend_stride"], s["blend_stride"], s["nw"], n)

    def fn():
        t0 = time.perf_counter()
        list(pool.map(work, range(threads)))
        return time.perf_counter() - t0, 0.0
    return fn, "port", threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    subs, palettes = load_sinbad()
    nverts = sum(s["nverts"] for s in subs)
    threads = os.cpu_count() or 1
    ninst = args.instances
    fn, kind, cores = reference_step_fn(subs, palettes, ninst, threads)
    for _ in range(min(args.warmup, 2)):
        fn()
    steps = args.steps
    times = []
    t_budget = time.time() + args.ref_seconds
    for k in range(steps):
        times.append(fn())
        if time.time() > t_budget and k + 1 >= 3:          # bounded: the whole run ends within a few minutes
            break
    done = len(times)
    total = sum(x[0] for x in times)
    tpal = sum(x[1] for x in times)
    value = ninst * nverts * done / total / 1e6
    line = {
        "impl": "reference", "metric": "skinned_Mvert_per_s", "value": value, "unit": "Mvert/s", "n_gpus": args.gpus,
        "steps": done, "warmup": min(args.warmup, 2), "ms_per_step": total / done * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(ninst, nverts, args.gpus),
        "cpu_baseline": {"value": value, "unit": "Mvert/s", "cores": cores, "kind": kind,
                         "sample": "all %d instances per step (%d vertices): SkeletonInstance::setAnimationState + _getBoneMatrices per "
                                   "instance on 1 thread (%.1f ms per step), softwareVertexMorph + softwareVertexSkinning through the "
                                   "real OptimisedUtilSSE on %d threads (%.1f ms per step)" % (
                                       ninst, ninst * nverts, tpal / done * 1e3, cores, (total - tpal) / done * 1e3)},
        "e2e": {"value": value, "unit": "Mvert/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ our arm

def run_ours(args):
    import torch
    import ogre_b200
    from ogre_b200 import AnimJob, DrawDesc, SkinJob, Skeleton, check, lib
    from datagen import look_at, perspective, viewport_matrix
    from goldenlib import load_skeleton

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    subs, palettes = load_sinbad()
    nverts = sum(s["nverts"] for s in subs)
    ninst = args.instances
    nb = palettes.shape[1]
    pal_h, t_h = instance_inputs(ninst, palettes, rank)

    ctx = ogre_b200.Context(local)          # raises if the CUDA library / device is missing: no fallback
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_config_option("skin_mode", args.skin_mode)
    L = lib()

    # ---- static data: mesh streams, keyframes, one skin plan per sub-mesh (built once from the static blend stream)
    with torch.cuda.stream(stream):
        pal_sets = [torch.from_numpy(pal_h).to(dev), torch.from_numpy(pal_h).to(dev)]
        t_sets = [torch.from_numpy(t_h).to(dev), torch.from_numpy(t_h).to(dev)]
        dsub = []
        for s in subs:
            n = s["nverts"]
            e = {"key1": torch.from_numpy(s["key1"]).to(dev), "key2": torch.from_numpy(s["key2"]).to(dev),
                 "blend": torch.from_numpy(s["blend"]).to(dev), "map": torch.from_numpy(s["bone_map"].view(np.int16)).to(dev),
                 "dst": torch.empty(ninst, n, 6, device=dev)}
            dsub.append(e)
    stream.synchronize()

    def frame_jobs(b):
        arr = (AnimJob * len(dsub))()
        for k, (e, s) in enumerate(zip(dsub, subs)):
            j = SkinJob()
            j.src_pos, j.src_norm = e["key1"].data_ptr(), e["key1"].data_ptr() + 12
            j.dst_pos, j.dst_norm = e["dst"].data_ptr(), e["dst"].data_ptr() + 12
            j.blend_weight, j.blend_index = e["blend"].data_ptr() + 4, e["blend"].data_ptr()
            j.palette, j.bone_map = pal_sets[b].data_ptr(), e["map"].data_ptr()
            j.src_pos_stride = j.dst_pos_stride = j.src_norm_stride = j.dst_norm_stride = 24
            j.blend_weight_stride = j.blend_index_stride = s["blend_stride"]
            j.num_weights, j.num_matrices, j.num_vertices = s["nw"], len(s["bone_map"]), s["nverts"]
            arr[k].skin = j
            arr[k].key2 = e["key2"].data_ptr()
            arr[k].morph_out, arr[k].morph_instance_stride_bytes = None, 0
            arr[k].dst_instance_stride_bytes = s["nverts"] * 24
        return arr
    jobs = [frame_jobs(0), frame_jobs(1)]
    plans = (C.c_void_p * len(dsub))()
    if args.skin_mode == 0:
        for k in range(len(dsub)):
            p = C.c_void_p()
            check(L.ogrecu_skin_plan_create(ctx.handle, C.byref(jobs[0][k].skin), C.byref(p)))
            plans[k] = p

    def anim(b):
        # one call per frame: morph + skin of all seven sub-meshes of every instance
        check(L.ogrecu_anim_frame_dev(ctx.handle, jobs[b], plans, len(dsub), t_sets[b].data_ptr(), ninst, nb * 16))

    def step():
        anim(0)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps between a barrier+sync on both sides, CUDA events on the launch stream, max over ranks."""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        barrier()
        ms = a.elapsed_time(b)
        if dist is not None:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    # ---- device-resident number
    for _ in range(max(args.warmup, 3)):
        step()
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    launches0 = ctx.launch_count
    ms = timed(step, args.steps)
    launches = ctx.launch_count - launches0
    value = world * ninst * nverts * args.steps / (ms * 1e-3) / 1e6

    # ---- end to end: host animation states in, rendered frame out, every step.
    # Two frames in flight, as a renderer submits them: step k's states are uploaded on a copy stream into set k&1 and its
    # palettes produced right behind them (second context on that stream) while step k-1's kernels run; step k's frame is
    # copied to pinned host memory behind its render and is waited for before step k+2 may reuse the set.
    skel_def, _ = load_skeleton("sinbad_skeleton.npz")
    assert skel_def["num_bones"] == nb
    cstream = torch.cuda.Stream(device=dev)
    ctx_c = ogre_b200.Context(local)
    ctx_c.set_stream(cstream.cuda_stream)
    skel = Skeleton(ctx_c, skel_def)
    anim_h, tpos_h, w_h = instance_states(ninst, rank, skel_def["anims"][0]["length"])
    st_n = 3 * ninst * 4
    hs, ht = C.c_void_p(), C.c_void_p()
    check(L.ogrecu_host_alloc(st_n, C.byref(hs)))
    check(L.ogrecu_host_alloc(t_h.nbytes, C.byref(ht)))
    st_h = np.ctypeslib.as_array(C.cast(hs.value, C.POINTER(C.c_uint8)), shape=(st_n,))
    st_h[:ninst * 4].view(np.int32)[:] = anim_h
    st_h[ninst * 4:2 * ninst * 4].view(np.float32)[:] = tpos_h
    st_h[2 * ninst * 4:].view(np.float32)[:] = w_h
    C.memmove(ht.value, t_h.ctypes.data, t_h.nbytes)
    frame_bytes = FRAME_W * FRAME_H * 3
    hframes = [ogre_b200.pinned_empty((FRAME_H, FRAME_W, 3)) for _ in range(2)]
    with torch.cuda.stream(stream):
        st_sets = [torch.empty(st_n, dtype=torch.uint8, device=dev) for _ in range(2)]
    rts = [ogre_b200.RenderTarget(ctx, FRAME_W, FRAME_H) for _ in range(2)]
    rt_ptr = []
    for r in rts:
        p = C.c_void_p()
        check(L.ogrecu_rt_colour_device_ptr(r.handle, C.byref(p)))
        rt_ptr.append(p.value)
    # the consuming _render: one instanced, indexed draw per sub-mesh over its [instances x vertices] output buffer (what
    # TinyRenderSystem::_render is handed per sub-entity, OgreTinyRenderSystem.cpp:349-414, for every entity at once), lit,
    # depth-tested; the crowd stands on a 300 x 300 grid and fills most of the frame, seen from above
    proj = perspective(45, FRAME_W / FRAME_H, 5.0, 3000.0)
    proj[1, :] *= -1          # OgreMain flips the projection for targets with requiresTextureFlipping (Tiny windows)
    view = look_at(np.array([148.5, 340.0, 330.0]), (148.5, 0.0, 135.0))
    mvp = (proj @ view).astype(np.float32)
    mvit = np.ascontiguousarray(np.linalg.inv(view).T.astype(np.float32))
    ldir = view[:3, :3] @ np.array([0.4, 0.8, 0.3]); ldir = (ldir / np.linalg.norm(ldir)).astype(np.float32)
    # Sinbad's real triangle lists (the edge-list fixture keeps every sub-mesh's index triples in mesh order)
    shadow = np.load(os.path.join(ROOT, "tests", "golden", "sinbad_shadow.npz"))
    draws, keep_idx = [], []
    tris_per_frame = 0
    for k, (e, s) in enumerate(zip(dsub, subs)):
        g = shadow["groups"][k]
        assert int(g[4]) == s["nverts"]
        tri = np.ascontiguousarray(shadow["tris"][int(g[2]):int(g[2]) + int(g[3])].astype(np.uint16).reshape(-1))   # IT_16BIT, as Sinbad.mesh has them
        with torch.cuda.stream(stream):
            d_idx = torch.from_numpy(tri.view(np.int16)).to(dev)
        keep_idx.append(d_idx)
        d = DrawDesc()
        d.positions, d.pos_stride = e["dst"].data_ptr(), 24
        d.normals, d.normal_stride = e["dst"].data_ptr() + 12, 24
        d.indices, d.index_type, d.draw_count = d_idx.data_ptr(), 2, len(tri)
        d.vertex_count = ninst * s["nverts"]
        d.instance_count, d.instance_vertex_count = ninst, s["nverts"]        # every instance: the same indices, its own vertices
        d.prim_type = ogre_b200.PRIM_TRIANGLE_LIST
        d.flags = ogre_b200.DRAW_DEPTH_TEST | ogre_b200.DRAW_DEPTH_WRITE | ogre_b200.DRAW_LIGHTING
        d.mvp[:] = mvp.reshape(16); d.mvit[:] = mvit.reshape(16)
        d.tex_matrix[:] = np.eye(4, dtype=np.float32).reshape(16)
        d.viewport[:] = viewport_matrix(FRAME_W, FRAME_H).reshape(16)
        d.light_dir[:] = ldir
        d.ambient[:] = (0.25, 0.25, 0.3, 1.0)
        draws.append(d)
        tris_per_frame += ninst * (len(tri) // 3)
    stream.synchronize()

    def render(b):
        rts[b].clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
        for d in draws:
            rts[b].draw(d)
        rts[b].flush()

    uploaded = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_k = [0]

    # every entity's parent-node transform: the crowd stands on a 100 x 100 grid, 3 units apart (static, resident)
    world_h = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (ninst, 1))
    world_h[:, 3] = (np.arange(ninst) % 100) * 3.0
    world_h[:, 11] = (np.arange(ninst) // 100) * 3.0
    with torch.cuda.stream(cstream):
        d_world = torch.from_numpy(world_h).to(dev)
    cstream.synchronize()

    def produce_palettes(b):
        # SkeletonInstance::setAnimationState + _getBoneMatrices, then Entity::cacheBoneMatrices' concatenation with the
        # entity's world transform (OgreEntity.cpp:1307-1330), for every instance; on the copy stream (ctx_c)
        p = st_sets[b].data_ptr()
        skel.palettes_dev(ninst, 1, p, p + ninst * 4, p + 2 * ninst * 4, pal_sets[b].data_ptr(), nb * 16)
        check(L.ogrecu_concat_affine_instanced_dev(ctx_c.handle, d_world.data_ptr(), pal_sets[b].data_ptr(), pal_sets[b].data_ptr(), nb, ninst,
                                                   nb * 16))

    def step_e2e():
        k = e2e_k[0]; e2e_k[0] += 1
        b = k & 1
        # copy stream; set b was last read by step k-2, already waited for below.  (ogrecu_set_stream would drain the
        # launch stream — it is a synchronisation point — so the upload names its stream explicitly.)
        check(L.ogrecu_memcpy_h2d_on(ctx.handle, cstream.cuda_stream, st_sets[b].data_ptr(), hs.value, st_n))
        check(L.ogrecu_memcpy_h2d_on(ctx.handle, cstream.cuda_stream, t_sets[b].data_ptr(), ht.value, t_h.nbytes))
        produce_palettes(b)                               # on the copy stream (ctx_c), behind the upload
        uploaded[b].record(cstream)
        stream.wait_event(uploaded[b])
        anim(b)
        render(b)
        check(L.ogrecu_memcpy_d2h_async(ctx.handle, hframes[b].ctypes.data, rt_ptr[b], frame_bytes))
        done[b].record(stream)
        if k > 0:
            done[b ^ 1].synchronize()                    # the previous step's frame is on the host now

    def step_e2e_skin_only():
        # the same without the consuming render: one instance's skinned body sub-mesh read back as the result
        k = e2e_k[0]; e2e_k[0] += 1
        b = k & 1
        check(L.ogrecu_memcpy_h2d_on(ctx.handle, cstream.cuda_stream, st_sets[b].data_ptr(), hs.value, st_n))
        check(L.ogrecu_memcpy_h2d_on(ctx.handle, cstream.cuda_stream, t_sets[b].data_ptr(), ht.value, t_h.nbytes))
        produce_palettes(b)
        uploaded[b].record(cstream)
        stream.wait_event(uploaded[b])
        anim(b)
        check(L.ogrecu_memcpy_d2h_async(ctx.handle, hframes[b].ctypes.data, dsub[1]["dst"].data_ptr() + (ninst - 1) * subs[1]["nverts"] * 24,
                                        subs[1]["nverts"] * 24))
        done[b].record(stream)
        if k > 0:
            done[b ^ 1].synchronize()

    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    for _ in range(3):
        step_e2e()
    torch.cuda.synchronize()
    launches1 = ctx.launch_count + ctx_c.launch_count
    ms_e2e = timed(step_e2e, e2e_steps)
    launches_e2e = (ctx.launch_count + ctx_c.launch_count - launches1) / e2e_steps
    e2e_value = world * ninst * nverts * e2e_steps / (ms_e2e * 1e-3) / 1e6
    torch.cuda.synchronize()
    covered = float((hframes[(e2e_k[0] - 1) & 1] != np.array([25, 51, 76], np.uint8)).any(axis=2).mean())
    for _ in range(3):
        step_e2e_skin_only()
    torch.cuda.synchronize()
    ms_skin_only = timed(step_e2e_skin_only, args.steps)
    skin_only_value = world * ninst * nverts * args.steps / (ms_skin_only * 1e-3) / 1e6
    clk = clocks.stop() if rank == 0 else None      # sampled across the timed regions (device-resident and e2e)
    # the two stages of the e2e step alone (device, events): palettes, render
    torch.cuda.synchronize()
    pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pa.record(cstream)
    for _ in range(20):
        produce_palettes(0)
    pb.record(cstream)
    cstream.synchronize()
    ms_pal = pa.elapsed_time(pb) / 20
    render(0); torch.cuda.synchronize()
    render_stats = rts[0].stats()
    ra, rb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ra.record(stream)
    for _ in range(5):
        render(0)
    rb.record(stream)
    stream.synchronize()
    ms_render = ra.elapsed_time(rb) / 5
    # restore the fixture palettes for the measurements below (the palette kernel overwrote the sets)
    for b in (0, 1):
        with torch.cuda.stream(stream):
            pal_sets[b].copy_(torch.from_numpy(pal_h))
    stream.synchronize()

    # ---- roofline: the dominant kernel alone (the call minus its palette repack launch; palettes unchanged between calls)
    peak, peak_kind = hbm_peak()
    roof, extra = None, {}
    if rank == 0:
        step(); stream.synchronize()
        ctx.set_config_option("lane_reuse_repack", 1)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for a, b in evs:
            a.record(stream); step(); b.record(stream)
        stream.synchronize()
        ctx.set_config_option("lane_reuse_repack", 0)
        dur = float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e-3
        used = sorted(set(int(x) for s in subs for x in s["bone_map"]))
        static = sum(s["nverts"] * (48 + s["blend_stride"]) for s in subs)
        alg = ninst * nverts * 24 + ninst * len(used) * 48 + static + ninst * 4
        try:
            traf = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        except Exception:
            traf = {}
        roof = {"bound": "hbm", "kernel": "skin_lane_kernel<morph> (one launch: 7 sub-meshes x %d instances, %d vertices each)" % (ninst, nverts),
                "achieved": alg / dur / 1e9, "peak": peak,
                "peak_source": (peak_kind + " (MEASURED_PEAKS.json hbm_gbs)") if peak_kind == "measured" else "fallback",
                "unit": "GB/s", "frac": alg / dur / 1e9 / peak, "traffic": traf.get("skin_lane_frame_bytes"), "launch_ms": dur * 1e3,
                "algorithmic_bytes_per_launch": alg,
                "algorithmic_bytes": "24 B written per vertex-instance + 48 B per (instance, referenced bone) of palette + the mesh "
                                     "streams once"}
        step_bytes = alg + ninst * len(used) * 64          # + the repack's read of the palettes
        extra["step_compulsory_GBps"] = step_bytes / (ms / args.steps * 1e-3) / 1e9
        extra["step_frac_of_hbm_peak"] = extra["step_compulsory_GBps"] / peak
        extra["e2e_stages_ms"] = {"palettes_10000x69": ms_pal, "morph_skin_frame_call": ms / args.steps, "render_of_all_instances": ms_render,
                                  "frame_d2h_bytes": frame_bytes}
        extra["e2e_render"] = {"triangles_per_frame": int(tris_per_frame), "target": "%dx%d" % (FRAME_W, FRAME_H), "covered_fraction": covered,
                               "Mtri_per_s": tris_per_frame / (ms_render * 1e-3) / 1e6, "stats": render_stats}

    # ---- BASELINE's render configs (c1, c4, c5) in the same run
    render_cfg = None
    if not args.no_render:
        render_cfg = run_render_configs(args, local, rank, world, peak)

    # ---- the same render configs through the C ABI's device group: ONE process (rank 0's) drives all N GPUs, the other ranks idle
    if world > 1 and not args.no_render:
        group_cfg = run_device_group(args, rank, world, dist)
        if rank == 0 and render_cfg is not None and group_cfg:
            for key in ("c4", "c5"):
                if key in render_cfg and key in group_cfg:
                    render_cfg[key]["capi_device_group"] = group_cfg[key]

    # ---- the drop-in plugin end to end: a crowd of Entity objects through Root::renderOneFrame (rank 0, N=1: one Root per process)
    if rank == 0 and world == 1 and not args.no_plugin:
        extra["plugin_e2e"] = run_plugin_e2e(args)

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = args.ref_instances
        fn, kind, cores = reference_step_fn(subs, palettes, sample, threads)
        fn()
        best, best_pal, t_end, reps = 1e30, 0.0, time.time() + args.cpu_seconds, 0
        while time.time() < t_end or reps < 3:
            tt, tp = fn(); reps += 1
            if tt < best:
                best, best_pal = tt, tp
        cpu = {"value": sample * nverts / best / 1e6, "unit": "Mvert/s", "cores": cores, "kind": kind,
               "sample": "%d of the %d instances (%d vertices): real SkeletonInstance palettes on 1 thread (%.1f ms) + morph + skin "
                         "through the real OptimisedUtilSSE on %d threads (%.1f ms), best of %d passes" % (
                             sample, ninst, sample * nverts, best_pal * 1e3, cores, (best - best_pal) * 1e3, reps),
               "morph_skin_only_value": sample * nverts / max(best - best_pal, 1e-9) / 1e6}

    if rank == 0:
        if render_cfg is not None:
            extra["render"] = render_cfg
        line = {
            "metric": "skinned_Mvert_per_s", "value": value, "unit": "Mvert/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(ninst, nverts, world),
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "Mvert/s", "h2d_bytes_per_step": int(st_n + t_h.nbytes),
                    "d2h_bytes_per_step": int(frame_bytes), "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                    "gpu_launches_per_step": launches_e2e,
                    "skin_only_value": skin_only_value, "skin_only_ms_per_step": ms_skin_only / args.steps,
                    "note": "host animation states (12 B per instance) + morph weights in -> bone palettes on the device -> morph + skin "
                            "of every instance (buffers stay in HBM) -> the consuming _render of all instances (%d triangles into a "
                            "%dx%d target) -> the frame read back to pinned host memory; two frames in flight.  skin_only_value = the "
                            "same without the render (one instance's vertices read back)" % (tris_per_frame, FRAME_W, FRAME_H)},
            "roofline": roof, "cpu_baseline": cpu, "extra": extra,
        }
        print(json.dumps(line))
    for k in range(len(dsub)):
        if plans[k]:
            L.ogrecu_skin_plan_destroy(plans[k])
    for r in rts:
        r.destroy()
    ctx.set_stream(None)
    skel.destroy()
    ctx_c.set_stream(None)
    ctx_c.close()
    for h in (hs, ht):
        L.ogrecu_host_free(h)
    if dist is not None:
        dist.destroy_process_group()


def run_device_group(args, rank, world, dist):
    """extra.render.{c4,c5}.capi_device_group (N > 1): tools/bench_group.py in a process of its own started by rank 0 — ogrecu_init_devices
    over the N GPUs of the node, config 4's views on targets placed round-robin, config 5's frame on a split target, every frame
    checked against a single-device context.  The other ranks wait on a host-side (gloo) barrier, their GPUs idle."""
    import subprocess
    out = {}
    try:
        gloo = dist.new_group(backend="gloo")
    except Exception:                                            # noqa: BLE001
        gloo = None
    if rank == 0:
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_group.py"), "--devices", ",".join(str(d) for d in range(world)),
                                "--views", str(args.views), "--c5-tris", str(args.c5_tris)], capture_output=True, text=True, timeout=420)
            for ln in r.stdout.strip().splitlines():
                if ln.startswith("{"):
                    d = json.loads(ln)
                    out["c4" if d["config"].startswith("c4") else "c5"] = d
            if not out:
                out = {"c4": {"error": (r.stderr or r.stdout)[-300:]}}
        except Exception as e:                                   # noqa: BLE001 - reported, not fatal to the bench line
            out = {"c4": {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}}
    if gloo is not None:
        dist.barrier(group=gloo)
    else:
        dist.barrier()
    return out


def run_plugin_e2e(args):
    """extra.plugin_e2e: `--plugin-entities` animated Sinbad Entity objects in a stock OGRE scene, every one re-skinned each frame by
    Entity::_updateAnimation -> Mesh::softwareVertexBlend -> OptimisedUtil (OgreEntity.cpp:747-911, OgreMesh.cpp:1792-1897) and
    drawn by Root::renderOneFrame, through RenderSystem_CUDA.so (OptimisedUtil calls queued and flushed as one launch; and unqueued)
    and through the unmodified RenderSystem_Tiny.so + OptimisedUtilSSE (the CPU baseline) in the same harness.  One process each."""
    import subprocess
    script = os.path.join(ROOT, "tests", "harness", "run_crowd.py")
    if not os.path.exists(os.path.join(ROOT, "tests", "harness", "libplugin_harness.so")):
        return {"unavailable": "tests/harness/libplugin_harness.so is not built (needs oracle/_ref: the reference's own OgreMain)"}
    out = {"scene": "%d Sinbad entities, %s target, Root::renderOneFrame per frame (animation advanced every frame)" % (
        args.plugin_entities, args.plugin_size)}
    runs = [("cuda_batched", ["--plugin", "cuda", "--batch", "1", "--frames", "10"]),
            ("cuda_unbatched", ["--plugin", "cuda", "--batch", "0", "--frames", "3"]),
            ("tiny_sse_cpu", ["--plugin", "tiny", "--frames", "2"])]
    frames = {}
    for name, argv in runs:
        npz = "/tmp/ogrecu_plugin_e2e_%s_%d.npz" % (name, os.getpid())
        try:
            r = subprocess.run([sys.executable, script, "--count", str(args.plugin_entities), "--size", args.plugin_size, "--out", npz] + argv,
                               capture_output=True, text=True, timeout=240)
            d = json.loads(r.stdout.strip().splitlines()[-1])
            out[name] = {"ms_per_frame": d["s_per_frame"] * 1e3, "Mvert_per_s": d["Mvert_per_s"], "vertices_per_frame": d["vertices_per_frame"],
                         "triangles": d["triangles"], "batches": d["batches"]}
            frames[name] = np.load(npz)["frame"]
        except Exception as e:                                   # noqa: BLE001 - reported, not fatal to the bench line
            out[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        finally:
            if os.path.exists(npz):
                os.remove(npz)
    if "cuda_batched" in frames and "cuda_unbatched" in frames:
        out["batched_frame_identical_to_unbatched"] = bool(np.array_equal(frames["cuda_batched"], frames["cuda_unbatched"]))
    if "cuda_batched" in frames and "tiny_sse_cpu" in frames:
        a, b = frames["cuda_batched"].astype(np.int16), frames["tiny_sse_cpu"].astype(np.int16)
        diff = np.abs(a - b).max(axis=2)
        out["vs_tiny_frame"] = {"pixels_differing": int((diff > 0).sum()), "max_abs_diff": int(diff.max()),
                                "note": "Tiny skins with OptimisedUtilSSE whose normals use rsqrtps (approximate): a lit pixel may "
                                        "differ by one 8-bit step; positions / coverage are bit-identical (tests/test_plugin_gpu.py)"}
    return out


def run_render_configs(args, local, rank, world, peak):
    """BASELINE.json's render configs through tools/bench_configs.py (same process, own context and stream): c1 ogrehead
    1280x720, c4 batched offscreen views (view i -> GPU i mod G, no collective), c5 one 4K frame of 20 M triangles
    (whole frame on one GPU; under torchrun contiguous strips per rank + NCCL gather, interleaved bands, sort-last)."""
    import ogre_b200
    import bench_configs as bc
    ctx = ogre_b200.Context(local)
    T = bc.Timer(ctx)
    a = types.SimpleNamespace(cpu=(rank == 0 and world == 1 and not args.no_cpu_baseline), views=args.views, chunk=32, contexts=1,
                              c3_tris=2000000, c4_tris=1000000, c5_tris=args.c5_tris)
    out = {}
    try:
        traf = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except Exception:
        traf = {}
    # ---- c1
    r = bc.c1(ctx, T, a)
    c1_bytes = 1280 * 720 * 3 + 70000                       # 2.76 MB colour write + 0.07 MB geometry
    out["c1"] = {"config": r["config"], "frames_per_s": r["frames_per_s_device"], "Mtri_per_s": r["Mtri_per_s_device"],
                 "ms_per_frame": r["ms_per_frame_device"], "image_identical_to_tiny": r["image_identical_to_tiny"],
                 "roofline": {"bound": "hbm", "achieved": c1_bytes / (r["ms_per_frame_device"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": c1_bytes / (r["ms_per_frame_device"] * 1e-3) / 1e9 / peak, "traffic": None,
                              "note": "latency-bound: 4 draws, 4 336 triangles"},
                 "e2e": {"value": r["frames_per_s_with_readback"], "unit": "frames/s", "d2h_bytes_per_step": 1280 * 720 * 3,
                         "ms_per_frame": r["ms_per_frame_with_readback"]},
                 "cpu_baseline": ({"value": 1e3 / r["cpu_tiny_ms_per_frame"], "unit": "frames/s", "cores": 1, "kind": "reference",
                                   "sample": r.get("cpu", "")} if "cpu_tiny_ms_per_frame" in r else None)}
    # ---- c4
    r = bc.c4(ctx, T, a)
    view_bytes = 12e6 + 16e6 + 640 * 480 * 7
    out["c4"] = {"config": r["config"], "n_gpus": world, "views_per_s": r["views_per_s_device"], "Mtri_per_s": r["Mtri_per_s_device"],
                 "ms_per_view": r["ms_per_view_device"], "sharding": r["sharding"], "covered_fraction": r["covered_fraction"],
                 "roofline": {"bound": "hbm", "achieved": r["compulsory_GBps_per_gpu"], "peak": peak, "unit": "GB/s",
                              "frac": r["compulsory_GBps_per_gpu"] / peak, "traffic": None,
                              "algorithmic_bytes_per_view": int(view_bytes)},
                 "e2e": {"value": r["views_per_s_with_readback"], "unit": "views/s", "d2h_bytes_per_step": int(r["views_per_gpu"] * 640 * 480 * 3)},
                 "cpu_baseline": ({"value": 1.0 / r["cpu_s_per_view_1_thread"], "unit": "views/s", "cores": 1, "kind": r["cpu_kind"],
                                   "sample": "1 of the %d views through TinyRenderSystem::_render" % args.views,
                                   "gpu_view_identical": r["identical_to_cpu"]} if "cpu_s_per_view_1_thread" in r else None)}
    # ---- c5
    r = bc.c5(ctx, T, a)
    if world == 1:
        out["c5"] = {"config": r["config"], "n_gpus": 1, "frames_per_s": r["frames_per_s_device"], "Mtri_per_s": r["Mtri_per_s_device"],
                     "ms_per_frame": r["ms_per_frame_device"], "covered_fraction": r["covered_fraction"],
                     "roofline": {"bound": "hbm", "achieved": r["compulsory_GBps"], "peak": peak, "unit": "GB/s", "frac": r["compulsory_GBps"] / peak,
                                  "traffic": traf.get("c5_frame_bytes"),
                                  "note": "whole frame (all kernels of the pipeline pass): algorithmic = indices + vertices + planes once; traffic = "
                                          "ncu DRAM bytes of one frame (profiles/r02e_traffic_c5.csv); the pass is bound by per-triangle "
                                          "instruction issue and per-tile latency, not by bytes"},
                     "e2e": {"value": 1e3 / r["ms_per_frame_with_readback"], "unit": "frames/s", "d2h_bytes_per_step": 3840 * 2160 * 3,
                             "ms_per_frame": r["ms_per_frame_with_readback"]},
                     "cpu_baseline": ({"value": 1.0 / r["cpu_s_per_frame_1_thread"], "unit": "frames/s", "cores": 1, "kind": r["cpu_kind"],
                                       "sample": "the whole frame through TinyRenderSystem::_render", "gpu_frame_identical": r["identical_to_cpu"]}
                                      if "cpu_s_per_frame_1_thread" in r else None)}
    else:
        best = min((r["ms_per_frame_with_composite"], "strips"), (r["interleaved"]["ms_per_frame_with_composite"], "interleaved bands"),
                   (r["sort_last"]["ms_per_frame_with_composite"], "sort-last"))
        out["c5"] = {"config": r["config"], "n_gpus": world, "frames_per_s": 1e3 / best[0], "Mtri_per_s": args.c5_tris / best[0] / 1e3,
                     "ms_per_frame": best[0], "split": best[1],
                     "strips": {k: r[k] for k in ("ms_per_frame_render_only", "ms_per_frame_with_composite", "identical_to_single_gpu_frame")},
                     "interleaved_bands": {k: r["interleaved"][k] for k in ("ms_per_frame_render_only", "ms_per_frame_with_composite",
                                                                            "identical_to_single_gpu_frame")},
                     "sort_last": {k: r["sort_last"][k] for k in ("ms_per_frame_render_only", "ms_per_frame_with_composite", "ms_composite_only",
                                                                  "identical_to_single_gpu_frame")}}
    ctx.set_stream(None)
    ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skin-mode", type=int, default=0, help="0 = lane kernel from plans (default); 1-3 = the tile kernels (A/B runs)")
    ap.add_argument("--instances", type=int, default=10000)
    ap.add_argument("--ref-instances", type=int, default=1000, help="bounded CPU sample of the instances for cpu_baseline")
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--ref-seconds", type=float, default=150.0, help="--impl reference: stop after this many seconds (>= 3 steps)")
    ap.add_argument("--e2e-steps", type=int, default=40, help="steps of the end-to-end measurement (each renders every instance)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-render", action="store_true", help="skip extra.render (c1, c4, c5)")
    ap.add_argument("--no-plugin", action="store_true", help="skip extra.plugin_e2e (the Entity crowd through Root::renderOneFrame)")
    ap.add_argument("--plugin-entities", type=int, default=1000)
    ap.add_argument("--plugin-size", default="1280x720")
    ap.add_argument("--views", type=int, default=1024)
    ap.add_argument("--c5-tris", type=int, default=20000000)
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: what libraries write to file descriptor 1 on their own (NCCL's version
    # banner under torchrun) goes to stderr; print() keeps the real stdout
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        real_stdout.flush()


if __name__ == "__main__":
    main()
