#!/usr/bin/env python3
"""bench.py — headline benchmark of the OGRE software-rendering hot path on B200.

Contract (driver):  python bench.py --gpus N --steps K --warmup W [--impl reference]
  N>1 is launched under torch.distributed.run (one rank per GPU, NCCL for barrier/max only —
  the path shards by instance with NO data-path collective, so scaling is "weak").

Workload (BASELINE.json configs[1]): Sinbad (7 sub-meshes, 5 534 vertices, 69 bones, 1/4/3/1/1/3/4
weights per vertex; the real mesh streams and 32 real animation poses from tests/golden/sinbad_skin.npz)
x 10 000 animated instances per GPU.  One STEP = one frame of vertex animation for every instance,
in the order Entity::updateAnimation runs it (OgreMain/src/OgreEntity.cpp:747-911):
  softwareVertexMorph   two shared keyframes -> the instance's animated source (pos+normal, stride 24)
  softwareVertexSkinning instance source + instance bone palette -> instance output (pos+normal)
metric = skinned vertices per second (Mvert/s), whole job.

Lines printed (rank 0, one JSON line):
  value      device-timed (CUDA events on the launch stream), inputs resident in HBM
  e2e        same step through the C ABI with HOST inputs: every step uploads all instance bone
             palettes + morph weights from pinned host memory and reads one instance's skinned
             vertices back (the skinned buffers themselves stay in HBM for _render)
  roofline   dominant kernel of the timed step (the 4-weight body sub-mesh launch) vs measured HBM peak;
             extra.roofline_other_form carries the same for the other --form; extra.roofline_stream_case is SURVEY §8d's
             synthetic roofline stream (one call, 2^24 distinct vertices, 68 B per vertex) measured the same way
  cpu_baseline  the unmodified reference (real OptimisedUtilSSE via oracle/_ref) on the host cores
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FIXTURE = os.path.join(ROOT, "tests", "golden", "sinbad_skin.npz")
FALLBACK_HBM_GBS = 6650.0     # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


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


# ------------------------------------------------------------------------------ reference arm

def reference_step_fn(subs, palettes, ninst, threads):
    """Returns (fn, kind, cores): fn() runs one morph+skin pass over `ninst` instances on the CPU with
    the unmodified reference (oracle/_ref) — or the oracle port when that is absent — and returns seconds."""
    from oracle import ref
    pal, t = instance_inputs(ninst, palettes, 0)
    nb = pal.shape[1]
    bufs = []
    for s in subs:
        n = s["nverts"]
        bufs.append((np.zeros((ninst, n, 6), np.float32), np.zeros((ninst, n, 6), np.float32)))
    if ref.available():
        l = ref.lib()

        def fn():
            tot = 0.0
            for s, (tmp, dst) in zip(subs, bufs):
                bl = s["blend"].ctypes.data
                tot += l.ref_bench_anim_sub(s["key1"].ctypes.data, s["key2"].ctypes.data, t.ctypes.data, bl + 4, bl,
                                            s["blend_stride"], s["blend_stride"], s["nw"], s["nverts"], pal.ctypes.data, nb * 16,
                                            s["bone_map"].ctypes.data, len(s["bone_map"]), ninst, tmp.ctypes.data, dst.ctypes.data,
                                            threads, 1)
            return tot
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
        return time.perf_counter() - t0
    return fn, "port", threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    subs, palettes = load_sinbad()
    nverts = sum(s["nverts"] for s in subs)
    threads = os.cpu_count() or 1
    ninst = args.ref_instances
    fn, kind, cores = reference_step_fn(subs, palettes, ninst, threads)
    for _ in range(args.warmup):
        fn()
    times = [fn() for _ in range(args.steps)]
    total = sum(times)
    value = ninst * nverts * args.steps / total / 1e6
    line = {
        "impl": "reference", "metric": "skinned_Mvert_per_s", "value": value, "unit": "Mvert/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(ninst, nverts, sample=True),
        "cpu_baseline": {"value": value, "unit": "Mvert/s", "cores": cores, "kind": kind,
                         "sample": "%d of 10000 instances per step (%d vertices), morph+skin, real OptimisedUtilSSE" % (ninst, ninst * nverts)},
        "e2e": {"value": value, "unit": "Mvert/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(ninst, nverts, sample=False):
    return {"workload": "config[1] skeletal skinning: Sinbad (7 sub-meshes, %d vertices, 69 bones, 1/4/3/1/1/3/4 weights) x %d animated "
                        "instances%s; step = softwareVertexMorph (2 shared keyframes -> instance source) + softwareVertexSkinning "
                        "(pos+normal shared-packed, per-instance palette)" % (nverts, ninst, " (bounded CPU sample of the 10000)" if sample else " per GPU"),
            "instances_per_gpu": ninst, "vertices_per_instance": nverts,
            "l2": "working set per step 2.7 GB per GPU (instance source + output) >> 126 MB L2: inputs larger than L2, no flush needed"}


# ------------------------------------------------------------------------------------ our arm

def run_ours(args):
    import torch
    import ogre_b200
    from ogre_b200 import AnimJob, SkinJob, check, lib

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
    ctx.set_config_option("skin_ipc", args.skin_ipc)
    L = lib()

    with torch.cuda.stream(stream):
        d_pal = torch.from_numpy(pal_h).to(dev)
        d_t = torch.from_numpy(t_h).to(dev)
        dsub = []
        for s in subs:
            n = s["nverts"]
            e = {"key1": torch.from_numpy(s["key1"]).to(dev), "key2": torch.from_numpy(s["key2"]).to(dev),
                 "blend": torch.from_numpy(s["blend"]).to(dev), "map": torch.from_numpy(s["bone_map"].view(np.int16)).to(dev),
                 "tmp": torch.empty(ninst, n, 6, device=dev), "dst": torch.empty(ninst, n, 6, device=dev)}
            j = SkinJob()
            j.src_pos, j.src_norm = e["tmp"].data_ptr(), e["tmp"].data_ptr() + 12
            j.dst_pos, j.dst_norm = e["dst"].data_ptr(), e["dst"].data_ptr() + 12
            j.blend_weight, j.blend_index = e["blend"].data_ptr() + 4, e["blend"].data_ptr()
            j.palette, j.bone_map = d_pal.data_ptr(), e["map"].data_ptr()
            j.src_pos_stride = j.dst_pos_stride = j.src_norm_stride = j.dst_norm_stride = 24
            j.blend_weight_stride = j.blend_index_stride = s["blend_stride"]
            j.num_weights, j.num_matrices, j.num_vertices = s["nw"], len(s["bone_map"]), n
            e["job"] = j
            jf = SkinJob.from_buffer_copy(j)
            jf.src_pos, jf.src_norm = e["key1"].data_ptr(), e["key1"].data_ptr() + 12
            e["job_fused"] = jf
            dsub.append(e)
    stream.synchronize()

    def morph(e, s):
        check(L.ogrecu_morph_instanced_dev(ctx.handle, d_t.data_ptr(), e["key1"].data_ptr(), e["key2"].data_ptr(), e["tmp"].data_ptr(),
                                           24, 24, 24, s["nverts"], 1, ninst, s["nverts"] * 24))

    def skin(e, s):
        check(L.ogrecu_skin_instanced_dev(ctx.handle, C.byref(e["job"]), ninst, nb * 16, s["nverts"] * 24, s["nverts"] * 24))

    def fused(e, s):
        # one pass per sub-mesh: source_i = morph(key1, key2, t[i]) in registers -> skin -> dst_i; no intermediate in HBM
        j = e["job_fused"]
        check(L.ogrecu_anim_instanced_dev(ctx.handle, C.byref(j), e["key2"].data_ptr(), d_t.data_ptr(), None, 0, ninst, nb * 16,
                                          s["nverts"] * 24))

    def step_separate():
        for e, s in zip(dsub, subs):
            morph(e, s)
            skin(e, s)

    def anim_jobs(which):
        arr = (AnimJob * len(dsub))()
        for k, (e, s) in enumerate(zip(dsub, subs)):
            arr[k].skin = e[which] if not isinstance(e[which], list) else e[which][0]
            arr[k].key2 = e["key2"].data_ptr()
            arr[k].morph_out, arr[k].morph_instance_stride_bytes = None, 0
            arr[k].dst_instance_stride_bytes = s["nverts"] * 24
        return arr
    frame_jobs = anim_jobs("job_fused")

    def step_fused():
        # one call per frame: the fused morph+skin of all seven sub-meshes (ogrecu_anim_instanced_multi_dev overlaps the
        # independent launches internally)
        check(L.ogrecu_anim_instanced_multi_dev(ctx.handle, frame_jobs, len(dsub), d_t.data_ptr(), ninst, nb * 16))

    step = step_fused if args.form == "fused" else step_separate

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

    # ---- end to end: host palettes + weights in, one instance's skinned vertices out, every step.
    # Two frames in flight, as a renderer submits them: step k's inputs are uploaded on a copy stream into palette set
    # k&1 while step k-1's kernels run; step k's result is copied to pinned host memory behind its kernels and is
    # waited for before step k+2's inputs may overwrite the same set.  Every step still moves all of its inputs H2D
    # and its result D2H inside the timed region.  The strictly serial form (upload -> kernels -> read-back, one
    # stream) is timed as well and reported as e2e.serial_value.
    hp, ht = C.c_void_p(), C.c_void_p()
    check(L.ogrecu_host_alloc(pal_h.nbytes, C.byref(hp)))
    check(L.ogrecu_host_alloc(t_h.nbytes, C.byref(ht)))
    out_bytes = subs[1]["nverts"] * 24                # body sub-mesh (1 888 vertices, 4 weights) of the last instance
    houts = [C.c_void_p(), C.c_void_p()]
    for h in houts:
        check(L.ogrecu_host_alloc(out_bytes, C.byref(h)))
    C.memmove(hp.value, pal_h.ctypes.data, pal_h.nbytes)
    C.memmove(ht.value, t_h.ctypes.data, t_h.nbytes)
    cstream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        pal_sets = [d_pal, torch.empty_like(d_pal)]
        t_sets = [d_t, torch.empty_like(d_t)]
    for e in dsub:
        e["jobs_fused"], e["jobs"] = [e["job_fused"]], [e["job"]]
        for name in ("jobs_fused", "jobs"):
            j2 = SkinJob.from_buffer_copy(e[name][0])
            j2.palette = pal_sets[1].data_ptr()
            e[name].append(j2)
    uploaded = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    last = dsub[1]

    set_jobs = []
    for b in (0, 1):
        arr = anim_jobs("job_fused")
        for k, e in enumerate(dsub):
            arr[k].skin = e["jobs_fused"][b]
        set_jobs.append(arr)

    def step_set(b):
        if args.form == "fused":
            check(L.ogrecu_anim_instanced_multi_dev(ctx.handle, set_jobs[b], len(dsub), t_sets[b].data_ptr(), ninst, nb * 16))
            return
        for e, s in zip(dsub, subs):
            check(L.ogrecu_morph_instanced_dev(ctx.handle, t_sets[b].data_ptr(), e["key1"].data_ptr(), e["key2"].data_ptr(),
                                               e["tmp"].data_ptr(), 24, 24, 24, s["nverts"], 1, ninst, s["nverts"] * 24))
            check(L.ogrecu_skin_instanced_dev(ctx.handle, C.byref(e["jobs"][b]), ninst, nb * 16, s["nverts"] * 24, s["nverts"] * 24))

    e2e_k = [0]

    def step_e2e():
        k = e2e_k[0]; e2e_k[0] += 1
        b = k & 1
        # copy stream; set b was last read by step k-2, already waited for below.  (ogrecu_set_stream would drain the
        # launch stream — it is a synchronisation point — so the upload names its stream explicitly.)
        check(L.ogrecu_memcpy_h2d_on(ctx.handle, cstream.cuda_stream, pal_sets[b].data_ptr(), hp.value, pal_h.nbytes))
        check(L.ogrecu_memcpy_h2d_on(ctx.handle, cstream.cuda_stream, t_sets[b].data_ptr(), ht.value, t_h.nbytes))
        uploaded[b].record(cstream)
        stream.wait_event(uploaded[b])
        step_set(b)
        check(L.ogrecu_memcpy_d2h_async(ctx.handle, houts[b].value, last["dst"].data_ptr() + (ninst - 1) * out_bytes, out_bytes))
        done[b].record(stream)
        if k > 0:
            done[b ^ 1].synchronize()                    # the previous step's result is on the host now

    def step_e2e_serial():
        check(L.ogrecu_memcpy_h2d(ctx.handle, d_pal.data_ptr(), hp.value, pal_h.nbytes))
        check(L.ogrecu_memcpy_h2d(ctx.handle, d_t.data_ptr(), ht.value, t_h.nbytes))
        step()
        # device -> pinned host, synchronous: the step's result is on the host when this returns
        check(L.ogrecu_memcpy_d2h(ctx.handle, houts[0].value, last["dst"].data_ptr() + (ninst - 1) * out_bytes, out_bytes))

    # ---- end to end with the bone palettes PRODUCED ON THE DEVICE: the host supplies only each
    # instance's animation state (animation index, time position, weight: 12 B) and morph weight; ogrecu_skel_palettes_dev
    # does SkeletonInstance::setAnimationState + _getBoneMatrices for all instances, the skinning kernels read the result.
    from ogre_b200 import Skeleton
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from goldenlib import load_skeleton
    skel_def, _ = load_skeleton("sinbad_skeleton.npz")
    assert skel_def["num_bones"] == nb
    # the skeleton lives in a second context whose stream is the copy stream: the palettes of step k+1 are produced right
    # behind their state upload while the launch stream still runs step k's skinning (the palette kernel is latency-bound
    # and overlaps almost for free)
    ctx_c = ogre_b200.Context(local)
    ctx_c.set_stream(cstream.cuda_stream)
    skel = Skeleton(ctx_c, skel_def)
    st_n = 3 * ninst * 4
    hs = C.c_void_p()
    check(L.ogrecu_host_alloc(st_n, C.byref(hs)))
    st_h = np.ctypeslib.as_array(C.cast(hs.value, C.POINTER(C.c_uint8)), shape=(st_n,))
    dance_len = np.float32(skel_def["anims"][0]["length"])
    st_h[:ninst * 4].view(np.int32)[:] = 0                                                   # animation 0 ("Dance")
    st_h[ninst * 4:2 * ninst * 4].view(np.float32)[:] = np.mod((np.arange(ninst) * 0.01 + rank).astype(np.float32), dance_len)
    st_h[2 * ninst * 4:].view(np.float32)[:] = 1.0
    with torch.cuda.stream(stream):
        st_sets = [torch.empty(st_n, dtype=torch.uint8, device=dev) for _ in range(2)]

    def produce_palettes(b):
        p = st_sets[b].data_ptr()
        skel.palettes_dev(ninst, 1, p, p + ninst * 4, p + 2 * ninst * 4, pal_sets[b].data_ptr(), nb * 16)

    st_k = [0]

    def step_states():
        k = st_k[0]; st_k[0] += 1
        b = k & 1
        check(L.ogrecu_memcpy_h2d_on(ctx.handle, cstream.cuda_stream, st_sets[b].data_ptr(), hs.value, st_n))
        check(L.ogrecu_memcpy_h2d_on(ctx.handle, cstream.cuda_stream, t_sets[b].data_ptr(), ht.value, t_h.nbytes))
        produce_palettes(b)                               # on the copy stream (ctx_c), behind the upload
        uploaded[b].record(cstream)
        stream.wait_event(uploaded[b])
        step_set(b)
        check(L.ogrecu_memcpy_d2h_async(ctx.handle, houts[b].value, last["dst"].data_ptr() + (ninst - 1) * out_bytes, out_bytes))
        done[b].record(stream)
        if k > 0:
            done[b ^ 1].synchronize()

    for _ in range(3):
        step_e2e_serial()
    ms_e2e_serial = timed(step_e2e_serial, args.steps)
    for _ in range(4):
        step_states()
    torch.cuda.synchronize()
    launches1 = ctx.launch_count + ctx_c.launch_count
    ms_states = timed(step_states, args.steps)
    launches_states = ctx.launch_count + ctx_c.launch_count - launches1
    states_value = world * ninst * nverts * args.steps / (ms_states * 1e-3) / 1e6
    torch.cuda.synchronize()
    pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pa.record(cstream)
    for _ in range(50):
        produce_palettes(0)
    pb.record(cstream)
    cstream.synchronize()
    ms_pal = pa.elapsed_time(pb) / 50
    # restore the host-palette inputs in both sets for the timing below (the palette kernel overwrote them)
    for b in (0, 1):
        check(L.ogrecu_memcpy_h2d(ctx.handle, pal_sets[b].data_ptr(), hp.value, pal_h.nbytes))
    ctx.sync()
    for _ in range(4):
        step_e2e()
    torch.cuda.synchronize()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = world * ninst * nverts * args.steps / (ms_e2e * 1e-3) / 1e6
    e2e_serial_value = world * ninst * nverts * args.steps / (ms_e2e_serial * 1e-3) / 1e6
    clk = clocks.stop() if rank == 0 else None      # sampled across the timed regions (device-resident and e2e)

    # ---- per-kernel durations (outside the timed regions) for the roofline line
    kern = {}
    if rank == 0:
        evs = []
        for rep in range(5):
            for k, (e, s) in enumerate(zip(dsub, subs)):
                for name, f in (("morph", morph), ("skin", skin), ("fused", fused)):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream); f(e, s); b.record(stream)
                    evs.append((name, k, a, b))
        stream.synchronize()
        for name, k, a, b in evs:
            kern.setdefault((name, k), []).append(a.elapsed_time(b))
    peak, peak_kind = hbm_peak()
    roof = None
    extra = {}
    if rank == 0:
        k = 1                                           # dominant launch: 4-weight body sub-mesh
        s = subs[k]
        nv = s["nverts"]
        shared_once = nv * s["blend_stride"]            # blend stream: read by every instance, L2-resident, counted once
        pal_bytes = ninst * len(s["bone_map"]) * 48     # palette rows actually referenced (3 of 4 rows, mapped bones)
        # algorithmic (compulsory DRAM) bytes per launch
        alg_sep = ninst * nv * 48 + pal_bytes + shared_once          # skin alone: 24 B/vertex source read + 24 B/vertex written
        alg_fused = ninst * nv * 24 + pal_bytes + shared_once + nv * 48 + ninst * 4   # fused: only the result is written; keys once
        try:
            traf = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        except Exception:
            traf = {}

        def roofline(name, label, alg, key):
            dur = float(np.mean(kern[(name, k)][1:])) * 1e-3
            return {"bound": "hbm", "kernel": label % (nv, ninst), "achieved": alg / dur / 1e9, "peak": peak,
                    "peak_source": (peak_kind + " (MEASURED_PEAKS.json hbm_gbs)") if peak_kind == "measured" else "fallback",
                    "unit": "GB/s", "frac": alg / dur / 1e9 / peak, "traffic": traf.get(key), "launch_ms": dur * 1e3,
                    "algorithmic_bytes_per_launch": alg}
        roof_fused = roofline("fused", "skin_pair_kernel<4,normals,fused morph> (sub-mesh 1: %d verts x %d instances, 4 weights)",
                              alg_fused, "skin_pair_fused_sub1_bytes")
        roof_sep = roofline("skin", "skin_pair_kernel<4,normals> (sub-mesh 1: %d verts x %d instances, 4 weights)", alg_sep,
                            "skin_pair_sub1_bytes")
        roof = roof_fused if args.form == "fused" else roof_sep
        extra["roofline_other_form"] = roof_sep if args.form == "fused" else roof_fused
        tot = {n: sum(np.mean(kern[(n, i)][1:]) for i in range(len(subs))) for n in ("skin", "morph", "fused")}
        extra["kernel_ms_per_step"] = tot
        extra["fused_ms_per_submesh"] = [float(np.mean(kern[("fused", i)][1:])) for i in range(len(subs))]
        extra["skin_ms_per_submesh"] = [float(np.mean(kern[("skin", i)][1:])) for i in range(len(subs))]
        extra["fused_Mvert_per_s"] = ninst * nverts / (tot["fused"] * 1e-3) / 1e6
        extra["separate_Mvert_per_s"] = ninst * nverts / ((tot["skin"] + tot["morph"]) * 1e-3) / 1e6
        extra["skin_only_Mvert_per_s"] = ninst * nverts / (tot["skin"] * 1e-3) / 1e6
        extra["morph_only_Mvert_per_s"] = ninst * nverts / (tot["morph"] * 1e-3) / 1e6
        pal_step = ninst * sum(len(s["bone_map"]) for s in subs) * 48
        bytes_step = ninst * nverts * (24 if args.form == "fused" else 72) + pal_step
        # ---- SURVEY §8d "synthetic stream (roofline case)": ONE softwareVertexSkinning call over 2^24 distinct vertices, 64
        # bones, 4 weights, uniformly random bone indices (mt19937-style: no locality), pos+normal stride 24, blend stride 20:
        # 24 + 20 + 24 = 68 algorithmic bytes per vertex.  Outside the timed regions; the kernel is skin_stream_kernel.
        try:
            with torch.cuda.stream(stream):
                sn, snb = 1 << 24, 64
                g = torch.Generator(device=dev).manual_seed(1234)
                s_src = torch.empty(sn, 6, device=dev).uniform_(-10, 10, generator=g)
                s_w = torch.rand(sn, 4, device=dev, generator=g) + 0.05
                s_w = s_w / s_w.sum(1, keepdim=True)
                s_blend = torch.empty(sn, 20, device=dev, dtype=torch.uint8)
                s_blend[:, :4] = torch.randint(0, snb, (sn, 4), device=dev, dtype=torch.uint8, generator=g)
                s_blend[:, 4:] = s_w.contiguous().view(torch.uint8).reshape(sn, 16)
                s_pal = torch.eye(4, device=dev).repeat(snb, 1, 1) + 0.01 * torch.randn(snb, 4, 4, device=dev, generator=g)
                s_pal[:, 3, :] = torch.tensor([0, 0, 0, 1.0], device=dev)
                s_pal = s_pal.contiguous()
                s_dst = torch.empty(sn, 6, device=dev)
            sj = SkinJob()
            sj.src_pos, sj.src_norm = s_src.data_ptr(), s_src.data_ptr() + 12
            sj.dst_pos, sj.dst_norm = s_dst.data_ptr(), s_dst.data_ptr() + 12
            sj.blend_weight, sj.blend_index = s_blend.data_ptr() + 4, s_blend.data_ptr()
            sj.palette, sj.bone_map = s_pal.data_ptr(), None
            sj.src_pos_stride = sj.dst_pos_stride = sj.src_norm_stride = sj.dst_norm_stride = 24
            sj.blend_weight_stride = sj.blend_index_stride = 20
            sj.num_weights, sj.num_matrices, sj.num_vertices = 4, snb, sn
            stream.synchronize()
            for _ in range(3):
                check(L.ogrecu_skin_instanced_dev(ctx.handle, C.byref(sj), 1, 0, 0, 0))
            sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
            for a, b in sev:
                a.record(stream)
                check(L.ogrecu_skin_instanced_dev(ctx.handle, C.byref(sj), 1, 0, 0, 0))
                b.record(stream)
            stream.synchronize()
            sdur = float(np.median([a.elapsed_time(b) for a, b in sev])) * 1e-3
            extra["roofline_stream_case"] = {
                "bound": "hbm", "kernel": "skin_stream_kernel<4,normals> (one call, 2^24 distinct vertices, 64 bones, random bone indices)",
                "achieved": sn * 68 / sdur / 1e9, "peak": peak, "unit": "GB/s", "frac": sn * 68 / sdur / 1e9 / peak, "launch_ms": sdur * 1e3,
                "traffic": traf.get("skin_stream_nw4_2p24_bytes"),
                "algorithmic_bytes_per_launch": sn * 68, "Mvert_per_s": sn / sdur / 1e6}
            del s_src, s_w, s_blend, s_pal, s_dst
        except Exception as ex:      # never lose the contract line over the side measurement
            extra["roofline_stream_case"] = {"error": repr(ex)}
        extra["step_compulsory_GBps"] = bytes_step / (ms / args.steps * 1e-3) / 1e9
        extra["step_frac_of_hbm_peak"] = extra["step_compulsory_GBps"] / peak

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = args.ref_instances
        fn, kind, cores = reference_step_fn(subs, palettes, sample, threads)
        fn()
        best, t_end, reps = 1e30, time.time() + args.cpu_seconds, 0
        while time.time() < t_end or reps < 3:
            best = min(best, fn()); reps += 1
        cpu = {"value": sample * nverts / best / 1e6, "unit": "Mvert/s", "cores": cores, "kind": kind,
               "sample": "%d of the %d instances (%d vertices), morph+skin through the real OptimisedUtilSSE, best of %d passes" % (
                   sample, ninst, sample * nverts, reps)}

    if rank == 0:
        line = {
            "metric": "skinned_Mvert_per_s", "value": value, "unit": "Mvert/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(ninst, nverts), form=args.form, parallelism="instances sharded per GPU, no collective (dp%d)" % world),
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "Mvert/s", "h2d_bytes_per_step": int(pal_h.nbytes + t_h.nbytes),
                    "d2h_bytes_per_step": int(out_bytes), "ms_per_step": ms_e2e / args.steps,
                    "serial_value": e2e_serial_value, "serial_ms_per_step": ms_e2e_serial / args.steps,
                    "states_value": states_value, "states_ms_per_step": ms_states / args.steps,
                    "states_h2d_bytes_per_step": int(st_n + t_h.nbytes), "states_gpu_launches": int(launches_states),
                    "states_note": "same step with the bone palettes produced on the device (ogrecu_skel_palettes_dev, %.1f us per "
                                   "launch for %d instances x %d bones): the host uploads each instance's animation state (12 B) and "
                                   "morph weight instead of its %d-byte palette" % (ms_pal * 1e3, ninst, nb, nb * 64),
                    "note": "two frames in flight (upload of step k+1 on a copy stream overlaps the kernels of step k); serial_value = "
                            "upload -> kernels -> read-back on one stream; skinned vertex buffers stay in HBM for _render, one instance "
                            "is read back per step as the result"},
            "roofline": roof, "cpu_baseline": cpu, "extra": extra,
        }
        print(json.dumps(line))
    ctx.set_stream(None)
    skel.destroy()
    ctx_c.set_stream(None)
    ctx_c.close()
    for h in [hp, ht, hs] + houts:
        L.ogrecu_host_free(h)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--form", default="fused", choices=["fused", "separate"],
                    help="fused: one morph+skin pass per sub-mesh (default); separate: softwareVertexMorph then softwareVertexSkinning")
    ap.add_argument("--skin-mode", type=int, default=0, help="kernel selection for A/B runs (ogrecu_set_config_option)")
    ap.add_argument("--skin-ipc", type=int, default=0, help="instances per CTA of the tiled kernel, 0 = auto")
    ap.add_argument("--instances", type=int, default=10000)
    ap.add_argument("--ref-instances", type=int, default=1000, help="bounded CPU sample of the instances")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
