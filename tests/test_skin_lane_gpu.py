"""GPU parity of the instance-pair-per-lane skinning kernel (ogre_b200/csrc/skin_lane.cu, reached through
ogrecu_skin_plan_create + ogrecu_anim_frame_dev) against the CPU oracle — bit patterns.

The kernel skips zero-weight slots and unit first weights where that is a bitwise no-op; the oracle (SSE formulation,
OgreOptimisedUtilSSE.cpp:278-368) multiplies everything through.  Covers: ragged instance groups (not a multiple of 64,
odd), odd vertex counts (instance rows alternate 8-byte phase: the two-tensor-map store), chunk splits by bone count,
zero weights trailing and interior, w0 == 1, -0 / denormal / non-finite palette entries, zero / tiny first weights,
a padded destination stride, layouts the lane kernel refuses (falls through to the tile kernels), and the multi-job frame.
"""
import ctypes as C

import numpy as np
import pytest

import oracle
from datagen import bits, rigid_matrices
from test_skin_tile_gpu import gpu_job, make_inputs, oracle_skin

pytestmark = pytest.mark.gpu


def localise_bones(rng, blend, n, nw, nb, window, drift):
    """Bone indices with the locality real meshes have (the plan needs it: a chunk holds at most 16 palette entries):
    vertex v draws from `window` consecutive bones starting at (v // drift) % nb."""
    base = (np.arange(n) // drift)[:, None]
    blend[:, :nw] = ((base + rng.integers(0, window, size=(n, nw))) % nb).astype(np.uint8)
    return blend


def make_plan(ctx, job):
    """ogrecu_skin_plan_create, or None when the mesh cannot be planned (no bone locality): the frame call then takes the tile kernels"""
    from ogre_b200 import OgrecuError, check, lib
    plan = C.c_void_p()
    try:
        check(lib().ogrecu_skin_plan_create(ctx.handle, C.byref(job), C.byref(plan)))
    except OgrecuError as e:
        assert e.status == 5, e
        return C.c_void_p(None)
    return plan


def sparsify_weights(rng, blend, n, nw, zero_frac, special):
    """zero weights the way real meshes have them: trailing slots exactly +0, the rest renormalised; some w0 == 1"""
    w = blend[:, 4:].view(np.float32).reshape(n, nw).copy()
    if nw > 1:
        keep = 1 + (rng.random((n, 1)) > zero_frac) * rng.integers(0, nw, size=(n, 1))
        w[np.arange(nw)[None, :] >= keep] = 0.0
        w = (w / w.sum(axis=1, keepdims=True)).astype(np.float32)
        w[np.arange(nw)[None, :] >= keep] = 0.0
        interior = rng.random(n) < 0.05                   # a zero in the middle, a non-zero after it
        if nw >= 3:
            w[interior, 1] = 0.0; w[interior, 2] = np.float32(0.25)
        neg = rng.random(n) < 0.03                        # a -0 weight is a zero too
        w[neg, nw - 1] = np.float32(-0.0)
    if special:
        wz = rng.random(n) < 0.1
        w[wz, 0] = np.array([0.0, -0.0, 1e-30, 1e-7], np.float32)[rng.integers(0, 4, size=int(wz.sum()))]
    blend[:, 4:] = w.view(np.uint8).reshape(n, 4 * nw)
    return blend


LANE_CASES = [
    # n,   nw, nb, ninst, fused, dst_off, zero_frac, special, pad, (bone window, drift) or None = no locality
    (1888, 4, 55, 70, True, 0, 0.5, False, 0, (6, 24)),     # Sinbad's body shape, ragged last instance group
    (1533, 4, 18, 129, True, 0, 0.7, False, 0, (5, 40)),    # odd vertex count (odd instances 8 mod 16: shifted quads) and odd instance count
    (583, 3, 6, 64, True, 0, 0.6, False, 0, (6, 1000)),
    (302, 1, 2, 96, True, 0, 0.0, False, 0, (2, 100)),
    (77, 2, 7, 65, True, 0, 0.5, False, 0, (7, 1000)),      # lerp form with w1 == 0
    (1001, 4, 30, 64, False, 0, 0.5, False, 0, (8, 16)),    # shared bind pose, no morph
    (640, 4, 40, 66, True, 8, 0.5, False, 0, (8, 16)),      # destination not 16-byte aligned: refused, tile kernels
    (512, 4, 24, 64, True, 0, 0.5, True, 0, (10, 32)),      # -0, denormals and non-finite entries in the palettes, zero / tiny first weights
    (512, 3, 24, 64, True, 0, 0.5, True, 0, (10, 32)),
    (512, 2, 24, 64, True, 0, 0.5, True, 0, (10, 32)),
    (300, 4, 256, 3, True, 0, 0.2, False, 0, (12, 3)),      # maximum palette, fast drift: chunks are cut by the bone limit
    (300, 4, 256, 3, True, 0, 0.2, False, 0, None),         # no locality at all: cannot be planned, tile kernels
    (255, 4, 200, 2, False, 0, 0.0, False, 40, (9, 8)),     # padded instance stride (multiple of 16), odd count, no zero weights
    (254, 4, 200, 5, True, 0, 0.3, False, 8, (9, 8)),       # padded stride 8 mod 16 with an even count: shifted quads, even tail
    (1, 4, 9, 5, True, 0, 0.0, False, 0, (9, 1)),
    (2, 3, 9, 1000, True, 0, 0.5, False, 8, (9, 1)),
    (3, 3, 9, 65, True, 0, 0.5, False, 0, (9, 1)),
    (5, 2, 9, 66, True, 0, 0.5, False, 0, (9, 1)),
    (6, 4, 9, 64, False, 0, 0.5, False, 8, (9, 1)),
    (7, 1, 9, 63, True, 0, 0.5, False, 0, (9, 1)),
    (129, 1, 3, 130, False, 0, 0.0, False, 0, (3, 50)),
    (131, 4, 12, 7, True, 0, 0.4, False, 0, (6, 20)),       # chunk boundary at 128 with an odd count
]


@pytest.mark.parametrize("case", LANE_CASES, ids=lambda c: "n%d_w%d_b%d_i%d_%s_o%d_z%d%s_p%d_%s" % (
    c[0], c[1], c[2], c[3], "F" if c[4] else "S", c[5], int(c[6] * 100), "_special" if c[7] else "", c[8],
    "loc%d_%d" % c[9] if c[9] else "noloc"))
def test_lane_kernel_equals_oracle(gpu_ctx, case):
    from ogre_b200 import AnimJob, check, lib
    n, nw, nb, ninst, fused, do, zero_frac, special, pad, loc = case
    ctx = gpu_ctx
    rng = np.random.default_rng(n * 7 + nw + ninst)
    keys, blend, pals, bone_map = make_inputs(rng, n, nw, nb, max(ninst, 2), True, True)
    pals = pals[:ninst]
    key1, key2 = np.ascontiguousarray(keys[0]), np.ascontiguousarray(keys[1])
    if loc:
        blend = localise_bones(rng, blend, n, nw, nb, *loc)
    blend = sparsify_weights(rng, blend, n, nw, zero_frac, special)
    if special:
        sp = np.array([0.0, -0.0, -1e-45, 1e-45, -3e-39, 1.1754944e-38, 3e38, -3.2e38, np.inf, -np.inf, np.nan, 1.0, -1.0], np.float32)
        pm = rng.random(pals.shape) < 0.02
        pm[: ninst // 2] &= rng.random(pals[: ninst // 2].shape) < 0.1      # some groups stay almost clean: -0 without non-finites
        pals[pm] = sp[rng.integers(0, len(sp), size=int(pm.sum()))]
    t = rng.uniform(0, 1, size=ninst).astype(np.float32)
    stride = n * 24 + pad
    dk1, dk2, dt = ctx.to_device(key1), ctx.to_device(key2), ctx.to_device(t)
    dbl, dpal, dmap = ctx.to_device(blend), ctx.to_device(pals), ctx.to_device(bone_map)
    total = ninst * stride
    ddst = ctx.alloc(total + 64)
    ddst.upload(np.full(total + 64, 0xCD, np.uint8))
    jobs = (AnimJob * 1)()
    jobs[0].skin = gpu_job(ctx, dk1.ptr, ddst.ptr + do, dbl, dpal, dmap, n, nw, nb, True)
    jobs[0].key2 = dk2.ptr if fused else None
    jobs[0].dst_instance_stride_bytes = stride
    plan = make_plan(ctx, jobs[0].skin)
    assert bool(plan.value) == bool(loc), "plannable exactly when the bones have locality"
    plans = (C.c_void_p * 1)(plan)
    launches0 = ctx.launch_count
    try:
        check(lib().ogrecu_anim_frame_dev(ctx.handle, jobs, plans, 1, dt.ptr if fused else None, ninst, (nb + 2) * 16))
        ctx.sync()
    finally:
        check(lib().ogrecu_skin_plan_destroy(plan))
    if loc and do == 0 and ninst >= 2:
        assert ctx.launch_count - launches0 == 2, "palette repack + the lane kernel"
    raw = ddst.download(np.uint8)
    assert (raw[:do] == 0xCD).all() and (raw[do + total:] == 0xCD).all(), "wrote outside the destination"
    rows = raw[do:do + total].reshape(ninst, stride)
    if pad:
        assert (rows[:, n * 24:] == 0xCD).all(), "wrote into the padding between instances"
    got = np.ascontiguousarray(rows[:, :n * 24]).view(np.float32).reshape(ninst, n, 6)
    for i in range(ninst):
        src = key1
        if fused:
            src = np.zeros((n, 6), np.float32)
            oracle.lib().oracle_morph(float(t[i]), key1.ctypes.data, key2.ctypes.data, src.ctypes.data, 24, 24, 24, n, 1)
        ref = oracle_skin(src, blend, np.ascontiguousarray(pals[i][bone_map]), nw, n, True)
        ok = ~np.isnan(ref)
        assert np.array_equal(np.isnan(got[i]), np.isnan(ref)), (case, i)     # NaN payloads are not part of the contract
        assert np.array_equal(bits(got[i])[ok], bits(ref)[ok]), (case, i)
    for b in (dk1, dk2, dt, dbl, dpal, dmap, ddst):
        b.free()


def test_lane_frame_of_many_sub_meshes_equals_tile_kernels(gpu_ctx):
    """ogrecu_anim_frame_dev with plans (one launch over all sub-meshes; more than 8 jobs -> two launches) ==
    ogrecu_anim_instanced_dev per sub-mesh through the tile kernels, bit for bit."""
    from ogre_b200 import AnimJob, check, lib
    ctx = gpu_ctx
    ninst, nb = 203, 40
    shapes = [(1888, 4), (583, 3), (302, 1), (77, 2), (1533, 4), (288, 1), (252, 3), (688, 1), (64, 4), (31, 2)]
    rng = np.random.default_rng(5)
    t = rng.uniform(0, 1, size=ninst).astype(np.float32)
    dt = ctx.to_device(t)
    pals = np.stack([rigid_matrices(rng, nb) for _ in range(ninst)])
    dpal = ctx.to_device(pals)
    arr = (AnimJob * len(shapes))()
    plans = (C.c_void_p * len(shapes))()
    keep, outs = [], []
    for k, (n, nw) in enumerate(shapes):
        keys, blend, _, bone_map = make_inputs(rng, n, nw, nb - 2, 2, True, True)
        blend = sparsify_weights(rng, localise_bones(rng, blend, n, nw, nb - 2, 7, 30), n, nw, 0.5, False)
        dk1, dk2, dbl, dmap = ctx.to_device(np.ascontiguousarray(keys[0])), ctx.to_device(np.ascontiguousarray(keys[1])), ctx.to_device(blend), ctx.to_device(bone_map)
        da, db = ctx.alloc(ninst * n * 24), ctx.alloc(ninst * n * 24)
        arr[k].skin = gpu_job(ctx, dk1.ptr, da.ptr, dbl, dpal, dmap, n, nw, nb - 2, True)
        arr[k].key2 = dk2.ptr
        arr[k].dst_instance_stride_bytes = n * 24
        plans[k] = make_plan(ctx, arr[k].skin)
        assert plans[k]
        jb = gpu_job(ctx, dk1.ptr, db.ptr, dbl, dpal, dmap, n, nw, nb - 2, True)
        keep += [dk1, dk2, dbl, dmap, da, db]
        outs.append((da, db, jb, dk2, n))
    launches0 = ctx.launch_count
    check(lib().ogrecu_anim_frame_dev(ctx.handle, arr, plans, len(shapes), dt.ptr, ninst, nb * 16))
    assert ctx.launch_count - launches0 == 3, "one palette repack, ten eligible jobs in two lane launches"
    for da, db, jb, dk2, n in outs:
        check(lib().ogrecu_anim_instanced_dev(ctx.handle, C.byref(jb), dk2.ptr, dt.ptr, None, 0, ninst, nb * 16, n * 24))
    ctx.sync()
    for da, db, jb, dk2, n in outs:
        assert np.array_equal(da.download(np.uint32), db.download(np.uint32)), n
    for k in range(len(shapes)):
        check(lib().ogrecu_skin_plan_destroy(plans[k]))
    for b in keep + [dt, dpal]:
        b.free()


def test_plan_rejects_bad_input(gpu_ctx):
    from ogre_b200 import OgrecuError, check, lib
    ctx = gpu_ctx
    rng = np.random.default_rng(1)
    _, blend, pals, bone_map = make_inputs(rng, 64, 4, 9, 2, True, True)
    blend[10, 2] = 200                                 # a blend index past the palette
    dbl, dpal, dmap = ctx.to_device(blend), ctx.to_device(pals), ctx.to_device(bone_map)
    j = gpu_job(ctx, dbl.ptr, dbl.ptr, dbl, dpal, dmap, 64, 4, 9, True)
    plan = C.c_void_p()
    with pytest.raises(OgrecuError):
        check(lib().ogrecu_skin_plan_create(ctx.handle, C.byref(j), C.byref(plan)))
    for b in (dbl, dpal, dmap):
        b.free()
