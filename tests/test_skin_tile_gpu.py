"""GPU parity of the tiled instanced skinning kernels (ogre_b200/csrc/skin_pair.cu: two instances per pass,
packed f32x2 arithmetic; ogre_b200/csrc/skin_tile.cu: one instance per pass) in all their transfer modes, and of
the fused morph+skin pass, against the CPU oracle — bit patterns.

Covers: full / partial / odd-sized tiles (odd vertex counts break the 16-byte TMA rule and must
take the cooperative path), 8- and 4-byte misaligned bases, shared bind pose vs per-instance
source, positions only, bone_map gather, instances-per-CTA > 1 with ragged last group, and the
four kernel selections of the "skin_mode" option (0 pair kernel, 1 without TMA, 2 one thread per vertex,
3 single-instance tile kernel), and signed zeros / denormals / extreme magnitudes through the packed arithmetic.
"""
import ctypes as C

import numpy as np
import pytest

import oracle
from datagen import bits, rigid_matrices

pytestmark = pytest.mark.gpu


def make_inputs(rng, n, nw, nb, ninst, norm, per_instance_src):
    vs = 24 if norm else 12
    nsrc = ninst if per_instance_src else 1
    src = rng.uniform(-10, 10, size=(nsrc, n, vs // 4)).astype(np.float32)
    if norm:
        nn = rng.normal(size=(nsrc, n, 3))
        src[:, :, 3:] = (nn / np.linalg.norm(nn, axis=2, keepdims=True)).astype(np.float32)
    w = rng.uniform(0.05, 1, size=(n, nw))
    w = (w / w.sum(axis=1, keepdims=True)).astype(np.float32)
    idx = rng.integers(0, nb, size=(n, nw)).astype(np.uint8)
    bs = 4 + 4 * nw
    blend = np.zeros((n, bs), np.uint8)
    blend[:, :nw] = idx
    blend[:, 4:] = w.view(np.uint8).reshape(n, 4 * nw)
    pals = np.stack([rigid_matrices(rng, nb + 2) for _ in range(ninst)])     # [ninst, nb+2, 16]
    bone_map = rng.permutation(nb + 2)[:nb].astype(np.uint16)
    return src, blend, pals, bone_map


def oracle_skin(src_i, blend, mats, nw, n, norm):
    vs = 24 if norm else 12
    out = np.zeros((n, vs // 4), np.float32)
    s = np.ascontiguousarray(src_i)
    sp, dp = s.ctypes.data, out.ctypes.data
    sn, dn = (sp + 12, dp + 12) if norm else (None, None)
    strides = (vs, vs, vs if norm else 0, vs if norm else 0, blend.shape[1], blend.shape[1])
    oracle.skin("sse", sp, dp, sn, dn, blend.ctypes.data + 4, blend.ctypes.data, mats, strides, nw, n)
    return out


def gpu_job(ctx, dsrc_ptr, ddst_ptr, dbl, dpal, dmap, n, nw, nb, norm):
    from ogre_b200 import SkinJob
    vs = 24 if norm else 12
    j = SkinJob()
    j.src_pos, j.dst_pos = dsrc_ptr, ddst_ptr
    if norm:
        j.src_norm, j.dst_norm = dsrc_ptr + 12, ddst_ptr + 12
        j.src_norm_stride = j.dst_norm_stride = vs
    j.src_pos_stride = j.dst_pos_stride = vs
    j.blend_weight, j.blend_index = dbl.ptr + 4, dbl.ptr
    j.blend_weight_stride = j.blend_index_stride = 4 + 4 * nw
    j.palette, j.bone_map = dpal.ptr, dmap.ptr
    j.num_weights, j.num_matrices, j.num_vertices = nw, nb, n
    return j


CASES = [
    # n,   nw, nb, ninst, norm, per_inst_src, src_off, dst_off, ipc
    (1888, 4, 55, 7, True, True, 0, 0, 3),      # TMA both ways, ragged last instance group
    (1888, 4, 55, 5, True, False, 0, 0, 0),     # shared bind pose
    (1001, 3, 20, 4, True, True, 0, 0, 2),      # odd n: instance bases alternate 8-byte misalignment -> cooperative 8
    (1001, 2, 9, 3, True, True, 8, 8, 0),       # 8-byte misaligned bases
    (513, 1, 3, 3, True, True, 4, 12, 2),       # 4-byte misaligned bases -> cooperative 4
    (256, 4, 64, 2, True, True, 0, 0, 0),       # exactly one full tile
    (255, 4, 200, 2, True, True, 0, 0, 0),      # partial tile, odd, big palette (> 48 KB smem path)
    (1, 2, 2, 3, True, True, 0, 0, 3),
    (777, 4, 30, 6, False, True, 0, 0, 4),      # positions only
    (1024, 3, 30, 3, False, False, 4, 4, 0),    # positions only, misaligned, shared source
    (300, 4, 256, 2, True, True, 0, 0, 0),      # maximum palette
    (600, 4, 100, 9, True, True, 0, 0, 0),      # palette of more than 64 matrices, odd instance count
    (1533, 4, 18, 11, True, True, 0, 0, 6),     # odd n and odd instances: half-empty last pass of every CTA
]


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_w%d_b%d_i%d_%s%s_o%d_%d_ipc%d" % (
    c[0], c[1], c[2], c[3], "pn" if c[4] else "p", "I" if c[5] else "S", c[6], c[7], c[8]))
def test_instanced_skinning_all_modes(gpu_ctx, case, mode):
    from ogre_b200 import check, lib
    n, nw, nb, ninst, norm, per_inst, so, do, ipc = case
    ctx = gpu_ctx
    vs = 24 if norm else 12
    rng = np.random.default_rng(n * 31 + nw)
    src, blend, pals, bone_map = make_inputs(rng, n, nw, nb, ninst, norm, per_inst)
    dsrc = ctx.alloc(src.nbytes + 64); dsrc.upload(src.reshape(-1).view(np.uint8), so)
    ddst = ctx.alloc(ninst * n * vs + 64)
    guard = np.full(ninst * n * vs + 64, 0xCD, np.uint8)
    ddst.upload(guard)
    dbl, dpal, dmap = ctx.to_device(blend), ctx.to_device(pals), ctx.to_device(bone_map)
    j = gpu_job(ctx, dsrc.ptr + so, ddst.ptr + do, dbl, dpal, dmap, n, nw, nb, norm)
    ctx.set_config_option("skin_mode", mode)
    ctx.set_config_option("skin_ipc", ipc)
    try:
        check(lib().ogrecu_skin_instanced_dev(ctx.handle, C.byref(j), ninst, (nb + 2) * 16, n * vs if per_inst else 0, n * vs))
        ctx.sync()
    finally:
        ctx.set_config_option("skin_mode", 0)
        ctx.set_config_option("skin_ipc", 0)
    raw = ddst.download(np.uint8)
    assert (raw[:do] == 0xCD).all() and (raw[do + ninst * n * vs:] == 0xCD).all(), "wrote outside the destination"
    got = raw[do:do + ninst * n * vs].view(np.float32).reshape(ninst, n, vs // 4)
    for i in range(ninst):
        ref = oracle_skin(src[i if per_inst else 0], blend, np.ascontiguousarray(pals[i][bone_map]), nw, n, norm)
        assert np.array_equal(bits(got[i]), bits(ref)), (case, mode, i)
    for b in (dsrc, ddst, dbl, dpal, dmap):
        b.free()


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("norm", [True, False])
@pytest.mark.parametrize("n,nw,nb,ninst,ipc,want_morph", [(1888, 4, 55, 9, 4, True), (1001, 3, 18, 5, 0, False),
                                                           (302, 1, 2, 33, 0, True), (77, 2, 7, 2, 1, False),
                                                           (583, 3, 6, 7, 0, True)])
def test_fused_morph_skin_equals_morph_then_skin(gpu_ctx, norm, n, nw, nb, ninst, ipc, want_morph, mode):
    from ogre_b200 import check, lib
    ctx = gpu_ctx
    vs = 24 if norm else 12
    rng = np.random.default_rng(n + 1000 * nw)
    keys, blend, pals, bone_map = make_inputs(rng, n, nw, nb, ninst, norm, True)
    key1, key2 = np.ascontiguousarray(keys[0]), np.ascontiguousarray(keys[1 % len(keys)] if ninst > 1 else keys[0] * 0.9)
    t = rng.uniform(0, 1, size=ninst).astype(np.float32)
    dk1, dk2, dt = ctx.to_device(key1), ctx.to_device(key2), ctx.to_device(t)
    dbl, dpal, dmap = ctx.to_device(blend), ctx.to_device(pals), ctx.to_device(bone_map)
    ddst = ctx.alloc(ninst * n * vs)
    dmo = ctx.alloc(ninst * n * vs) if want_morph else None
    j = gpu_job(ctx, dk1.ptr, ddst.ptr, dbl, dpal, dmap, n, nw, nb, norm)
    ctx.set_config_option("skin_ipc", ipc)
    ctx.set_config_option("skin_mode", mode)
    try:
        check(lib().ogrecu_anim_instanced_dev(ctx.handle, C.byref(j), dk2.ptr, dt.ptr, dmo.ptr if dmo else None, n * vs, ninst,
                                              (nb + 2) * 16, n * vs))
        ctx.sync()
    finally:
        ctx.set_config_option("skin_ipc", 0)
        ctx.set_config_option("skin_mode", 0)
    got = ddst.download(np.float32).reshape(ninst, n, vs // 4)
    gmo = dmo.download(np.float32).reshape(ninst, n, vs // 4) if dmo else None
    for i in range(ninst):
        mo = np.zeros((n, vs // 4), np.float32)
        oracle.lib().oracle_morph(float(t[i]), key1.ctypes.data, key2.ctypes.data, mo.ctypes.data, vs, vs, vs, n, int(norm))
        ref = oracle_skin(mo, blend, np.ascontiguousarray(pals[i][bone_map]), nw, n, norm)
        assert np.array_equal(bits(got[i]), bits(ref)), (n, nw, i)
        if gmo is not None:
            assert np.array_equal(bits(gmo[i]), bits(mo))
    for b in (dk1, dk2, dt, dbl, dpal, dmap, ddst, dmo):
        if b is not None:
            b.free()


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("nw", [1, 2, 3, 4])
def test_special_values_through_packed_arithmetic(gpu_ctx, nw, mode):
    """Signed zeros, denormals, tiny and huge magnitudes (no overflow, so no NaN payload questions): the packed
    multiply is fma(a, b, -0.0) and must round exactly like the scalar multiply the oracle does."""
    from ogre_b200 import check, lib
    ctx = gpu_ctx
    n, nb, ninst, norm = 512, 16, 4, True
    rng = np.random.default_rng(77 + nw)
    src, blend, pals, bone_map = make_inputs(rng, n, nw, nb, ninst, norm, True)
    special = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-39, -3e-39, 1.1754944e-38, 1e-30, -1e-25, 1e15, -3e15, 1.0, -1.0],
                       np.float32)
    pick = lambda shape: special[rng.integers(0, len(special), size=shape)]
    m = rng.random(src.shape) < 0.4
    src[m] = pick(int(m.sum()))
    pm = rng.random(pals.shape) < 0.3
    pals[pm] = pick(int(pm.sum()))
    w = blend[:, 4:].view(np.float32).reshape(n, nw).copy()
    wm = rng.random(w.shape) < 0.3
    w[wm] = pick(int(wm.sum()))
    blend[:, 4:] = w.view(np.uint8).reshape(n, 4 * nw)
    dsrc, ddst = ctx.to_device(src), ctx.alloc(ninst * n * 24)
    dbl, dpal, dmap = ctx.to_device(blend), ctx.to_device(pals), ctx.to_device(bone_map)
    j = gpu_job(ctx, dsrc.ptr, ddst.ptr, dbl, dpal, dmap, n, nw, nb, norm)
    ctx.set_config_option("skin_mode", mode)
    try:
        check(lib().ogrecu_skin_instanced_dev(ctx.handle, C.byref(j), ninst, (nb + 2) * 16, n * 24, n * 24))
        ctx.sync()
    finally:
        ctx.set_config_option("skin_mode", 0)
    got = ddst.download(np.float32).reshape(ninst, n, 6)
    for i in range(ninst):
        ref = oracle_skin(src[i], blend, np.ascontiguousarray(pals[i][bone_map]), nw, n, norm)
        ok = ~np.isnan(ref)                      # 0/0 normals: NaN on both sides, payload is not part of the contract
        assert np.array_equal(np.isnan(got[i]), np.isnan(ref))
        assert np.array_equal(bits(got[i])[ok], bits(ref)[ok]), (nw, mode, i)
    for b in (dsrc, ddst, dbl, dpal, dmap):
        b.free()


def test_multi_job_frame_equals_single_calls(gpu_ctx):
    """ogrecu_anim_instanced_multi_dev (all sub-meshes of a frame in one call, launches overlapped over side streams and
    joined back into the context's stream) == one ogrecu_anim_instanced_dev per sub-mesh, bit for bit, incl. the optional
    morphed intermediate; work queued on the context's stream before and after the call stays ordered."""
    from ogre_b200 import AnimJob, check, lib
    ctx = gpu_ctx
    ninst, nb = 70, 40
    shapes = [(1888, 4), (583, 3), (302, 1), (77, 2), (1533, 4)]
    rng = np.random.default_rng(99)
    t = rng.uniform(0, 1, size=ninst).astype(np.float32)
    dt = ctx.to_device(t)
    pals = np.stack([rigid_matrices(rng, nb) for _ in range(ninst)])
    dpal = ctx.to_device(pals)
    keep, jobs_single, outs = [], [], []
    arr = (AnimJob * len(shapes))()
    for k, (n, nw) in enumerate(shapes):
        keys, blend, _, bone_map = make_inputs(rng, n, nw, nb - 2, 2, True, True)
        dk1, dk2, dbl, dmap = ctx.to_device(np.ascontiguousarray(keys[0])), ctx.to_device(np.ascontiguousarray(keys[1])), ctx.to_device(blend), ctx.to_device(bone_map)
        da, db = ctx.alloc(ninst * n * 24), ctx.alloc(ninst * n * 24)
        mo_a = ctx.alloc(ninst * n * 24) if k == 1 else None
        mo_b = ctx.alloc(ninst * n * 24) if k == 1 else None
        ja = gpu_job(ctx, dk1.ptr, da.ptr, dbl, dpal, dmap, n, nw, nb - 2, True)
        jb = gpu_job(ctx, dk1.ptr, db.ptr, dbl, dpal, dmap, n, nw, nb - 2, True)
        arr[k].skin = ja
        arr[k].key2 = dk2.ptr
        arr[k].morph_out, arr[k].morph_instance_stride_bytes = (mo_a.ptr if mo_a else None), n * 24
        arr[k].dst_instance_stride_bytes = n * 24
        jobs_single.append((jb, dk2, mo_b, n))
        outs.append((da, db, mo_a, mo_b, n))
        keep += [dk1, dk2, dbl, dmap]
    check(lib().ogrecu_anim_instanced_multi_dev(ctx.handle, arr, len(shapes), dt.ptr, ninst, nb * 16))
    for jb, dk2, mo_b, n in jobs_single:
        check(lib().ogrecu_anim_instanced_dev(ctx.handle, C.byref(jb), dk2.ptr, dt.ptr, mo_b.ptr if mo_b else None, n * 24, ninst, nb * 16, n * 24))
    ctx.sync()
    for da, db, mo_a, mo_b, n in outs:
        assert np.array_equal(da.download(np.uint32), db.download(np.uint32))
        if mo_a:
            assert np.array_equal(mo_a.download(np.uint32), mo_b.download(np.uint32))
    # ordering against the context's stream: overwrite the destinations right after the call, then read
    da0 = outs[0][0]
    check(lib().ogrecu_anim_instanced_multi_dev(ctx.handle, arr, len(shapes), dt.ptr, ninst, nb * 16))
    marker = np.full(16, 0x7f, np.uint8)
    check(lib().ogrecu_memcpy_h2d(ctx.handle, da0.ptr, marker.ctypes.data, 16))
    ctx.sync()
    assert (da0.download(np.uint8, 16) == 0x7f).all(), "a side-stream launch finished after work queued behind the call"


STREAM_CASES = [
    # n,       nw, nb, norm   (one instance; n >= 2048 and aligned bases take skin_stream_kernel in skin_mode 0)
    (2048, 4, 64, True),          # exactly eight tiles, no tail
    (2049, 4, 64, True),          # one-vertex tail on the tile kernel
    (4099, 3, 55, True),
    (5000, 2, 18, True),
    (2303, 1, 7, True),
    (70000, 4, 256, True),        # maximum palette (> 48 KB of shared memory)
    (9001, 4, 30, False),         # positions only
    (3000, 2, 9, False),
    (256 * 2500 + 77, 4, 64, True),   # more tiles than resident CTAs: every CTA walks several tiles, the 3-stage ring wraps
    (256 * 1900, 1, 64, False),
]


@pytest.mark.parametrize("case", STREAM_CASES, ids=lambda c: "n%d_w%d_b%d_%s" % (c[0], c[1], c[2], "pn" if c[3] else "p"))
def test_single_stream_kernel(gpu_ctx, case):
    """ogre_b200/csrc/skin_stream.cu: one softwareVertexSkinning call over a long stream (persistent CTAs, resident palette,
    TMA pipeline over tiles) vs the oracle, bit patterns; and the same call forced onto the tile kernel (skin_mode 3) and the
    one-thread-per-vertex kernel (skin_mode 2) gives the same bytes, guard bytes around the destination untouched."""
    from ogre_b200 import check, lib
    n, nw, nb, norm = case
    ctx = gpu_ctx
    vs = 24 if norm else 12
    rng = np.random.default_rng(n + 7 * nw)
    src, blend, pals, bone_map = make_inputs(rng, n, nw, nb, 1, norm, True)
    dsrc = ctx.to_device(src.reshape(-1).view(np.uint8))
    dbl, dpal, dmap = ctx.to_device(blend), ctx.to_device(pals), ctx.to_device(bone_map)
    ref = oracle_skin(src[0], blend, np.ascontiguousarray(pals[0][bone_map]), nw, n, norm)
    results = {}
    for mode in (0, 3, 2):
        ddst = ctx.alloc(n * vs + 128)
        ddst.upload(np.full(n * vs + 128, 0xCD, np.uint8))
        j = gpu_job(ctx, dsrc.ptr, ddst.ptr + 64, dbl, dpal, dmap, n, nw, nb, norm)
        ctx.set_config_option("skin_mode", mode)
        try:
            check(lib().ogrecu_skin_instanced_dev(ctx.handle, C.byref(j), 1, 0, 0, 0))
            ctx.sync()
        finally:
            ctx.set_config_option("skin_mode", 0)
        raw = ddst.download(np.uint8)
        assert (raw[:64] == 0xCD).all() and (raw[64 + n * vs:] == 0xCD).all(), (case, mode, "wrote outside the destination")
        results[mode] = raw[64:64 + n * vs].copy()
        ddst.free()
    got = results[0].view(np.float32).reshape(n, vs // 4)
    assert np.array_equal(bits(got), bits(ref)), case
    assert np.array_equal(results[0], results[3]) and np.array_equal(results[0], results[2]), case
    for b in (dsrc, dbl, dpal, dmap):
        b.free()


def test_single_stream_kernel_special_values(gpu_ctx):
    # signed zeros, denormals, huge and non-finite values through the stream kernel = the tile kernel, byte for byte
    from ogre_b200 import check, lib
    ctx = gpu_ctx
    n, nw, nb = 4096, 4, 16
    rng = np.random.default_rng(3)
    src, blend, pals, bone_map = make_inputs(rng, n, nw, nb, 1, True, True)
    special = np.array([0.0, -0.0, 1e-42, -1e-42, 3e38, -3e38, np.inf, -np.inf, np.nan, 1.0], np.float32)
    src.reshape(-1)[::5] = special[rng.integers(0, len(special), src.reshape(-1)[::5].size)]
    pals.reshape(-1)[::7] = special[rng.integers(0, len(special), pals.reshape(-1)[::7].size)]
    w = blend[:, 4:].view(np.float32)
    w[::3, 0] = 0.0; w[1::9, 1] = -0.0
    dsrc, dbl, dpal, dmap = ctx.to_device(src.reshape(-1).view(np.uint8)), ctx.to_device(blend), ctx.to_device(pals), ctx.to_device(bone_map)
    out = {}
    for mode in (0, 3):
        ddst = ctx.alloc(n * 24)
        j = gpu_job(ctx, dsrc.ptr, ddst.ptr, dbl, dpal, dmap, n, nw, nb, True)
        ctx.set_config_option("skin_mode", mode)
        try:
            check(lib().ogrecu_skin_instanced_dev(ctx.handle, C.byref(j), 1, 0, 0, 0))
            ctx.sync()
        finally:
            ctx.set_config_option("skin_mode", 0)
        out[mode] = ddst.download(np.uint8)[:n * 24].view(np.uint32).copy()
        ddst.free()
    # NaN payloads may differ between code paths only if the arithmetic differs; it is the same code, so compare raw bits
    assert np.array_equal(out[0], out[3])
    for b in (dsrc, dbl, dpal, dmap):
        b.free()


@pytest.mark.parametrize("nw_used", [1, 2, 3])
def test_single_stream_kernel_wider_blend_stride(gpu_ctx, nw_used):
    # a blend stream laid out for four weights (stride 20) skinned with fewer weights per vertex: the stream kernel takes
    # any stride that is a multiple of 4 bytes
    from ogre_b200 import check, lib
    ctx = gpu_ctx
    n, nb = 256 * 700 + 31, 40
    rng = np.random.default_rng(50 + nw_used)
    src, blend, pals, bone_map = make_inputs(rng, n, 4, nb, 1, True, True)
    dsrc, dbl, dpal, dmap = ctx.to_device(src.reshape(-1).view(np.uint8)), ctx.to_device(blend), ctx.to_device(pals), ctx.to_device(bone_map)
    ddst = ctx.alloc(n * 24)
    j = gpu_job(ctx, dsrc.ptr, ddst.ptr, dbl, dpal, dmap, n, 4, nb, True)
    j.num_weights = nw_used                      # strides stay 20
    check(lib().ogrecu_skin_instanced_dev(ctx.handle, C.byref(j), 1, 0, 0, 0))
    ctx.sync()
    got = ddst.download(np.uint8)[:n * 24].view(np.float32).reshape(n, 6)
    ref = np.zeros((n, 6), np.float32)
    s = np.ascontiguousarray(src[0])
    oracle.skin("sse", s.ctypes.data, ref.ctypes.data, s.ctypes.data + 12, ref.ctypes.data + 12, blend.ctypes.data + 4, blend.ctypes.data,
                np.ascontiguousarray(pals[0][bone_map]), (24, 24, 24, 24, 20, 20), nw_used, n)
    assert np.array_equal(bits(got), bits(ref))
    for b in (dsrc, dbl, dpal, dmap, ddst):
        b.free()
