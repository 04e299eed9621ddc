"""GPU parity: CUDA OptimisedUtil (through the C ABI) vs the CPU oracle, bit-exact.

The oracle restates the SSE operation order (skinning positions, concat, facing) and the
General one where SSE is approximate (normals, extrusion) — oracle/ogre_oracle.h.  Because
the kernels replay the same binary32 operation sequence without FMA, the comparison is on
BIT PATTERNS (sign of zero canonicalised), which is stricter than north_star's 1e-5.
"""
import numpy as np
import pytest

import oracle
from datagen import SkinCase, bits, edge_triangles, rigid_matrices

pytestmark = pytest.mark.gpu


def run_skin_oracle(case, kind="sse"):
    dst, dstn = case.new_dst()
    sp, dp, sn, dn, bw, bi = case.pointers(dst, dstn)
    oracle.skin(kind, sp, dp, sn, dn, bw, bi, case.mats, case.strides(), case.nw, case.n)
    return dst, dstn


def run_skin_gpu(ctx, case):
    import ogre_b200
    ou = ogre_b200.OptimisedUtil(ctx)
    dst, dstn = case.new_dst()
    sp, dp, sn, dn, bw, bi = case.pointers(dst, dstn)
    ou.softwareVertexSkinning(sp, dp, sn, dn, bw, bi, list(case.mats), *case.strides(), case.nw, case.n)
    return dst, dstn


@pytest.mark.parametrize("layout", ["shared", "separate", "posonly", "fat"])
@pytest.mark.parametrize("nw", [1, 2, 3, 4])
def test_skinning_bit_exact_vs_oracle(gpu_ctx, layout, nw):
    for n, head in [(1, 0), (3, 8), (17, 4), (1000, 0), (4099, 12)]:
        case = SkinCase(seed=100 * nw + n, n=n, nw=nw, layout=layout, head=head)
        ref, refn = run_skin_oracle(case)
        got, gotn = run_skin_gpu(gpu_ctx, case)
        # whole destination buffers, including bytes that must stay untouched (gaps, guard)
        assert np.array_equal(got, ref), (layout, nw, n, head)
        if refn is not None:
            assert np.array_equal(gotn, refn)


def test_skinning_within_1e5_of_general(gpu_ctx):
    # north_star tolerance: positions within 1e-5 relative of the reference (General differs
    # from SSE by up to 1e-5 itself)
    for nw in (1, 2, 3, 4):
        case = SkinCase(seed=7 + nw, n=20000, nw=nw, layout="shared")
        ref, _ = run_skin_oracle(case, "general")
        got, _ = run_skin_gpu(gpu_ctx, case)
        pr, nr = case.dst_views(ref, None)
        pg, ng = case.dst_views(got, None)
        scale = np.maximum(np.abs(pr).max(), 1.0)
        assert np.abs(pg - pr).max() / scale <= 1e-5
        if nw > 1:   # General skips the multiply when nw==1? no: it scales by the weight; SSE ignores it
            assert np.abs(ng - nr).max() <= 1e-5


def test_skinning_separate_blend_streams(gpu_ctx):
    case = SkinCase(seed=5, n=777, nw=3, layout="shared", blend_interleaved=False)
    ref, _ = run_skin_oracle(case)
    got, _ = run_skin_gpu(gpu_ctx, case)
    assert np.array_equal(got, ref)


def test_skin_dev_batched_and_instanced(gpu_ctx):
    import ctypes as C
    import ogre_b200
    from ogre_b200 import SkinJob, check, lib
    ctx = gpu_ctx
    cases = [SkinCase(seed=40 + k, n=n, nw=nw, layout="shared", nbones=nb)
             for k, (n, nw, nb) in enumerate([(288, 1, 2), (1888, 4, 55), (583, 3, 6), (1, 2, 7), (1533, 4, 18)])]
    jobs = (SkinJob * len(cases))()
    keep = []
    for k, c in enumerate(cases):
        dsrc = ctx.to_device(c.src); dbl = ctx.to_device(c.blend); dpal = ctx.to_device(c.mats)
        ddst = ctx.alloc(c.n * 24)
        keep.append((dsrc, dbl, dpal, ddst))
        j = jobs[k]
        j.src_pos, j.src_norm = dsrc.ptr, dsrc.ptr + 12
        j.dst_pos, j.dst_norm = ddst.ptr, ddst.ptr + 12
        j.blend_weight, j.blend_index = dbl.ptr + 4, dbl.ptr
        j.palette, j.bone_map = dpal.ptr, None
        j.src_pos_stride = j.dst_pos_stride = j.src_norm_stride = j.dst_norm_stride = 24
        j.blend_weight_stride = j.blend_index_stride = 4 + 4 * c.nw
        j.num_weights, j.num_matrices, j.num_vertices = c.nw, len(c.mats), c.n
    check(lib().ogrecu_skin_dev(ctx.handle, jobs, len(cases)))
    ctx.sync()
    for c, (_, _, _, ddst) in zip(cases, keep):
        ref, _ = run_skin_oracle(c)
        got = ddst.download(np.uint8)
        assert np.array_equal(got, ref[:c.n * 24])

    # instanced: 5 instances of case 1 with different palettes, bone_map permutation
    c = cases[1]
    rng = np.random.default_rng(9)
    ninst, nb = 5, len(c.mats)
    pals = np.stack([rigid_matrices(rng, nb + 3) for _ in range(ninst)])          # [ninst, nb+3, 16]
    bone_map = rng.permutation(nb + 3)[:nb].astype(np.uint16)
    dsrc, dbl = keep[1][0], keep[1][1]
    dpal = ctx.to_device(pals); dmap = ctx.to_device(bone_map)
    ddst = ctx.alloc(ninst * c.n * 24)
    j = SkinJob()
    j.src_pos, j.src_norm = dsrc.ptr, dsrc.ptr + 12
    j.dst_pos, j.dst_norm = ddst.ptr, ddst.ptr + 12
    j.blend_weight, j.blend_index = dbl.ptr + 4, dbl.ptr
    j.palette, j.bone_map = dpal.ptr, dmap.ptr
    j.src_pos_stride = j.dst_pos_stride = j.src_norm_stride = j.dst_norm_stride = 24
    j.blend_weight_stride = j.blend_index_stride = 4 + 4 * c.nw
    j.num_weights, j.num_matrices, j.num_vertices = c.nw, nb, c.n
    check(lib().ogrecu_skin_instanced_dev(ctx.handle, C.byref(j), ninst, (nb + 3) * 16, 0, c.n * 24))
    ctx.sync()
    got = ddst.download(np.uint8).reshape(ninst, c.n * 24)
    for i in range(ninst):
        c.mats = np.ascontiguousarray(pals[i][bone_map])
        ref, _ = run_skin_oracle(c)
        assert np.array_equal(got[i], ref[:c.n * 24]), i


@pytest.mark.parametrize("normals", [False, True])
def test_morph_bit_exact(gpu_ctx, normals):
    import ogre_b200
    from ogre_b200 import addr
    ou = ogre_b200.OptimisedUtil(gpu_ctx)
    rng = np.random.default_rng(3)
    vs = 24 if normals else 12
    for n, stride in [(1, vs), (2, vs), (5, vs), (1001, vs), (65537, vs), (333, vs + 8)]:
        a = rng.uniform(-10, 10, size=n * stride // 4 + 4).astype(np.float32)
        b = rng.uniform(-10, 10, size=n * stride // 4 + 4).astype(np.float32)
        ref = np.full(n * stride // 4 + 4, 7.0, dtype=np.float32); got = ref.copy()
        oracle.lib().oracle_morph(0.37, addr(a), addr(b), addr(ref), stride, stride, stride, n, int(normals))
        ou.softwareVertexMorph(0.37, addr(a), addr(b), addr(got), stride, stride, stride, n, normals)
        assert np.array_equal(bits(got), bits(ref)), (n, stride)


def test_concat_affine(gpu_ctx):
    import ogre_b200
    from ogre_b200 import addr
    ou = ogre_b200.OptimisedUtil(gpu_ctx)
    rng = np.random.default_rng(4)
    base = rigid_matrices(rng, 1)[0]
    for n in (1, 69, 256, 1000):
        src = rigid_matrices(rng, n)
        ref = np.zeros_like(src); got = np.ones_like(src)
        oracle.lib().oracle_concat_affine_sse(addr(base), addr(src), addr(ref), n)
        ou.concatenateAffineMatrices(base, addr(src), addr(got), n)
        assert np.array_equal(bits(got), bits(ref))
        gen = np.zeros_like(src)
        oracle.lib().oracle_concat_affine_general(addr(base), addr(src), addr(gen), n)
        assert np.abs(got - gen).max() <= 1e-5 * max(1.0, np.abs(gen).max())


def test_concat_affine_instanced_equals_the_oracle_per_instance(gpu_ctx):
    """ogrecu_concat_affine_instanced_dev (Entity::cacheBoneMatrices for many entities in one launch) == the SSE restatement
    called once per instance with that instance's base matrix; in place, padded instance stride."""
    import ctypes as C
    from ogre_b200 import addr, check, lib
    ctx = gpu_ctx
    rng = np.random.default_rng(14)
    for ninst, n, pad in ((1, 69, 0), (37, 69, 2), (300, 5, 0)):
        bases = rigid_matrices(rng, ninst)
        bases[:, 3] += rng.uniform(-300, 300, size=ninst).astype(np.float32)
        src = np.stack([np.concatenate([rigid_matrices(rng, n), np.full((pad, 16), 7.0, np.float32)]) for _ in range(ninst)])
        ref = src.copy()
        for i in range(ninst):
            out = np.zeros((n, 16), np.float32)
            oracle.lib().oracle_concat_affine_sse(addr(np.ascontiguousarray(bases[i])), addr(np.ascontiguousarray(src[i, :n])), addr(out), n)
            ref[i, :n] = out
        db, dm = ctx.to_device(bases), ctx.to_device(src)
        check(lib().ogrecu_concat_affine_instanced_dev(ctx.handle, db.ptr, dm.ptr, dm.ptr, n, ninst, (n + pad) * 16))
        ctx.sync()
        got = dm.download(np.float32).reshape(ninst, n + pad, 16)
        assert np.array_equal(bits(got), bits(ref))
        db.free(); dm.free()


def test_face_normals_light_facing_extrude(gpu_ctx):
    import ogre_b200
    from ogre_b200 import addr
    ou = ogre_b200.OptimisedUtil(gpu_ctx)
    rng = np.random.default_rng(5)
    for nv, nt in [(3, 1), (50, 7), (5000, 20001), (100000, 262147)]:
        pos = rng.uniform(-10, 10, size=(nv, 3)).astype(np.float32)
        tris = edge_triangles(rng, nt, nv)
        ref = np.zeros((nt, 4), np.float32); got = np.ones((nt, 4), np.float32)
        oracle.lib().oracle_face_normals(addr(pos), addr(tris), addr(ref), nt)
        ou.calculateFaceNormals(addr(pos), addr(tris), addr(got), nt)
        assert np.array_equal(bits(got), bits(ref))
        for light in ([30.0, 2.0, -11.0, 1.0], [0.3, -0.9, 0.1, 0.0]):
            l = np.array(light, np.float32)
            fr = np.zeros(nt + 8, np.int8); fg = np.full(nt + 8, 9, np.int8); fr[nt:] = 9
            oracle.lib().oracle_light_facing_sse(addr(l), addr(ref), addr(fr), nt)
            ou.calculateLightFacing(l, addr(got), addr(fg), nt)
            assert np.array_equal(fg, fr)
            er = np.zeros((nv, 3), np.float32); eg = np.ones((nv, 3), np.float32)
            oracle.lib().oracle_extrude(addr(l), 250.0, addr(pos), addr(er), nv)
            ou.extrudeVertices(l, 250.0, addr(pos), addr(eg), nv)
            assert np.array_equal(bits(eg), bits(er))


def test_extrude_same_buffer_tail(gpu_ctx):
    # OGRE extrudes into the second half of the same buffer: dst = src + numVertices*3
    # (OgreShadowCaster.cpp:528-537), so dst is generally not 16-byte aligned
    import ogre_b200
    from ogre_b200 import addr
    ou = ogre_b200.OptimisedUtil(gpu_ctx)
    rng = np.random.default_rng(6)
    for nv in (1, 2, 3, 5, 1023, 4097):
        buf = rng.uniform(-5, 5, size=nv * 6).astype(np.float32)
        ref = buf.copy(); got = buf.copy()
        l = np.array([1.0, 20.0, 3.0, 1.0], np.float32)
        oracle.lib().oracle_extrude(addr(l), 100.0, addr(ref), addr(ref, nv * 12), nv)
        ou.extrudeVertices(l, 100.0, addr(got), addr(got, nv * 12), nv)
        assert np.array_equal(bits(got), bits(ref))


def test_extrude_lengths_outside_the_fast_range(gpu_ctx):
    """The extrusion takes its exact branch-free sqrt / reciprocal sequences only when the four squared lengths of a thread are
    in their range; a vertex on the light (length 0: Vector3::normalise leaves it alone), denormal, huge (squared length
    overflows) distances take the library path.  Mixed inside and across the groups of four, bit for bit against
    the oracle, single-light and multi-light entry points."""
    import ogre_b200
    from ogre_b200 import addr, capi, check, lib
    ctx = gpu_ctx
    ou = ogre_b200.OptimisedUtil(ctx)
    rng = np.random.default_rng(11)
    light = np.array([1.5, -2.0, 0.25, 1.0], np.float32)
    nv = 4099
    pos = rng.uniform(-10, 10, size=(nv, 3)).astype(np.float32)
    special = [light[:3],                                   # on the light: length 0
               light[:3] + np.float32(1e-30),                # denormal squared length
               light[:3] + np.array([1e-21, 0, 0], np.float32),
               np.array([3e19, 1e19, -2e19], np.float32),    # squared length overflows to inf
               np.array([1e25, 0, 0], np.float32),
               light[:3] + np.array([2.0 ** -20, 0, 0], np.float32),   # the ends of the fast range
               light[:3] + np.array([2.0 ** 19.9, 0, 0], np.float32)]
    where = rng.choice(nv, size=6 * len(special), replace=False)
    for i, w in enumerate(where):
        pos[w] = special[i % len(special)]
    pos[0:4] = special[0:4]                                  # a whole group of four off the fast path
    ref = np.zeros((nv, 3), np.float32); got = np.ones((nv, 3), np.float32)
    with np.errstate(all="ignore"):
        oracle.lib().oracle_extrude(addr(light), 250.0, addr(pos), addr(ref), nv)
    ou.extrudeVertices(light, 250.0, addr(pos), addr(got), nv)
    assert np.array_equal(bits(got), bits(ref))
    lights = np.stack([light, np.array([-4.0, 9.0, 2.0, 1.0], np.float32), np.array([0.3, -0.9, 0.1, 0.0], np.float32)])
    dpos, dout = ctx.to_device(pos), ctx.alloc(3 * 3 * nv * 4)
    check(lib().ogrecu_extrude_multi_dev(ctx.handle, lights.ctypes.data_as(capi.c_float_p), 3, 250.0, dpos.ptr, dout.ptr, 3 * nv, nv))
    gm = dout.download(np.float32).reshape(3, nv, 3)
    for k in range(3):
        r = np.zeros((nv, 3), np.float32)
        l = np.ascontiguousarray(lights[k])
        oracle.lib().oracle_extrude(addr(l), 250.0, addr(pos), addr(r), nv)
        assert np.array_equal(bits(gm[k]), bits(r)), k
    dpos.free(); dout.free()


def test_extrude_multi_light_equals_per_light(gpu_ctx):
    """ogrecu_extrude_multi_dev: 1..8 lights in one launch (positions read once), per light bit-identical to the oracle;
    point and directional lights mixed, aligned and unaligned strides, ragged vertex counts."""
    import ctypes as C
    from ogre_b200 import addr, capi, check, lib
    ctx = gpu_ctx
    rng = np.random.default_rng(8)
    for nv, pad in [(1, 0), (5, 1), (1023, 0), (4097, 3), (100000, 0)]:
        pos = rng.uniform(-10, 10, size=(nv, 3)).astype(np.float32)
        for nl in (1, 3, 8):
            lights = rng.uniform(-30, 30, size=(nl, 4)).astype(np.float32)
            lights[:, 3] = (np.arange(nl) % 3 != 1)                 # every third light directional
            stride = 3 * nv + pad
            dpos, dout = ctx.to_device(pos), ctx.alloc(nl * stride * 4)
            check(lib().ogrecu_extrude_multi_dev(ctx.handle, lights.ctypes.data_as(capi.c_float_p), nl, 250.0, dpos.ptr, dout.ptr, stride, nv))
            got = dout.download(np.float32).reshape(nl, stride)[:, :3 * nv].reshape(nl, nv, 3)
            for k in range(nl):
                ref = np.zeros((nv, 3), np.float32)
                l = np.ascontiguousarray(lights[k])
                oracle.lib().oracle_extrude(addr(l), 250.0, addr(pos), addr(ref), nv)
                assert np.array_equal(bits(got[k]), bits(ref)), (nv, nl, k)
            dpos.free(); dout.free()
