"""CPU, build container only: the oracle restatement vs the LIVE unmodified reference in
oracle/_ref (real TinyRenderSystem::_render, real OptimisedUtil).  Skipped where oracle/_ref
has not been built (it cannot be rebuilt without /root/reference); the committed golden
fixtures (test_oracle_golden.py) carry the same pinning everywhere else."""
import numpy as np
import pytest

import oracle
from oracle import ref
from datagen import SkinCase, bits, clip_case, degenerate_soup, make_scene
from rasterlib import assert_images_equal, render_oracle

pytestmark = [pytest.mark.ref, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]


@pytest.mark.parametrize("w,h", [(257, 193), (64, 64), (33, 17)])
def test_oracle_raster_equals_real_tiny(w, h):
    for seed in range(4):
        cases = make_scene(seed * 7 + w, w, h, ndraws=6, ntri=150)
        oc, oz, _ = render_oracle(cases, w, h)
        rc, rz = ref.tiny_render(cases, w, h)
        assert_images_equal(oc, oz, rc, rz, "seed %d" % seed)


def test_oracle_raster_corner_cases_equal_real_tiny():
    w, h = 128, 96
    clip = degenerate_soup(np.random.default_rng(2), w, h)
    for kw in (dict(lighting=False), dict(lighting=True, blend_add=True), dict(depth_write=False), dict(depth_test=False)):
        cases = clip_case(clip, w, h, **kw)
        oc, oz, _ = render_oracle(cases, w, h)
        rc, rz = ref.tiny_render(cases, w, h)
        assert_images_equal(oc, oz, rc, rz, repr(kw))


def test_oracle_raster_tiny_triangles_and_no_clear():
    w, h = 320, 240
    cases = make_scene(5, w, h, ndraws=3, ntri=20000, size=0.02)
    oc, oz, _ = render_oracle(cases, w, h)
    rc, rz = ref.tiny_render(cases, w, h)
    assert_images_equal(oc, oz, rc, rz, "tiny triangles")
    rng = np.random.default_rng(9)
    init = (rng.integers(0, 256, size=(80, 100, 3), dtype=np.uint8), rng.uniform(0.3, 1.0, size=(80, 100)).astype(np.float32))
    cases = make_scene(3, 100, 80, ndraws=4, ntri=60)
    oc, oz, _ = render_oracle(cases, 100, 80, clear=None, init=init)
    rc, rz = ref.tiny_render(cases, 100, 80, clear=None, init=init)
    assert_images_equal(oc, oz, rc, rz, "no clear")


def test_oracle_skinning_equals_real_sse_large():
    l = ref.lib()
    for nw in (1, 2, 3, 4):
        c = SkinCase(seed=nw, n=50001, nw=nw, layout="shared")
        dst, _ = c.new_dst(); rdst, _ = c.new_dst()
        sp, dp, sn, dn, bw, bi = c.pointers(dst, None)
        oracle.skin("sse", sp, dp, sn, dn, bw, bi, c.mats, c.strides(), nw, c.n)
        sp, dp, sn, dn, bw, bi = c.pointers(rdst, None)
        keep, pal = oracle.palette_ptrs(c.mats)
        l.ref_skin(ref.SSE, sp, dp, sn, dn, bw, bi, pal, *c.strides(), nw, c.n)
        po, no = c.dst_views(dst, None); pr, nr = c.dst_views(rdst, None)
        assert np.array_equal(bits(po), bits(pr))
        assert np.abs(no - nr).max() <= 4e-4


@pytest.mark.parametrize("mesh", ["Sinbad.mesh", "jaiqua.mesh"])
def test_oracle_bone_palettes_equal_real_skeleton_instance(mesh):
    """300 random poses (1-3 states, random order, times and weights, both blend modes) through the real
    SkeletonInstance::setAnimationState + _getBoneMatrices vs the restatement: bit-identical."""
    d = ref.skeleton_def(mesh)
    sk = oracle.Skeleton(d)
    assert np.array_equal(bits(sk.bind_inverse), bits(d["bind_inverse"]))
    rng = np.random.default_rng(11)
    for trial in range(300):
        ns = int(rng.integers(1, 4))
        an = rng.choice(len(d["anims"]), ns, replace=False)
        tm = [rng.uniform(0, 1.2 * d["anims"][a]["length"]) for a in an]       # beyond the end: clamped by AnimationState
        tm = [min(t, d["anims"][a]["length"]) for t, a in zip(tm, an)]
        w = rng.uniform(0.0, 1.0, ns) if trial % 3 else np.ones(ns)
        mode = trial % 2
        ref.lib().ref_skel_set_blend_mode(mode)
        r = ref.skeleton_pose_states(d["num_bones"], an, tm, w)
        o = sk.palette(an, tm, w, bind_inverse=d["bind_inverse"], blend_mode=mode)
        assert np.array_equal(bits(r), bits(o)), (mesh, trial)
    ref.lib().ref_skel_set_blend_mode(d["blend_mode"])


@pytest.mark.parametrize("mesh", ["jaiqua.mesh", "ogrehead.mesh", "Sinbad.mesh"])
def test_oracle_shadow_volume_equals_real_generate_shadow_volume(mesh):
    """60 random (light, bounds, flags, offset) cases per mesh through the real ShadowCaster::generateShadowVolume on the
    mesh's own EdgeData vs the restatement: identical index buffers, ranges and buffer positions."""
    ed = ref.shadow_open(mesh)
    rng = np.random.default_rng(2)
    for trial in range(60):
        if trial % 3 == 0:
            v = rng.normal(size=3); light = np.array([*(v / np.linalg.norm(v)), 0], np.float32)
        else:
            light = np.array([*(rng.uniform(-1, 1, 3) * rng.choice([5, 40, 300])), 1], np.float32)
        half = rng.uniform(3, 40, 3); c = rng.uniform(-5, 5, 3)
        bounds = np.array([*(c - half), *(c + half)], np.float32)
        flags = int(rng.integers(0, 8)); sep = bool(rng.integers(0, 2)); cap = 200000; used = int(rng.choice([0, 1000, 199000]))
        ri, rr, ru, fac = ref.shadow_volume(ed, light, bounds, flags, sep, cap, used)
        oi, orr, ou, n, mc = oracle.shadow_volume(ed, fac, light, bounds, flags, sep, cap, used)
        assert np.array_equal(ri, oi) and ru == ou, (mesh, trial)
        cols = 4 if (sep and flags & 1) else 2
        assert np.array_equal(rr[:, :cols], orr[:, :cols]), (mesh, trial)


def test_oracle_pose_blend_equals_real_mesh_pose_blend():
    """Mesh::softwareVertexPoseBlend (public static) on a real HardwareVertexBuffer vs the restatement: weights incl. 0
    (early return), 1, negative; positions only and positions + normals; signed zeros and denormals in the buffer."""
    from datagen import pose_case
    l = ref.lib()
    for seed, nv, frac, normals in [(1, 4000, 0.3, True), (2, 17, 1.0, True), (3, 1000, 0.01, False), (4, 1, 1.0, True)]:
        buf, oi, off, ni, nrm = pose_case(seed, nv, frac, normals)
        fpv = buf.shape[1]
        for w in (0.0, 1.0, 0.37, -0.5, 1e-30):
            r = buf.copy(); o = buf.copy()
            assert l.ref_pose_blend(w, len(oi), oi.ctypes.data, off.ctypes.data, len(ni), ni.ctypes.data, nrm.ctypes.data, r.ctypes.data, nv, fpv,
                                    int(normals)) == 0
            oracle.lib().oracle_pose_blend(w, oi.ctypes.data, off.ctypes.data, len(oi), o.ctypes.data, fpv)
            if normals:
                oracle.lib().oracle_pose_blend(w, ni.ctypes.data, nrm.ctypes.data, len(ni), o.ctypes.data + 12, fpv)
            assert np.array_equal(bits(r), bits(o)), (seed, w)


@pytest.mark.parametrize("seed,nv,nposes,normals", [(1, 500, 3, True), (2, 500, 3, False), (3, 37, 1, True), (4, 2000, 5, True)])
def test_oracle_pose_frame_equals_real_entity(seed, nv, nposes, normals):
    """The whole pose-animation path of a real Entity on a hand-built mesh (initialisePoseVertexData -> pose blends in
    key-frame order -> finalisePoseNormals, through Entity::_updateAnimation) vs the restatements chained: identical."""
    from datagen import pose_entity_case, pose_entity_oracle
    base, offs, idx, off, nrm, infl = pose_entity_case(seed, nv, nposes)
    out = np.zeros((nv, 6), np.float32)
    rc = ref.lib().ref_pose_entity(nv, base.ctypes.data, nposes, offs.ctypes.data, idx.ctypes.data, off.ctypes.data, nrm.ctypes.data,
                                   infl.ctypes.data, int(normals), out.ctypes.data)
    assert rc == 0
    assert np.array_equal(bits(out), bits(pose_entity_oracle(oracle, base, offs, idx, off, nrm, infl, normals)))


def test_oracle_image_scale_equals_real_image_scale():
    """LinearResampler_Byte restated vs the real Image::scale(FILTER_BILINEAR): random sizes, 1-4 channel byte formats, sub-boxes,
    padded destination rows; and scale + format conversion (Image::scale's temp buffer + bulkPixelConversion tail) vs the
    restated scale followed by the shift packing ogrecu_rt_read_colour_scaled uses."""
    from datagen import PACKED32, pack32
    rng = np.random.default_rng(1)
    for c, fmt in ((3, "PF_BYTE_RGB"), (4, "PF_BYTE_RGBA"), (1, "PF_L8"), (2, "PF_BYTE_LA")):
        for _ in range(25):
            h, w = (int(v) for v in rng.integers(1, 90, 2)); dh, dw = (int(v) for v in rng.integers(1, 150, 2))
            src = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
            assert np.array_equal(ref.image_scale(src, dw, dh, fmt, dst_pad=3), oracle.scale_bilinear(src, dw, dh, dst_pad=3))
            if w > 4 and h > 4:
                box = (1, 2, w - 1, h - 1)
                assert np.array_equal(ref.image_scale(src, dw, dh, fmt, box=box), oracle.scale_bilinear(src, dw, dh, box=box))
    src = rng.integers(0, 256, (45, 77, 3), dtype=np.uint8)
    for name, layout in PACKED32.items():
        want = ref.image_scale(src, 120, 31, "PF_BYTE_RGB", name).view(np.uint32)[..., 0]
        assert np.array_equal(pack32(oracle.scale_bilinear(src, 120, 31), *layout), want), name


@pytest.mark.parametrize("im,rim", [(0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("mesh", ["Sinbad.mesh", "jaiqua.mesh"])
def test_oracle_spline_and_spherical_interpolation_equal_real_skeleton_instance(mesh, im, rim):
    """RIM_SPHERICAL (Quaternion::Slerp in key interpolation and in the weight application, OgreAnimationTrack.cpp:417-421,
    495-498) and IM_SPLINE (SimpleSpline / RotationalSpline over all keys of a track, :438-450,515-549): every animation of the
    real skeleton switched to the mode, 200 random multi-state poses, the restatement bit-identical (same libm in one
    process).  The device path refuses these modes (tests/test_skeleton_gpu.py) — libm's transcendentals have no
    bit-identical device counterpart."""
    d = ref.skeleton_def(mesh)
    try:
        for a in range(len(d["anims"])):
            assert ref.lib().ref_skel_set_interpolation(a, im, rim) == 0
        ds = ref.skeleton_def()
        assert all(a["rotation_interpolation"] == rim and a["interpolation"] == im for a in ds["anims"])
        sk = oracle.Skeleton(ds)
        lin = oracle.Skeleton(d)
        rng = np.random.default_rng(23)
        differs = 0
        for trial in range(200):
            ns = int(rng.integers(1, 4))
            an = rng.choice(len(ds["anims"]), ns, replace=False)
            tm = [rng.uniform(0, ds["anims"][a]["length"]) for a in an]
            w = rng.uniform(0.0, 1.0, ns) if trial % 3 else np.ones(ns)
            mode = trial % 2
            ref.lib().ref_skel_set_blend_mode(mode)
            r = ref.skeleton_pose_states(ds["num_bones"], an, tm, w)
            o = sk.palette(an, tm, w, bind_inverse=ds["bind_inverse"], blend_mode=mode)
            assert np.array_equal(bits(r), bits(o)), (mesh, trial)
            differs += int(not np.array_equal(bits(o), bits(lin.palette(an, tm, w, bind_inverse=d["bind_inverse"], blend_mode=mode))))
        assert differs > 100          # the mode really is a different interpolation
    finally:
        for a in range(len(d["anims"])):
            ref.lib().ref_skel_set_interpolation(a, d["anims"][a]["interpolation"], d["anims"][a]["rotation_interpolation"])
        ref.lib().ref_skel_set_blend_mode(d["blend_mode"])


def random_manual_bones(rng, d, k):
    """k manually controlled bones with local transforms near their binding pose"""
    handles = rng.choice(d["num_bones"], k, replace=False).astype(np.uint32)
    x = np.array(d["bind"], np.float32).reshape(-1, 10)[handles].copy()
    x[:, 0:3] += rng.uniform(-0.5, 0.5, (k, 3)).astype(np.float32)
    q = x[:, 3:7].astype(np.float64) + rng.uniform(-0.3, 0.3, (k, 4))
    x[:, 3:7] = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    x[:, 7:10] *= rng.uniform(0.8, 1.25, (k, 3)).astype(np.float32)
    slot = np.full(d["num_bones"], -1, np.int32)
    slot[handles] = np.arange(k, dtype=np.int32)
    return handles, x, slot


def held_by_the_bone(x):
    """what a Bone holds after setPosition / setOrientation / setScale: Node::setOrientation normalises the quaternion
    (OgreNode.cpp: mOrientation = q; mOrientation.normalise()), in float32 as Quaternion::normalise does: (1.0f / len) * q"""
    x = np.array(x, np.float32).copy()
    if len(x):
        q = x[:, 3:7]
        n2 = ((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2]) + q[:, 3] * q[:, 3]
        s = np.float32(1.0) / np.sqrt(n2, dtype=np.float32)
        x[:, 3:7] = s[:, None] * q
    return x


@pytest.mark.parametrize("mesh", ["Sinbad.mesh", "jaiqua.mesh"])
def test_oracle_blend_masks_and_manual_bones_equal_real_skeleton_instance(mesh):
    """AnimationState blend masks (Animation::apply with a mask, OgreAnimation.cpp:344-357: mask[bone] * weight per track) and
    manually controlled bones (Skeleton::reset leaves them alone, OgreSkeleton.cpp:233-240; their tracks apply on top of what they
    hold): 150 random multi-state poses against the real SkeletonInstance, masks with zeros and ones among the values, 0-4 manual
    bones per pose — bit-identical."""
    d = ref.skeleton_def(mesh)
    sk = oracle.Skeleton(d)
    rng = np.random.default_rng(41)
    nb = d["num_bones"]
    plain_differs = 0
    for trial in range(150):
        ns = int(rng.integers(1, 4))
        an = rng.choice(len(d["anims"]), ns, replace=False)
        tm = [rng.uniform(0, d["anims"][a]["length"]) for a in an]
        w = rng.uniform(0.0, 1.0, ns) if trial % 3 else np.ones(ns)
        mode = trial % 2
        mask = None
        if trial % 4 != 3:
            mask = rng.uniform(0, 1, (ns, nb)).astype(np.float32)
            mask[rng.uniform(size=(ns, nb)) < 0.2] = 0.0
            mask[rng.uniform(size=(ns, nb)) < 0.2] = 1.0
        k = int(rng.integers(0, 5))
        handles, xf, slot = random_manual_bones(rng, d, k)
        ref.lib().ref_skel_set_blend_mode(mode)
        r = ref.skeleton_pose_states_ex(nb, an, tm, w, mask, handles, xf)
        o = sk.palette(an, tm, w, bind_inverse=d["bind_inverse"], blend_mode=mode, mask=mask, manual_slot=slot if k else None,
                       manual=held_by_the_bone(xf) if k else None)
        assert np.array_equal(bits(r), bits(o)), (mesh, trial)
        plain_differs += int(not np.array_equal(bits(o), bits(sk.palette(an, tm, w, bind_inverse=d["bind_inverse"], blend_mode=mode))))
        # the engine is left as it was: the next plain pose is the plain pose
        if trial % 25 == 0:
            assert np.array_equal(bits(ref.skeleton_pose_states(nb, an, tm, w)), bits(sk.palette(an, tm, w, bind_inverse=d["bind_inverse"], blend_mode=mode)))
    assert plain_differs > 100
    ref.lib().ref_skel_set_blend_mode(d["blend_mode"])


def test_oracle_streams_fuzz_against_real_optimised_util():
    """Live fuzz of the remaining OptimisedUtil restatements against the real implementations (which = 0 SSE, 1 General in
    oracle/ref_driver.cpp): random counts incl. 1-17 and counts = 0..3 mod 4, random values incl. signed zeros and large
    magnitudes.  Morph and face normals: bit-identical to both; concat: bit-identical per association; light facing: identical
    per summation order; extrusion: bit-identical to General, within the rsqrt bound of SSE."""
    from ogre_b200 import addr
    from datagen import edge_triangles
    L, R = oracle.lib(), ref.lib()
    SSE, GEN = 0, 1
    rng = np.random.default_rng(77)
    sizes = list(range(1, 18)) + [31, 64, 257, 1000, 4099]

    def values(shape):
        v = (rng.normal(size=shape) * 10.0 ** rng.integers(-3, 4, size=shape)).astype(np.float32)
        v[rng.random(shape) < 0.02] = 0.0
        v[rng.random(shape) < 0.01] = -0.0
        return v
    for n in sizes:
        # morph, positions only and positions + normals (General normalises the morphed normals exactly)
        for normals in (0, 1):
            c = 6 if normals else 3
            a, b = values((n, c)), values((n, c))
            t = np.float32(rng.uniform(0, 1))
            d, g, s = np.zeros_like(a), np.zeros_like(a), np.zeros_like(a)
            L.oracle_morph(t, addr(a), addr(b), addr(d), 4 * c, 4 * c, 4 * c, n, normals)
            R.ref_morph(GEN, t, addr(a), addr(b), addr(g), 4 * c, 4 * c, 4 * c, n, normals)
            R.ref_morph(SSE, t, addr(a), addr(b), addr(s), 4 * c, 4 * c, 4 * c, n, normals)
            assert np.array_equal(bits(d), bits(g)), ("morph general", n, normals)
            assert np.array_equal(bits(d[:, :3]), bits(s[:, :3])), ("morph sse positions", n, normals)
        # concatenateAffineMatrices
        base = values((16,)); base[12:] = (0, 0, 0, 1)
        src = values((n, 16)); src[:, 12:] = (0, 0, 0, 1)
        for which, fn in ((SSE, L.oracle_concat_affine_sse), (GEN, L.oracle_concat_affine_general)):
            d, r = np.zeros_like(src), np.zeros_like(src)
            fn(addr(base), addr(src), addr(d), n)
            R.ref_concat(which, addr(base), addr(src), addr(r), n)
            assert np.array_equal(bits(d[:, :12]), bits(r[:, :12])), ("concat", which, n)
        # face normals, light facing, extrusion
        nv = max(3, n)
        pos = values((nv, 3))
        tris = edge_triangles(rng, n, nv)
        fn_o, fn_g, fn_s = (np.zeros((n, 4), np.float32) for _ in range(3))
        L.oracle_face_normals(addr(pos), addr(tris), addr(fn_o), n)
        R.ref_face_normals(GEN, addr(pos), addr(tris), addr(fn_g), n)
        R.ref_face_normals(SSE, addr(pos), addr(tris), addr(fn_s), n)
        assert np.array_equal(bits(fn_o), bits(fn_g)) and np.array_equal(bits(fn_o), bits(fn_s)), ("face normals", n)
        for light in (np.array([*rng.uniform(-50, 50, 3), 1.0], np.float32), np.array([*rng.normal(size=3), 0.0], np.float32)):
            fo, fr = np.zeros(n, np.int8), np.zeros(n, np.int8)
            L.oracle_light_facing_sse(addr(light), addr(fn_o), addr(fo), n); R.ref_light_facing(SSE, addr(light), addr(fn_o), addr(fr), n)
            assert np.array_equal(fo, fr), ("facing sse", n)
            L.oracle_light_facing_general(addr(light), addr(fn_o), addr(fo), n); R.ref_light_facing(GEN, addr(light), addr(fn_o), addr(fr), n)
            assert np.array_equal(fo, fr), ("facing general", n)
            pts = rng.uniform(-20, 20, (nv, 3)).astype(np.float32)
            eo, eg, es = (np.zeros((nv, 3), np.float32) for _ in range(3))
            L.oracle_extrude(addr(light), 100.0, addr(pts), addr(eo), nv)
            R.ref_extrude(GEN, addr(light), 100.0, addr(pts), addr(eg), nv)
            R.ref_extrude(SSE, addr(light), 100.0, addr(pts), addr(es), nv)
            assert np.array_equal(bits(eo), bits(eg)), ("extrude general", n)
            assert np.abs(eo - es).max() <= 3.7e-4 * 100.0, ("extrude sse", n)


def test_oracle_skinning_fuzz_all_layouts_against_real_sse_and_general():
    """Live fuzz of both skinning restatements over every layout the SSE dispatcher distinguishes (shared-packed, separated,
    positions only, general strides), misaligned heads, counts 1-17 and larger, 1-4 weights: positions bit-identical to the
    implementation restated (SSE: collapse then transform; General: transform then blend), General normals bit-identical,
    SSE normals within its rsqrt bound, bytes outside the written elements untouched by all four."""
    l = ref.lib()
    rng = np.random.default_rng(5)
    for trial in range(160):
        layout = ("shared", "separate", "posonly", "fat")[trial % 4]
        n = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 33, 100, 1001]))
        nw = int(rng.integers(1, 5))
        head = int(rng.choice([0, 4, 8, 12]))
        c = SkinCase(seed=1000 + trial, n=n, nw=nw, layout=layout, head=head, nbones=int(rng.integers(1, 65)),
                     blend_interleaved=bool(trial % 3))
        keep, pal = oracle.palette_ptrs(c.mats)
        for kind, which in (("sse", ref.SSE), ("general", ref.GENERAL)):
            dst, dstn = c.new_dst(); rdst, rdstn = c.new_dst()
            sp, dp, sn, dn, bw, bi = c.pointers(dst, dstn)
            oracle.skin(kind, sp, dp, sn, dn, bw, bi, c.mats, c.strides(), nw, n)
            sp, dp, sn, dn, bw, bi = c.pointers(rdst, rdstn)
            l.ref_skin(which, sp, dp, sn, dn, bw, bi, pal, *c.strides(), nw, n)
            po, no = c.dst_views(dst, dstn); pr, nr = c.dst_views(rdst, rdstn)
            assert np.array_equal(bits(po), bits(pr)), (trial, kind, layout, n, nw, head)
            if no is not None:
                if kind == "general":
                    assert np.array_equal(bits(no), bits(nr)), (trial, kind)
                else:
                    assert np.abs(no - nr).max() <= 4e-4, (trial, kind)
            mask = np.ones(len(dst), bool)
            for v in range(n):
                o = c.head + v * c.dp_stride
                mask[o:o + (24 if layout in ("shared", "fat") else 12)] = False
            assert np.array_equal(dst[mask], rdst[mask]), (trial, kind, "bytes outside the vertices")
