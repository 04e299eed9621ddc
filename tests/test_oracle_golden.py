"""CPU: the oracle restatement vs golden vectors dumped from the UNMODIFIED reference
(tests/golden/make_golden.py).  This is what pins oracle/ogre_oracle.c; it needs neither a GPU
nor /root/reference."""
import os

import numpy as np
import pytest

import oracle
from datagen import SkinCase, bits
from goldenlib import GOLDEN, SkelFixture, load_scene
from rasterlib import assert_images_equal, render_oracle

LAYOUTS = ("shared", "separate", "posonly", "fat")


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(GOLDEN, "optutil_golden.npz"))


def test_ogrehead_oracle_replay_is_image_identical_to_tiny():
    for name in ("ogrehead_320x180.npz", "ogrehead_1280x720.npz"):
        cases, colour, depth, (w, h) = load_scene(name)
        assert sum(c.draw_count // 3 for c in cases) == 4336 and len(cases) == 4
        oc, oz, frags = render_oracle(cases, w, h, clear=(0.1, 0.2, 0.3, 1.0), depth=1.0)
        assert_images_equal(oc, oz, colour, depth, name)


def test_skinning_oracle_vs_real_sse_and_general(G):
    worst_n = 0.0
    for k in range(int(G["nskin"])):
        seed, n, nw, head, li = [int(v) for v in G["skin%d_args" % k]]
        c = SkinCase(seed=seed, n=n, nw=nw, layout=LAYOUTS[li], head=head, nbones=24)
        for kind, tag in (("sse", "sse"), ("general", "gen")):
            dst, dstn = c.new_dst()
            sp, dp, sn, dn, bw, bi = c.pointers(dst, dstn)
            oracle.skin(kind, sp, dp, sn, dn, bw, bi, c.mats, c.strides(), nw, n)
            ref = G["skin%d_%s" % (k, tag)]
            refn = G["skin%d_%s_n" % (k, tag)] if ("skin%d_%s_n" % (k, tag)) in G else None
            po, no = c.dst_views(dst, dstn)
            pr, nr = c.dst_views(ref, refn)
            # positions: bit-identical to the implementation whose formulation is restated
            assert np.array_equal(bits(po), bits(pr)), (k, tag)
            if no is not None:
                if tag == "gen":
                    assert np.array_equal(bits(no), bits(nr)), (k, tag)
                else:   # SSE normalises with the ~12-bit rsqrt estimate (<= 4e-4)
                    worst_n = max(worst_n, float(np.abs(no - nr).max()))
                    assert np.abs(no - nr).max() <= 4e-4
            # bytes outside the written elements are untouched in both
            mask = np.ones(len(dst), bool)
            for v in range(n):
                o = c.head + v * c.dp_stride
                mask[o:o + (24 if c.layout in ("shared", "fat") else 12)] = False
            assert np.array_equal(dst[mask], ref[mask])
    assert worst_n > 1e-6      # the SSE estimate really is approximate on this CPU


def test_morph_oracle_vs_reference(G):
    from ogre_b200 import addr
    for j in range(6):
        n, normals = [int(v) for v in G["morph%d_args" % j]]
        vs = 24 if normals else 12
        a, b = G["morph%d_a" % j].copy(), G["morph%d_b" % j].copy()
        d = np.zeros_like(a)
        oracle.lib().oracle_morph(0.37, addr(a), addr(b), addr(d), vs, vs, vs, n, normals)
        assert np.array_equal(bits(d), bits(G["morph%d_gen" % j]))                 # General: exact
        if not normals:
            assert np.array_equal(bits(d), bits(G["morph%d_sse" % j]))             # SSE positions: exact
        else:   # SSE morph normals are wrong (length^2 = x^2+y^2+2z^2): positions still exact
            dv, sv = d.reshape(n, 6), G["morph%d_sse" % j].reshape(n, 6)
            assert np.array_equal(bits(dv[:, :3]), bits(sv[:, :3]))


def test_concat_face_normals_facing_extrude_vs_reference(G):
    from ogre_b200 import addr
    base, src = G["concat_base"].copy(), G["concat_src"].copy()
    d = np.zeros_like(src)
    oracle.lib().oracle_concat_affine_sse(addr(base), addr(src), addr(d), len(src))
    assert np.array_equal(bits(d), bits(G["concat_sse"]))
    oracle.lib().oracle_concat_affine_general(addr(base), addr(src), addr(d), len(src))
    assert np.array_equal(bits(d), bits(G["concat_gen"]))
    pos, tris = G["fn_pos"].copy(), G["fn_tris"].copy()
    nt, nv = len(tris), len(pos)
    fn = np.zeros((nt, 4), np.float32)
    oracle.lib().oracle_face_normals(addr(pos), addr(tris), addr(fn), nt)
    assert np.array_equal(bits(fn), bits(G["fn_sse"])) and np.array_equal(bits(fn), bits(G["fn_gen"]))
    for li, light in enumerate(([30.0, 2.0, -11.0, 1.0], [0.3, -0.9, 0.1, 0.0])):
        lv = np.array(light, np.float32)
        fac = np.zeros(nt, np.int8)
        oracle.lib().oracle_light_facing_sse(addr(lv), addr(fn), addr(fac), nt)
        assert np.array_equal(fac, G["facing%d_sse" % li])
        oracle.lib().oracle_light_facing_general(addr(lv), addr(fn), addr(fac), nt)
        assert np.array_equal(fac, G["facing%d_gen" % li])
        for cnt in (nv, 3, 2, 1):
            ex = np.zeros((cnt, 3), np.float32)
            oracle.lib().oracle_extrude(addr(lv), 250.0, addr(pos), addr(ex), cnt)
            assert np.array_equal(bits(ex), bits(G["extrude%d_%d_gen" % (li, cnt)]))
            # SSE extrusion uses rsqrt: abs error <= 3.7e-4 * extrudeDist
            assert np.abs(ex - G["extrude%d_%d_sse" % (li, cnt)]).max() <= 3.7e-4 * 250.0


@pytest.mark.parametrize("name", ["sinbad_skin.npz", "jaiqua_skin.npz"])
def test_oracle_reproduces_ogres_own_skinned_entities(name):
    """End to end through the real call site: Entity::_updateAnimation -> Mesh::softwareVertexBlend
    -> OptimisedUtilSSE.  Bone palette = bones[bone_map] (Mesh::prepareMatricesForVertexBlend)."""
    F = SkelFixture(name)
    for t in (3, 17):
        for s in F.subs:
            n, nw = s["nverts"], s["nw"]
            pal = np.ascontiguousarray(F.palettes[t][s["bone_map"]])
            dst = np.zeros(n * 24, np.uint8)
            sp = s["src"].ctypes.data
            bl = s["blend"].ctypes.data
            oracle.skin("sse", sp, dst.ctypes.data, sp + s["norm_off"], dst.ctypes.data + 12,
                        bl + s["weight_off"], bl + s["index_off"], pal,
                        (s["src_stride"], 24, s["src_stride"], 24, s["blend_stride"], s["blend_stride"]), nw, n)
            got = dst.view(np.float32).reshape(n, 6)
            ref = s["skinned"][t].view(np.float32).reshape(n, 6)
            assert np.array_equal(bits(got[:, :3]), bits(ref[:, :3])), (name, t, nw)
            assert np.abs(got[:, 3:] - ref[:, 3:]).max() <= 4e-4


@pytest.mark.parametrize("name", ["sinbad_skeleton.npz", "jaiqua_skeleton.npz"])
def test_oracle_reproduces_ogres_own_bone_palettes(name):
    """Skeleton::setAnimationState + _getBoneMatrices: 96 poses per skeleton dumped from the real
    SkeletonInstance — 1-3 enabled states, both blend modes, times on keys / at 0 / at the end, zero weights — bit for bit;
    the binding-pose inverses the engine computed in Bone::setBindingPose as well."""
    from goldenlib import load_skeleton
    d, poses = load_skeleton(name)
    sk = oracle.Skeleton(d)
    assert np.array_equal(bits(sk.bind_inverse), bits(d["bind_inverse"]))
    for p in range(len(poses["nstates"])):
        n = int(poses["nstates"][p])
        got = sk.palette(poses["anim"][p, :n], poses["time"][p, :n], poses["weight"][p, :n], blend_mode=poses["blend_mode"][p])
        assert np.array_equal(bits(got), bits(poses["palette"][p])), (name, p)


@pytest.mark.parametrize("name", ["jaiqua_shadow.npz", "sinbad_shadow.npz"])
def test_oracle_reproduces_ogres_own_shadow_volume_indices(name):
    """ShadowCaster::generateShadowVolume: 14 cases per mesh dumped from the real engine — every flag
    combination, McGuire fan and full caps, separate light caps, offsets into the shared buffer including a wrap."""
    from goldenlib import load_shadow
    ed, cases = load_shadow(name)
    for k, c in enumerate(cases):
        idx, ranges, used, n, mc = oracle.shadow_volume(ed, c["facings"], c["light"], c["bounds"], c["flags"], c["separate"], c["capacity"],
                                                        c["used_in"])
        assert (n, used, mc) == (c["emitted"], c["used_out"], c["mcguire"]), (name, k)
        assert np.array_equal(idx[used - n:used], c["indices"]), (name, k)
        assert (idx[:used - n] == 0xCDCD).all() and (idx[used:] == 0xCDCD).all()
        cols = 4 if (c["separate"] and c["flags"] & 1) else 2
        assert np.array_equal(ranges[:, :cols], c["ranges"][:, :cols]), (name, k)


def test_oracle_reproduces_ogres_own_pose_blend():
    """Mesh::softwareVertexPoseBlend: outputs of the real static method on seeded poses."""
    from datagen import pose_case
    z = np.load(os.path.join(GOLDEN, "pose_blend.npz"))
    for k, (seed, nv, frac, normals) in enumerate(z["cases"]):
        buf, oi, off, ni, nrm = pose_case(int(seed), int(nv), float(frac), bool(normals))
        for j, w in enumerate((0.0, 1.0, 0.37, -0.5)):
            o = buf.copy()
            oracle.lib().oracle_pose_blend(w, oi.ctypes.data, off.ctypes.data, len(oi), o.ctypes.data, buf.shape[1])
            if normals:
                oracle.lib().oracle_pose_blend(w, ni.ctypes.data, nrm.ctypes.data, len(ni), o.ctypes.data + 12, buf.shape[1])
            assert np.array_equal(bits(o), bits(z["c%d_w%d" % (k, j)])), (k, w)


def test_oracle_reproduces_ogres_own_pose_animated_entities():
    """Whole pose-animated frames dumped from a real Entity (Entity::_updateAnimation: initialisePoseVertexData, the pose
    blends of the key frame in order, finalisePoseNormals) vs the restatements chained."""
    from datagen import pose_entity_case, pose_entity_oracle
    z = np.load(os.path.join(GOLDEN, "pose_blend.npz"))
    for k, (seed, nv, nposes, normals) in enumerate(z["frames"]):
        base, offs, idx, off, nrm, infl = pose_entity_case(int(seed), int(nv), int(nposes))
        got = pose_entity_oracle(oracle, base, offs, idx, off, nrm, infl, bool(normals))
        assert np.array_equal(bits(got), bits(z["f%d" % k])), k


def test_oracle_reproduces_ogres_own_image_scale():
    """Image::scale(FILTER_BILINEAR) on byte images (the scaled blitToMemory read-back): outputs of the real
    method on seeded images — same format (with padded destination rows) and into two 32-bit formats (scale, then the
    bulkPixelConversion tail restated as shifts)."""
    from datagen import SCALE_CASES, PACKED32, pack32, scale_image
    z = np.load(os.path.join(GOLDEN, "image_scale.npz"))
    for k, (seed, h, w, box, dw, dh) in enumerate(SCALE_CASES):
        src = scale_image(seed, h, w)
        got = oracle.scale_bilinear(src, dw, dh, box=box, dst_pad=k % 3)
        assert np.array_equal(got, z["c%d_rgb" % k]), k
        tight = oracle.scale_bilinear(src, dw, dh, box=box)
        assert np.array_equal(pack32(tight, *PACKED32["PF_A8R8G8B8"]), z["c%d_PF_A8R8G8B8" % k].view(np.uint32)[..., 0]), k
        assert np.array_equal(pack32(tight, *PACKED32["PF_A8B8G8R8"]), z["c%d_PF_BYTE_RGBA" % k].view(np.uint32)[..., 0]), k
