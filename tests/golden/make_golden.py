#!/usr/bin/env python3
"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref, built
by oracle/build_ref.py + oracle/build_driver.sh from /root/reference).  Run in the build
container only; the GPU box and the CPU test suite just read the .npz files.

  ogrehead_1280x720.npz  every TinyRenderSystem::_render call of the SURVEY §8d C1 scene
                         (buffers, state, DefaultShader uniforms, textures) + Tiny's own
                         colour and depth images
  ogrehead_320x180.npz   same scene, small target (fast CPU test)
  sinbad_skin.npz        Sinbad.mesh (7 sub-meshes, 69 bones): source/blend streams, bone
  jaiqua_skin.npz        maps, 32 bone palettes sampled from the first animation, and OGRE's
                         own software-skinned vertices (Entity::_updateAnimation with the SSE
                         OptimisedUtil) for two of the poses
  sinbad_skeleton.npz    the skeletons themselves (bone table, binding pose, the engine's own binding-pose
  jaiqua_skeleton.npz    inverses, every animation's node tracks and key frames) + 96 poses each: 1-3 enabled
                         animation states (order, time position, weight, both blend modes) and the palette
                         the real SkeletonInstance::setAnimationState + _getBoneMatrices produced
  pose_blend.npz         outputs of the real Mesh::softwareVertexPoseBlend on seeded poses (datagen.pose_case)
  jaiqua_shadow.npz      the meshes' edge lists (Mesh::buildEdgeList) + 14 shadow-volume cases each (lights, bounds,
  sinbad_shadow.npz      flag combinations, shared-buffer offsets incl. a wrap): light facings from the real EdgeData and the
                         index buffer contents / ShadowRenderable ranges the real ShadowCaster::generateShadowVolume wrote
  image_scale.npz        outputs of the real Image::scale(FILTER_BILINEAR) on seeded RGB8 images (datagen.SCALE_CASES: up- and
                         down-scaling, sub-boxes, 1-pixel extents), same format and into PF_A8R8G8B8 / PF_BYTE_RGBA
  skel_frames.npz        the frames the real Tiny plugin renders of the software-skinned Sinbad / jaiqua entities (poses at t = 0.3 s
                         and 1.7 s, 256x256) and of ogrehead at 400x300 from z = 90, through tests/harness/run_scene.py on the
                         build host
  optutil_golden.npz     outputs of the real OptimisedUtilSSE / OptimisedUtilGeneral on small
                         seeded inputs (all six functions, edge-case counts and alignments)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from oracle import ref  # noqa: E402
from datagen import SkinCase, edge_triangles, rigid_matrices  # noqa: E402


def save_scene(name, mesh, w, h):
    colour, z, draws, texs, stats = ref.capture_scene(mesh, w, h)
    out = {"colour": colour, "depth": z, "ndraws": np.array(len(draws)), "ntex": np.array(len(texs)),
           "triangles": np.array(stats["triangles"]), "size": np.array([w, h])}
    for t, tx in enumerate(texs):
        out["tex%d" % t] = tx
    for i, d in enumerate(draws):
        for k, v in d.items():
            out["d%d_%s" % (i, k)] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, stats, "covered %.3f" % (z != 1).mean())


def save_skeleton(name, mesh, ntimes=32, dt=0.1):
    out = {}
    pals = []
    for k in range(ntimes):
        bones, subs = ref.skeleton_pose(mesh, k * dt, reopen=(k == 0))
        pals.append(bones)
        if k == 0:
            out["nsubs"] = np.array(len(subs))
            for i, s in enumerate(subs):
                n = s["nverts"]
                out["s%d_src" % i] = s["src"][:n * s["src_stride"]]
                out["s%d_blend" % i] = s["blend"][:n * s["blend_stride"]]
                out["s%d_bone_map" % i] = s["bone_map"]
                out["s%d_meta" % i] = np.array([s[x] for x in ("src_stride", "norm_off", "blend_stride", "weight_off",
                                                               "index_off", "nw", "nverts", "has_norm", "dst_stride")], dtype=np.uint32)
        if k in (3, 17):
            for i, s in enumerate(subs):
                out["s%d_skinned_t%d" % (i, k)] = s["skinned"][:s["nverts"] * s["dst_stride"]]
    out["palettes"] = np.stack(pals).astype(np.float32)     # [ntimes, nbones, 16]
    out["times"] = np.arange(ntimes, dtype=np.float32) * dt
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, out["palettes"].shape, [int(out["s%d_meta" % i][6]) for i in range(len(subs))])


def save_skeleton_def(name, mesh, nposes=96, seed=7):
    d = ref.skeleton_def(mesh)
    out = {"num_bones": np.array(d["num_bones"]), "blend_mode": np.array(d["blend_mode"]), "parent": d["parent"], "bind": d["bind"],
           "inherit": d["inherit"], "bind_inverse": d["bind_inverse"], "nanims": np.array(len(d["anims"]))}
    for i, a in enumerate(d["anims"]):
        out["a%d_name" % i] = np.array(a["name"])
        out["a%d_meta" % i] = np.array([a["length"], a["interpolation"], a["rotation_interpolation"]], np.float32)
        for k in ("track_bone", "track_shortest", "key_offset", "keys"):
            out["a%d_%s" % (i, k)] = a[k]
    rng = np.random.default_rng(seed)
    S = 3
    pa = -np.ones((nposes, S), np.int32); pt = np.zeros((nposes, S), np.float32); pw = np.zeros((nposes, S), np.float32)
    pn = np.zeros(nposes, np.int32); pm = np.zeros(nposes, np.int32); pal = np.zeros((nposes, d["num_bones"], 16), np.float32)
    for p in range(nposes):
        ns = 1 if p < nposes // 3 else int(rng.integers(1, S + 1))
        an = rng.choice(len(d["anims"]), ns, replace=False)
        tm = np.array([rng.uniform(0, d["anims"][a]["length"]) for a in an], np.float32)
        if p % 7 == 0:
            tm[0] = d["anims"][an[0]]["length"]           # exactly the last key
        if p % 11 == 0:
            tm[0] = 0.0
        if p % 13 == 0 and len(d["anims"][an[0]]["keys"]) > 3:
            tm[0] = d["anims"][an[0]]["keys"][2, 0]        # exactly on a key time
        w = np.ones(ns, np.float32) if p % 3 == 0 else rng.uniform(0.05, 1.0, ns).astype(np.float32)
        if p % 17 == 5:
            w[-1] = 0.0                                    # zero weight: the track is skipped
        mode = int(p % 4 == 3)
        ref.lib().ref_skel_set_blend_mode(mode)
        pal[p] = ref.skeleton_pose_states(d["num_bones"], an, tm, w)
        pa[p, :ns], pt[p, :ns], pw[p, :ns], pn[p], pm[p] = an, tm, w, ns, mode
    ref.lib().ref_skel_set_blend_mode(d["blend_mode"])
    out.update(pose_anim=pa, pose_time=pt, pose_weight=pw, pose_nstates=pn, pose_blend_mode=pm, pose_palette=pal)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, d["num_bones"], "bones", len(d["anims"]), "animations", nposes, "poses")


def save_shadow(name, mesh, ncases=14, seed=4):
    """Edge list of `mesh` + shadow-volume cases: light, light-cap bounds, flags -> the light facings the real EdgeData
    computed and the indices / ranges the real ShadowCaster::generateShadowVolume wrote."""
    ed = ref.shadow_open(mesh)
    out = {"groups": ed["groups"], "edges": ed["edges"], "tris": ed["tris"], "ncases": np.array(ncases)}
    rng = np.random.default_rng(seed)
    for c in range(ncases):
        if c % 3 == 0:
            v = rng.normal(size=3); light = np.array([*(v / np.linalg.norm(v)), 0], np.float32)
        else:
            light = np.array([*(rng.uniform(-1, 1, 3) * rng.choice([5, 40, 300])), 1], np.float32)
        half = rng.uniform(3, 40, 3); ctr = rng.uniform(-5, 5, 3)
        bounds = np.array([*(ctr - half), *(ctr + half)], np.float32)
        flags = c % 8
        sep = bool(c % 2)
        cap = 120000
        used = [0, 1000, cap - 500][c % 3]
        idx, ranges, used_out, fac = ref.shadow_volume(ed, light, bounds, flags, sep, cap, used)
        oi, orr, ou, n, mc = oracle.shadow_volume(ed, fac, light, bounds, flags, sep, cap, used)     # the emitted count comes from the restatement
        assert np.array_equal(oi, idx) and ou == used_out, "oracle and reference disagree: fix that first"
        start = used_out - n
        out["c%d_meta" % c] = np.array([flags, int(sep), cap, used, used_out, n, mc], np.int64)
        out["c%d_light" % c], out["c%d_bounds" % c], out["c%d_facings" % c] = light, bounds, fac
        out["c%d_ranges" % c], out["c%d_indices" % c] = ranges, idx[start:start + n]
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, len(ed["groups"]), "groups", len(ed["edges"]), "edges", ncases, "cases")


def save_pose_blend(name):
    """Mesh::softwareVertexPoseBlend (the real static method) on seeded poses: inputs are regenerated by
    datagen.pose_case(seed, ...), the outputs are stored."""
    from datagen import pose_case
    l = ref.lib()
    out = {}
    cases = [(1, 4000, 0.3, 1), (2, 17, 1.0, 1), (3, 1000, 0.01, 0), (4, 1, 1.0, 1)]
    out["cases"] = np.array(cases, np.float64)
    for k, (seed, nv, frac, normals) in enumerate(cases):
        buf, oi, off, ni, nrm = pose_case(seed, nv, frac, bool(normals))
        for j, w in enumerate((0.0, 1.0, 0.37, -0.5)):
            r = buf.copy()
            assert l.ref_pose_blend(w, len(oi), oi.ctypes.data, off.ctypes.data, len(ni), ni.ctypes.data, nrm.ctypes.data, r.ctypes.data, nv,
                                    buf.shape[1], normals) == 0
            out["c%d_w%d" % (k, j)] = r
    # whole pose-animated frames of a real Entity (initialisePoseVertexData -> blends -> finalisePoseNormals)
    from datagen import pose_entity_case
    frames = [(1, 500, 3, 1), (2, 500, 3, 0), (3, 37, 1, 1), (4, 2000, 5, 1)]
    out["frames"] = np.array(frames, np.int64)
    for k, (seed, nv, nposes, normals) in enumerate(frames):
        base, offs, idx, off, nrm, infl = pose_entity_case(seed, nv, nposes)
        res = np.zeros((nv, 6), np.float32)
        assert l.ref_pose_entity(nv, base.ctypes.data, nposes, offs.ctypes.data, idx.ctypes.data, off.ctypes.data, nrm.ctypes.data,
                                 infl.ctypes.data, normals, res.ctypes.data) == 0
        out["f%d" % k] = res
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, len(cases), "cases", len(frames), "entity frames")


def save_image_scale(name):
    from datagen import SCALE_CASES, scale_image
    out = {}
    for k, (seed, h, w, box, dw, dh) in enumerate(SCALE_CASES):
        src = scale_image(seed, h, w)
        out["c%d_rgb" % k] = ref.image_scale(src, dw, dh, "PF_BYTE_RGB", box=box, dst_pad=k % 3)
        for fmt in ("PF_A8R8G8B8", "PF_BYTE_RGBA"):
            out["c%d_%s" % (k, fmt)] = ref.image_scale(src, dw, dh, "PF_BYTE_RGB", fmt, box=box)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, len(out), "arrays")


def save_skel_frames(name):
    import subprocess
    import tempfile
    out = {}
    for mesh in ("Sinbad.mesh", "jaiqua.mesh"):
        with tempfile.TemporaryDirectory() as tmp:
            f = os.path.join(tmp, "t.npz")
            subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "harness", "run_scene.py"), "--plugin", "tiny", "--mesh", mesh,
                                   "--size", "256x256", "--cam-z", "30", "--skel", "0.3,1.7", "--out", f],
                                  env=dict(os.environ, MALLOC_PERTURB_="255"))
            out[mesh] = np.load(f)["colour"]
    # the side-by-side scene of tests/test_plugin_gpu.py: ogrehead, 400x300, camera at z = 90
    with tempfile.TemporaryDirectory() as tmp:
        f = os.path.join(tmp, "t.npz")
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "harness", "run_scene.py"), "--plugin", "tiny", "--mesh", "ogrehead.mesh",
                               "--size", "400x300", "--cam-z", "90", "--out", f], env=dict(os.environ, MALLOC_PERTURB_="255"))
        out["ogrehead_400x300_z90"] = np.load(f)["colour"]
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, len(out), "arrays")


def save_optutil(name):
    l = ref.lib()
    rng = np.random.default_rng(2024)
    out = {}
    # skinning: every layout / weight count / ragged count / misaligned head the SSE dispatcher distinguishes
    k = 0
    for layout in ("shared", "separate", "posonly", "fat"):
        for nw in (1, 2, 3, 4):
            for n, head in ((1, 0), (5, 8), (17, 4), (18, 0), (103, 12)):
                c = SkinCase(seed=1000 + k, n=n, nw=nw, layout=layout, head=head, nbones=24)
                for which, tag in ((ref.SSE, "sse"), (ref.GENERAL, "gen")):
                    dst, dstn = c.new_dst()
                    sp, dp, sn, dn, bw, bi = c.pointers(dst, dstn)
                    keep, pal = oracle.palette_ptrs(c.mats)
                    l.ref_skin(which, sp, dp, sn, dn, bw, bi, pal, *c.strides(), nw, n)
                    out["skin%d_%s" % (k, tag)] = dst
                    if dstn is not None:
                        out["skin%d_%s_n" % (k, tag)] = dstn
                out["skin%d_args" % k] = np.array([1000 + k, n, nw, head, ("shared", "separate", "posonly", "fat").index(layout)])
                k += 1
    out["nskin"] = np.array(k)
    # morph
    for j, (n, normals) in enumerate(((1, 0), (7, 0), (64, 0), (1, 1), (9, 1), (130, 1))):
        vs = 24 if normals else 12
        a = rng.uniform(-10, 10, size=n * vs // 4).astype(np.float32)
        b = rng.uniform(-10, 10, size=n * vs // 4).astype(np.float32)
        for which, tag in ((ref.SSE, "sse"), (ref.GENERAL, "gen")):
            d = np.zeros_like(a)
            l.ref_morph(which, 0.37, a.ctypes.data, b.ctypes.data, d.ctypes.data, vs, vs, vs, n, normals)
            out["morph%d_%s" % (j, tag)] = d
        out["morph%d_a" % j], out["morph%d_b" % j] = a, b
        out["morph%d_args" % j] = np.array([n, normals])
    # concat
    base = rigid_matrices(rng, 1)[0]; src = rigid_matrices(rng, 69)
    out["concat_base"], out["concat_src"] = base, src
    for which, tag in ((ref.SSE, "sse"), (ref.GENERAL, "gen")):
        d = np.tile(np.array([0, 0, 0, 1], np.float32), (69, 4)).reshape(69, 16).copy()
        l.ref_concat(which, base.ctypes.data, src.ctypes.data, d.ctypes.data, 69)
        out["concat_" + tag] = d
    # face normals / light facing / extrude
    nv, nt = 500, 1003
    pos = rng.uniform(-10, 10, size=(nv, 3)).astype(np.float32)
    tris = edge_triangles(rng, nt, nv)
    out["fn_pos"], out["fn_tris"] = pos, tris
    for which, tag in ((ref.SSE, "sse"), (ref.GENERAL, "gen")):
        fn = np.zeros((nt, 4), np.float32)
        l.ref_face_normals(which, pos.ctypes.data, tris.ctypes.data, fn.ctypes.data, nt)
        out["fn_" + tag] = fn
        for li, light in enumerate(([30.0, 2.0, -11.0, 1.0], [0.3, -0.9, 0.1, 0.0])):
            lv = np.array(light, np.float32)
            fac = np.zeros(nt, np.int8)
            l.ref_light_facing(which, lv.ctypes.data, fn.ctypes.data, fac.ctypes.data, nt)
            out["facing%d_%s" % (li, tag)] = fac
            for cnt in (nv, 3, 2, 1):
                ex = np.zeros((cnt, 3), np.float32)
                l.ref_extrude(which, lv.ctypes.data, 250.0, pos.ctypes.data, ex.ctypes.data, cnt)
                out["extrude%d_%d_%s" % (li, cnt, tag)] = ex
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, len(out), "arrays")


if __name__ == "__main__":
    if not ref.available():
        sys.exit("oracle/_ref is not built: run oracle/build_ref.py and oracle/build_driver.sh first")
    if sys.argv[1:] == ["image_scale"]:
        save_image_scale("image_scale.npz")
        sys.exit(0)
    if sys.argv[1:] == ["skel_frames"]:
        save_skel_frames("skel_frames.npz")
        sys.exit(0)
    save_optutil("optutil_golden.npz")
    save_image_scale("image_scale.npz")
    save_skel_frames("skel_frames.npz")
    save_scene("ogrehead_320x180.npz", "ogrehead.mesh", 320, 180)
    save_scene("ogrehead_1280x720.npz", "ogrehead.mesh", 1280, 720)
    save_skeleton("sinbad_skin.npz", "Sinbad.mesh")
    save_skeleton("jaiqua_skin.npz", "jaiqua.mesh")
    save_skeleton_def("sinbad_skeleton.npz", "Sinbad.mesh")
    save_skeleton_def("jaiqua_skeleton.npz", "jaiqua.mesh")
    save_pose_blend("pose_blend.npz")
    save_shadow("jaiqua_shadow.npz", "jaiqua.mesh")
    save_shadow("sinbad_shadow.npz", "Sinbad.mesh")
