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
    save_optutil("optutil_golden.npz")
    save_scene("ogrehead_320x180.npz", "ogrehead.mesh", 320, 180)
    save_scene("ogrehead_1280x720.npz", "ogrehead.mesh", 1280, 720)
    save_skeleton("sinbad_skin.npz", "Sinbad.mesh")
    save_skeleton("jaiqua_skin.npz", "jaiqua.mesh")
