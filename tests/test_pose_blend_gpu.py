"""GPU parity of pose blending (ogrecu_pose_blend_dev / _instanced_dev, SURVEY §8f rank 3: Mesh::softwareVertexPoseBlend,
OgreMain/src/OgreMesh.cpp:1948-2007) against the CPU oracle (pinned against the real static method in
tests/test_oracle_vs_ref.py) — bit patterns."""
import numpy as np
import pytest

import oracle
import ogre_b200
from datagen import bits, pose_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,nv,frac,normals", [(1, 4000, 0.3, True), (2, 17, 1.0, True), (3, 100000, 0.01, False), (4, 1, 1.0, True)])
def test_pose_blend_bit_exact(gpu_ctx, seed, nv, frac, normals):
    from ogre_b200 import check, lib
    ctx = gpu_ctx
    buf, oi, off, ni, nrm = pose_case(seed, nv, frac, normals)
    fpv = buf.shape[1]
    doi, doff = ctx.to_device(oi), ctx.to_device(off)
    dni, dnrm = (ctx.to_device(ni), ctx.to_device(nrm)) if normals else (None, None)
    for w in (0.0, 1.0, 0.37, -0.5, 1e-30):
        ref = buf.copy()
        oracle.lib().oracle_pose_blend(w, oi.ctypes.data, off.ctypes.data, len(oi), ref.ctypes.data, fpv)
        if normals:
            oracle.lib().oracle_pose_blend(w, ni.ctypes.data, nrm.ctypes.data, len(ni), ref.ctypes.data + 12, fpv)
        d = ctx.to_device(buf)
        check(lib().ogrecu_pose_blend_dev(ctx.handle, w, doi.ptr, doff.ptr, len(oi), d.ptr, fpv * 4))
        if normals:
            check(lib().ogrecu_pose_blend_dev(ctx.handle, w, dni.ptr, dnrm.ptr, len(ni), d.ptr + 12, fpv * 4))
        got = d.download(np.float32).reshape(nv, fpv)
        assert np.array_equal(bits(got), bits(ref)), (seed, w)
        d.free()


def test_two_poses_in_sequence_and_instances(gpu_ctx):
    """Several poses are applied one after the other (VertexAnimationTrack::applyPoseToVertexData per pose reference);
    the instanced form gives every instance its own weight, zero weights leave the instance untouched."""
    from ogre_b200 import check, lib
    ctx = gpu_ctx
    nv, ninst = 5000, 37
    buf, oi1, off1, ni1, nrm1 = pose_case(11, nv, 0.4, True)
    _, oi2, off2, ni2, nrm2 = pose_case(12, nv, 0.2, True)
    rng = np.random.default_rng(5)
    w1 = rng.uniform(0, 1, ninst).astype(np.float32); w2 = rng.uniform(0, 1, ninst).astype(np.float32)
    w1[::5] = 0.0; w2[3::7] = 0.0
    inst = np.stack([buf + np.float32(i) for i in range(ninst)]).astype(np.float32)
    d = ctx.to_device(inst)
    dw1, dw2 = ctx.to_device(w1), ctx.to_device(w2)
    devs = [ctx.to_device(a) for a in (oi1, off1, ni1, nrm1, oi2, off2, ni2, nrm2)]
    for (dw, k) in ((dw1, 0), (dw2, 4)):
        idx, off, nidx, nrm = devs[k:k + 4]
        n_off, n_nrm = (len(oi1), len(ni1)) if k == 0 else (len(oi2), len(ni2))
        check(lib().ogrecu_pose_blend_instanced_dev(ctx.handle, dw.ptr, idx.ptr, off.ptr, n_off, d.ptr, 24, ninst, nv * 24))
        check(lib().ogrecu_pose_blend_instanced_dev(ctx.handle, dw.ptr, nidx.ptr, nrm.ptr, n_nrm, d.ptr + 12, 24, ninst, nv * 24))
    got = d.download(np.float32).reshape(ninst, nv, 6)
    for i in range(ninst):
        ref = inst[i].copy()
        for w, (oi, off, ni, nrm) in ((w1[i], (oi1, off1, ni1, nrm1)), (w2[i], (oi2, off2, ni2, nrm2))):
            oracle.lib().oracle_pose_blend(float(w), oi.ctypes.data, off.ctypes.data, len(oi), ref.ctypes.data, 6)
            oracle.lib().oracle_pose_blend(float(w), ni.ctypes.data, nrm.ctypes.data, len(ni), ref.ctypes.data + 12, 6)
        assert np.array_equal(bits(got[i]), bits(ref)), i


@pytest.mark.parametrize("seed,nv,nposes", [(1, 500, 3), (4, 20000, 5), (3, 37, 1)])
def test_whole_pose_frame_many_instances(gpu_ctx, seed, nv, nposes):
    """A frame of pose animation for many instances on the device: copy of the base mesh, normals reset
    (ogrecu_pose_normals_reset_dev), one instanced blend per pose and element with per-instance influences, then
    ogrecu_pose_normals_finalise_dev — against the oracle chain (which equals the real Entity, tests/test_oracle_vs_ref.py)."""
    from datagen import pose_entity_case, pose_entity_oracle
    from ogre_b200 import check, lib
    ctx = gpu_ctx
    ninst = 9
    base, offs, idx, off, nrm, infl = pose_entity_case(seed, nv, nposes)
    rng = np.random.default_rng(seed + 100)
    inf_i = (infl[None, :] * rng.uniform(0.3, 1.0, size=(ninst, nposes))).astype(np.float32)
    inf_i[0] = infl
    inf_i[2, 0] = 0.0
    d = ctx.to_device(np.stack([base] * ninst))
    dbase, didx, doff, dnrm = ctx.to_device(base), ctx.to_device(idx), ctx.to_device(off), ctx.to_device(nrm)
    dinf = ctx.to_device(np.ascontiguousarray(inf_i.T))                 # [pose][instance]
    L = lib()
    check(L.ogrecu_pose_normals_reset_dev(ctx.handle, d.ptr + 12, nv, 24, ninst, nv * 24))
    for p in range(nposes):
        a, b = int(offs[p]), int(offs[p + 1])
        w = dinf.ptr + p * ninst * 4
        check(L.ogrecu_pose_blend_instanced_dev(ctx.handle, w, didx.ptr + a * 4, doff.ptr + a * 12, b - a, d.ptr, 24, ninst, nv * 24))
        check(L.ogrecu_pose_blend_instanced_dev(ctx.handle, w, didx.ptr + a * 4, dnrm.ptr + a * 12, b - a, d.ptr + 12, 24, ninst, nv * 24))
    check(L.ogrecu_pose_normals_finalise_dev(ctx.handle, d.ptr + 12, dbase.ptr + 12, nv, 24, 24, ninst, nv * 24))
    got = d.download(np.float32).reshape(ninst, nv, 6)
    for i in range(ninst):
        ref = pose_entity_oracle(oracle, base, offs, idx, off, nrm, inf_i[i], True)
        assert np.array_equal(bits(got[i]), bits(ref)), i
