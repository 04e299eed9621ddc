"""GPU parity of the device bone-palette production (ogrecu_skel_*) against the golden palettes the
real SkeletonInstance produced (tests/golden/*_skeleton.npz) and against the CPU oracle on seeded random states.
Gate: bit-identical (all arithmetic is float32 in the reference's association, no FMA)."""
import numpy as np
import pytest

import oracle
import ogre_b200
from goldenlib import load_skeleton

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def ctx():
    c = ogre_b200.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("skel_mode", [0, 1])
@pytest.mark.parametrize("name", ["sinbad_skeleton.npz", "jaiqua_skeleton.npz"])
def test_golden_poses_bit_exact(ctx, name, skel_mode):
    # skel_mode 0: the kernel in use (a half-warp = 16 instances of one bone); 1: the first version (a lane per bone)
    ctx.set_config_option("skel_mode", skel_mode)
    d, poses = load_skeleton(name)
    sk = ogre_b200.Skeleton(ctx, d)
    # Bone::setBindingPose restated on the host side of the library
    assert np.array_equal(bits(sk.bind_inverse()), bits(d["bind_inverse"]))
    for mode in (0, 1):
        sel = np.nonzero(poses["blend_mode"] == mode)[0]
        sk.set_blend_mode(mode)
        got = sk.palettes(poses["anim"][sel], poses["time"][sel], poses["weight"][sel])      # disabled slots are -1
        assert np.array_equal(bits(got), bits(poses["palette"][sel])), (name, mode)
    sk.destroy()
    ctx.set_config_option("skel_mode", 0)


@pytest.mark.parametrize("name", ["sinbad_skeleton.npz", "jaiqua_skeleton.npz"])
def test_many_instances_equal_the_oracle(ctx, name):
    d, _ = load_skeleton(name)
    sk = ogre_b200.Skeleton(ctx, d, use_given_bind_inverse=True)
    osk = oracle.Skeleton(d)
    rng = np.random.default_rng(5)
    ni, ns = 3001, 4                                       # ragged against the instances-per-CTA grouping
    na = len(d["anims"])
    anim = rng.integers(-1, na, size=(ni, ns)).astype(np.int32)
    for i in range(ni):                                    # a state set holds an animation at most once
        seen = set()
        for s in range(ns):
            if anim[i, s] in seen:
                anim[i, s] = -1
            seen.add(int(anim[i, s]))
    length = np.array([a["length"] for a in d["anims"]] + [0.0], np.float32)
    time = (rng.uniform(0, 1.3, size=(ni, ns)) * length[anim]).astype(np.float32)          # some beyond the end: wrapped by fmod
    weight = rng.uniform(0, 1, size=(ni, ns)).astype(np.float32)
    weight[rng.uniform(size=(ni, ns)) < 0.1] = 0.0
    weight[rng.uniform(size=(ni, ns)) < 0.2] = 1.0
    for mode in (0, 1):
        sk.set_blend_mode(mode)
        got = sk.palettes(anim, time, weight)
        for i in range(0, ni, 7):
            on = anim[i] >= 0
            want = osk.palette(anim[i][on], time[i][on], weight[i][on], bind_inverse=d["bind_inverse"], blend_mode=mode)
            assert np.array_equal(bits(got[i]), bits(want)), (name, mode, i)
    sk.destroy()


def test_palettes_feed_skinning(ctx):
    """Palettes produced on the device, consumed in place by the instanced skinning kernel: identical to OGRE's own
    skinned vertices of the fixture (Entity::_updateAnimation with the SSE OptimisedUtil) for the same pose."""
    from goldenlib import SkelFixture
    F = SkelFixture("sinbad_skin.npz")
    d, _ = load_skeleton("sinbad_skeleton.npz")
    sk = ogre_b200.Skeleton(ctx, d)
    t = np.float32(3 * 0.1)                                # pose 3 of the fixture: first animation (Dance), time 3 * 0.1
    pal = sk.palettes(np.zeros((2, 1), np.int32), np.full((2, 1), t, np.float32), np.ones((2, 1), np.float32))
    assert np.array_equal(bits(pal[0]), bits(F.palettes[3]))
    sk.destroy()


# IM_SPLINE / RIM_SPHERICAL: the oracle restates them and is pinned BIT FOR BIT against the real SkeletonInstance (same libm, same
# process: tests/test_oracle_vs_ref.py).  The device evaluates Quaternion::Slerp / Squad with CUDA's sinf / atan2f (<= 2 ulp each)
# instead of libm's, so these modes are gated by a TOLERANCE, stated here: every palette element within 2e-5 of the oracle's,
# relative to the largest element of its matrix (the translation column can be tens of units).  What does not go through sin / atan2 is
# still exact: spline translations and scales (Hermite arithmetic on host-made tangents) — checked on poses whose rotation keys
# are all identical below.
SPLINE_TOL = 2e-5


@pytest.mark.parametrize("name", ["sinbad_skeleton.npz", "jaiqua_skeleton.npz"])
@pytest.mark.parametrize("interp,rot", [(0, 1), (1, 0), (1, 1)])
def test_spline_and_spherical_modes_within_tolerance_of_the_oracle(ctx, name, interp, rot):
    d, _ = load_skeleton(name)
    d = dict(d, anims=[dict(a, interpolation=interp, rotation_interpolation=rot) for a in d["anims"]])
    sk = ogre_b200.Skeleton(ctx, d, use_given_bind_inverse=True)
    osk = oracle.Skeleton(d)
    rng = np.random.default_rng(17 + 2 * interp + rot)
    ni, ns = 601, 3
    na = len(d["anims"])
    anim = rng.integers(-1, na, size=(ni, ns)).astype(np.int32)
    for i in range(ni):
        seen = set()
        for s in range(ns):
            if anim[i, s] in seen:
                anim[i, s] = -1
            seen.add(int(anim[i, s]))
    length = np.array([a["length"] for a in d["anims"]] + [0.0], np.float32)
    time = (rng.uniform(0, 1.2, size=(ni, ns)) * length[anim]).astype(np.float32)
    weight = rng.uniform(0, 1, size=(ni, ns)).astype(np.float32)
    weight[rng.uniform(size=(ni, ns)) < 0.2] = 1.0
    worst = 0.0
    for mode in (0, 1):
        sk.set_blend_mode(mode)
        got = sk.palettes(anim, time, weight)
        assert np.isfinite(got).all()
        for i in range(0, ni, 3):
            on = anim[i] >= 0
            want = osk.palette(anim[i][on], time[i][on], weight[i][on], bind_inverse=d["bind_inverse"], blend_mode=mode)
            g = got[i].reshape(-1, 16)[:, :12]; w = np.asarray(want, np.float32).reshape(-1, 16)[:, :12]
            scale = np.maximum(np.abs(w).max(axis=1, keepdims=True), 1.0)
            worst = max(worst, float((np.abs(g - w) / scale).max()))
    assert worst <= SPLINE_TOL, (name, interp, rot, worst)
    sk.destroy()


def test_spline_translation_and_scale_are_exact(ctx):
    """IM_SPLINE with every rotation key of a track equal: Squad degenerates to its nlerp branch (all dot products are 1), so the
    whole palette is sin / atan2-free and must equal the oracle bit for bit — the Hermite evaluation and the host-made tangents"""
    d, _ = load_skeleton("jaiqua_skeleton.npz")
    anims = []
    for a in d["anims"]:
        k = np.array(a["keys"], np.float32).copy()
        ko = a["key_offset"]
        for t in range(len(ko) - 1):
            k[ko[t]:ko[t + 1], 4:8] = k[ko[t], 4:8]                  # one rotation per track
        rng = np.random.default_rng(3)
        k[:, 8:11] = (1.0 + 0.2 * rng.uniform(-1, 1, size=(len(k), 3))).astype(np.float32)     # animated scales
        anims.append(dict(a, keys=k, interpolation=1, rotation_interpolation=0))
    d = dict(d, anims=anims)
    sk = ogre_b200.Skeleton(ctx, d, use_given_bind_inverse=True)
    osk = oracle.Skeleton(d)
    rng = np.random.default_rng(9)
    ni = 300
    anim = rng.integers(0, len(anims), size=(ni, 1)).astype(np.int32)
    length = np.array([a["length"] for a in anims], np.float32)
    time = (rng.uniform(0, 1, size=(ni, 1)) * length[anim]).astype(np.float32)
    weight = np.ones((ni, 1), np.float32)
    got = sk.palettes(anim, time, weight)
    for i in range(ni):
        want = osk.palette(anim[i], time[i], weight[i], bind_inverse=d["bind_inverse"], blend_mode=0)
        assert np.array_equal(bits(got[i]), bits(want)), i
    sk.destroy()


@pytest.mark.parametrize("name", ["sinbad_skeleton.npz", "jaiqua_skeleton.npz"])
@pytest.mark.parametrize("shared_mask", [False, True])
def test_blend_masks_and_manual_bones_bit_exact(ctx, name, shared_mask):
    """ogrecu_skel_palettes_ex_dev: per-state blend masks (per instance, or one set shared by all) and manually controlled bones with
    per-instance local transforms — the oracle for these is pinned bit for bit against the real SkeletonInstance
    (tests/test_oracle_vs_ref.py::test_oracle_blend_masks_and_manual_bones_equal_real_skeleton_instance); the device equals it."""
    d, _ = load_skeleton(name)
    sk = ogre_b200.Skeleton(ctx, d, use_given_bind_inverse=True)
    osk = oracle.Skeleton(d)
    rng = np.random.default_rng(77)
    ni, ns, nb = 403, 3, d["num_bones"]
    na = len(d["anims"])
    anim = rng.integers(-1, na, size=(ni, ns)).astype(np.int32)
    for i in range(ni):
        seen = set()
        for s in range(ns):
            if anim[i, s] in seen:
                anim[i, s] = -1
            seen.add(int(anim[i, s]))
    length = np.array([a["length"] for a in d["anims"]] + [0.0], np.float32)
    time = (rng.uniform(0, 1.2, size=(ni, ns)) * length[anim]).astype(np.float32)
    weight = rng.uniform(0, 1, size=(ni, ns)).astype(np.float32)
    weight[rng.uniform(size=(ni, ns)) < 0.2] = 1.0
    mask = rng.uniform(0, 1, size=((ns, nb) if shared_mask else (ni, ns, nb))).astype(np.float32)
    mask[rng.uniform(size=mask.shape) < 0.2] = 0.0
    mask[rng.uniform(size=mask.shape) < 0.2] = 1.0
    k = 4
    handles = rng.choice(nb, k, replace=False).astype(np.uint32)
    bind = np.array(d["bind"], np.float32).reshape(-1, 10)
    manual = np.repeat(bind[handles][None], ni, axis=0).copy()
    manual[:, :, 0:3] += rng.uniform(-0.5, 0.5, (ni, k, 3)).astype(np.float32)
    q = manual[:, :, 3:7].astype(np.float64) + rng.uniform(-0.3, 0.3, (ni, k, 4))
    manual[:, :, 3:7] = (q / np.linalg.norm(q, axis=2, keepdims=True)).astype(np.float32)
    manual[:, :, 7:10] *= rng.uniform(0.8, 1.25, (ni, k, 3)).astype(np.float32)
    slot = np.full(nb, -1, np.int32); slot[handles] = np.arange(k, dtype=np.int32)
    for mode in (0, 1):
        sk.set_blend_mode(mode)
        got = sk.palettes(anim, time, weight, mask=mask, manual_bones=handles, manual=manual)
        only_mask = sk.palettes(anim, time, weight, mask=mask)
        for i in range(0, ni, 5):
            on = anim[i] >= 0
            # the oracle takes the masks of the ENABLED states in order (disabled slots are skipped on both sides)
            mi = (mask if shared_mask else mask[i])[on]
            want = osk.palette(anim[i][on], time[i][on], weight[i][on], bind_inverse=d["bind_inverse"], blend_mode=mode, mask=mi,
                               manual_slot=slot, manual=manual[i])
            assert np.array_equal(bits(got[i]), bits(want)), (name, mode, i)
            want = osk.palette(anim[i][on], time[i][on], weight[i][on], bind_inverse=d["bind_inverse"], blend_mode=mode, mask=mi)
            assert np.array_equal(bits(only_mask[i]), bits(want)), (name, mode, i, "mask only")
    sk.destroy()


def test_bad_modes_and_the_first_kernel_fail_loudly(ctx):
    d, _ = load_skeleton("jaiqua_skeleton.npz")
    sk = ogre_b200.Skeleton(ctx, dict(d, anims=[]))
    with pytest.raises(ogre_b200.OgrecuError):
        sk.add_animation(dict(d["anims"][0], interpolation=2))
    sk.add_animation(dict(d["anims"][0], interpolation=1))
    ctx.set_config_option("skel_mode", 1)              # the first-version kernel (A/B only) knows the linear modes only
    try:
        with pytest.raises(ogre_b200.OgrecuError):
            sk.palettes(np.zeros((1, 1), np.int32), np.zeros((1, 1), np.float32), np.ones((1, 1), np.float32))
    finally:
        ctx.set_config_option("skel_mode", 0)
    sk.destroy()
