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


def test_unsupported_modes_fail_loudly(ctx):
    d, _ = load_skeleton("jaiqua_skeleton.npz")
    sk = ogre_b200.Skeleton(ctx, dict(d, anims=[]))
    a = dict(d["anims"][0], interpolation=1)
    with pytest.raises(ogre_b200.OgrecuError) as e:
        sk.add_animation(a)
    assert e.value.status == 5
    sk.destroy()
