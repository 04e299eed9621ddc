"""CPU: the DEFINITION of the opt-in "beyond Tiny" render state (oracle_draw_state) checked through
properties that tie it back to the pinned Tiny restatement — the reference ignores this state, so there is nothing
of its own to pin against: with every extension off, or set to the value Tiny hard-wires, the result must be Tiny's image."""
import numpy as np

from datagen import make_scene, with_states
from rasterlib import assert_images_equal, render_oracle

W, H = 96, 64
LESS_EQUAL, ALWAYS_PASS, ALWAYS_FAIL, GREATER_EQUAL = 3, 1, 0, 6
CULL_NONE, CULL_CW, CULL_ACW = 1, 2, 3


def scene(seed=3, **kw):
    return make_scene(seed, W, H, ndraws=5, ntri=120, **kw)


def states(cases, **st):
    for c in cases:
        c.state = dict(st)
    return cases


def test_states_that_restate_tinys_hard_wired_rules_change_nothing():
    plain = render_oracle(scene(), W, H)
    # lists: Tiny's cull rule is CULL_CLOCKWISE; its depth rule CMPF_LESS_EQUAL; its alpha discard `alpha >= 1`
    for st in (dict(depth_func=LESS_EQUAL), dict(alpha=(ALWAYS_PASS, 77)), dict(alpha=(GREATER_EQUAL, 1)), dict(scissor=(0, 0, W, H)),
               dict(scissor=(0, 0, 40000, 40000)), dict(depth_func=LESS_EQUAL, alpha=(GREATER_EQUAL, 0), scissor=(0, 0, W, H))):
        got = render_oracle(states(scene(), **st), W, H)
        assert_images_equal(got[0], got[1], plain[0], plain[1], repr(st))
        assert got[2] == plain[2]
    lists = render_oracle(scene(strip=False), W, H)
    got = render_oracle(states(scene(strip=False), cull_mode=CULL_CW), W, H)
    assert_images_equal(got[0], got[1], lists[0], lists[1], "CULL_CLOCKWISE on lists")
    strips = render_oracle(scene(strip=True), W, H)
    got = render_oracle(states(scene(strip=True), cull_mode=CULL_NONE), W, H)
    assert_images_equal(got[0], got[1], strips[0], strips[1], "CULL_NONE on strips")


def test_scissor_confines_the_draws_to_the_rectangle():
    # without blending every pixel is independent: inside the rectangle the frame is Tiny's, outside it is the cleared target
    kw = dict(blend_add=False)
    plain_c, plain_z, _ = render_oracle(scene(**kw), W, H)
    clear_c, clear_z, _ = render_oracle([], W, H)
    for rect in ((10, 7, 70, 50), (0, 0, 1, 1), (95, 63, 96, 64), (30, 30, 30, 60), (50, 10, 20, 40), (0, 20, 96, 21)):
        c, z, _ = render_oracle(states(scene(**kw), scissor=rect), W, H)
        l, t, r, b = rect
        inside = np.zeros((H, W), bool)
        if r > l and b > t:
            inside[t:b, l:r] = True
        assert np.array_equal(c[inside], plain_c[inside]) and np.array_equal(z[inside].view(np.uint32), plain_z[inside].view(np.uint32)), rect
        assert np.array_equal(c[~inside], clear_c[~inside]) and np.array_equal(z[~inside], clear_z[~inside]), rect


def test_cull_modes_partition_the_triangles():
    # depth test off, no blending, no discard: the fragment counts of the two one-sided modes add up to CULL_NONE's
    kw = dict(depth_test=False, blend_add=False, texture=None, strip=False)
    n = {m: render_oracle(states(scene(**kw), cull_mode=m), W, H)[2] for m in (CULL_NONE, CULL_CW, CULL_ACW)}
    assert n[CULL_CW] > 0 and n[CULL_ACW] > 0 and n[CULL_CW] + n[CULL_ACW] == n[CULL_NONE]
    # strips: with a mode set, odd triangles count with the opposite sign, so both modes draw something of a random strip
    kw["strip"] = True
    s = {m: render_oracle(states(scene(**kw), cull_mode=m), W, H)[2] for m in (CULL_NONE, CULL_CW, CULL_ACW)}
    assert s[CULL_CW] > 0 and s[CULL_ACW] > 0 and s[CULL_CW] + s[CULL_ACW] == s[CULL_NONE]


def test_depth_functions():
    kw = dict(blend_add=False)
    off = render_oracle(scene(depth_test=False, **kw), W, H)
    got = render_oracle(states(scene(depth_test=True, **kw), depth_func=ALWAYS_PASS), W, H)
    assert_images_equal(got[0], got[1], off[0], off[1], "CMPF_ALWAYS_PASS = depth test off")
    cleared = render_oracle([], W, H)
    kw["depth_test"] = True      # a depth function only acts while the depth test is on (as _setDepthBufferParams' depthTest)
    got = render_oracle(states(scene(**kw), depth_func=ALWAYS_FAIL), W, H)
    assert_images_equal(got[0], got[1], cleared[0], cleared[1], "CMPF_ALWAYS_FAIL draws nothing")
    assert got[2] == 0
    # CMPF_GREATER against a target cleared to depth 0 keeps the FARTHEST fragment: every written depth is >= CMPF_LESS's
    far = render_oracle(states(scene(depth_write=True, **kw), depth_func=7), W, H, depth=0.0)
    near = render_oracle(states(scene(depth_write=True, **kw), depth_func=2), W, H, depth=1.0)
    both = (far[1] > 0.0) & (near[1] < 1.0)
    assert both.sum() > 100 and (far[1][both] >= near[1][both]).all()
    # CMPF_EQUAL against the depth a first pass wrote re-touches exactly those pixels
    first = render_oracle(scene(lighting=False, depth_write=True, texture=None, **kw), W, H)
    again = render_oracle(states(scene(lighting=False, depth_write=True, texture=None, **kw), depth_func=4), W, H, clear=None,
                          init=(first[0], first[1]))
    assert np.array_equal(again[1].view(np.uint32), first[1].view(np.uint32)) and again[2] > 0


def test_alpha_rejection_and_vertex_colours():
    kw = dict(blend_add=False)
    cleared = render_oracle([], W, H)
    got = render_oracle(states(scene(**kw), alpha=(ALWAYS_FAIL, 0)), W, H)
    assert_images_equal(got[0], got[1], cleared[0], cleared[1], "alpha CMPF_ALWAYS_FAIL")
    # untextured draws have alpha 255: `alpha < 255` rejects all of their fragments, `alpha >= 255` none
    plain = render_oracle(scene(texture=None, **kw), W, H)
    none = render_oracle(states(scene(texture=None, **kw), alpha=(2, 255)), W, H)
    allf = render_oracle(states(scene(texture=None, **kw), alpha=(GREATER_EQUAL, 255)), W, H)
    assert none[2] == 0 and allf[2] == plain[2]
    assert_images_equal(allf[0], allf[1], plain[0], plain[1], "alpha >= 255 on untextured draws")
    # a higher threshold never adds fragments
    tex = np.random.default_rng(9).integers(0, 256, (32, 48, 4), dtype=np.uint8)
    counts = [render_oracle(states(scene(texture=tex, **kw), alpha=(GREATER_EQUAL, ref)), W, H)[2] for ref in (0, 64, 128, 200, 255)]
    assert counts == sorted(counts, reverse=True) and counts[0] > counts[-1]
    # black vertex colours, unlit: every covered pixel turns black, coverage and depth stay Tiny's
    cases = scene(lighting=False, **kw)
    ref = render_oracle(scene(lighting=False, **kw), W, H)
    for c in cases:
        c.state = dict(colours=np.zeros((len(c.verts), 4), np.uint8))
    got = render_oracle(cases, W, H)
    assert np.array_equal(got[1].view(np.uint32), ref[1].view(np.uint32)) and got[2] == ref[2]
    covered = ref[1] != cleared[1]
    assert covered.sum() > 100 and (got[0][covered] == 0).all()


def test_random_states_are_deterministic_and_differ_from_tiny():
    a = render_oracle(with_states(scene(), 5, W, H), W, H)
    b = render_oracle(with_states(scene(), 5, W, H), W, H)
    assert_images_equal(a[0], a[1], b[0], b[1], "repeat")
    plain = render_oracle(scene(), W, H)
    assert (a[0] != plain[0]).any()
