"""GPU parity of the opt-in "beyond Tiny" render state (ogrecu_draw_with_state) against its definition in
the oracle (oracle_tiny_draw_state; tied back to the pinned Tiny restatement by tests/test_beyond_tiny_cpu.py): coverage,
depth and colour bit-exact, on both raster paths."""
import numpy as np
import pytest

import ogre_b200
from datagen import make_scene, with_states
from rasterlib import GpuScene, assert_images_equal, render_gpu, render_oracle

pytestmark = pytest.mark.gpu

FIELDS = ("depth_func", "cull_mode", "scissor", "alpha", "colours")


@pytest.mark.parametrize("w,h", [(256, 192), (257, 193), (33, 17)])
def test_random_states_exact(gpu_ctx, w, h):
    for seed in range(4):
        mk = lambda: with_states(make_scene(seed * 11 + w, w, h, ndraws=6, ntri=150), seed + 40, w, h)
        rc, rz, frags = render_oracle(mk(), w, h)
        for mode in (0, 1):
            gpu_ctx.set_config_option("raster_mode", mode)
            try:
                gc, gz, _ = render_gpu(gpu_ctx, mk(), w, h)
            finally:
                gpu_ctx.set_config_option("raster_mode", 0)
            assert_images_equal(gc, gz, rc, rz, "seed %d %dx%d raster_mode %d" % (seed, w, h, mode))


@pytest.mark.parametrize("field", FIELDS)
def test_each_field_alone_exact(gpu_ctx, field):
    w, h = 320, 240
    total = 0
    for seed in range(3):
        kw = dict(depth_test=True) if field == "depth_func" else {}
        mk = lambda: with_states(make_scene(seed + 70, w, h, ndraws=5, ntri=250, **kw), seed, w, h, what=[field])
        rc, rz, frags = render_oracle(mk(), w, h)
        gc, gz, _ = render_gpu(gpu_ctx, mk(), w, h)
        assert_images_equal(gc, gz, rc, rz, "%s seed %d" % (field, seed))
        total += frags
    assert total > 0


def test_every_depth_function_and_cull_mode(gpu_ctx):
    w, h = 200, 120
    for func in range(8):
        for cull in (1, 2, 3):
            def mk():
                cases = make_scene(func * 3 + cull, w, h, ndraws=4, ntri=120, depth_test=True)
                for c in cases:
                    c.state = dict(depth_func=func, cull_mode=cull)
                return cases
            # two passes over the same target so that every function meets depths other than the clear value
            init = render_oracle(make_scene(99, w, h, ndraws=3, ntri=100, blend_add=False), w, h)[:2]
            rc, rz, _ = render_oracle(mk(), w, h, clear=None, init=init)
            gc, gz, _ = render_gpu(gpu_ctx, mk(), w, h, clear=None, init=init)
            assert_images_equal(gc, gz, rc, rz, "depth_func %d cull_mode %d" % (func, cull))


def test_state_off_is_tiny_and_plain_draws_are_unaffected_by_neighbours(gpu_ctx):
    # enable == 0 through ogrecu_draw_with_state = ogrecu_draw; and a flush mixing plain and stateful draws keeps the plain ones Tiny's
    w, h = 256, 160
    plain = render_oracle(make_scene(5, w, h), w, h)
    cases = make_scene(5, w, h)
    for c in cases:
        c.state = {"scissor": (0, 0, w, h)}       # a state that changes nothing
    gc, gz, _ = render_gpu(gpu_ctx, cases, w, h)
    assert_images_equal(gc, gz, plain[0], plain[1], "no-op state")
    scene = GpuScene(gpu_ctx, make_scene(5, w, h))
    rt = ogre_b200.RenderTarget(gpu_ctx, w, h)
    rt.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
    for d in scene.descs:
        rt.draw(d, ogre_b200.DrawState())          # enable = 0
    rt.flush()
    assert_images_equal(rt.read_colour(), rt.read_depth(), plain[0], plain[1], "enable == 0")
    scene.free(); rt.destroy()
    def mk():
        cases = with_states(make_scene(8, w, h, ndraws=6), 3, w, h)
        for c in cases[::2]:
            c.state = None
        return cases
    rc, rz, _ = render_oracle(mk(), w, h)
    gc, gz, _ = render_gpu(gpu_ctx, mk(), w, h)
    assert_images_equal(gc, gz, rc, rz, "mixed plain / stateful draws")


def test_bad_state_is_rejected(gpu_ctx):
    scene = GpuScene(gpu_ctx, make_scene(1, 64, 64, ndraws=1, ntri=10))
    rt = ogre_b200.RenderTarget(gpu_ctx, 64, 64)
    for st in (ogre_b200.DrawState(enable=1, depth_func=8), ogre_b200.DrawState(enable=2, cull_mode=0), ogre_b200.DrawState(enable=8, alpha_func=3, alpha_ref=256),
               ogre_b200.DrawState(enable=16), ogre_b200.DrawState(enable=32)):
        with pytest.raises(ogre_b200.OgrecuError):
            rt.draw(scene.descs[0], st)
    scene.free(); rt.destroy()


def test_full_hd_scissored_alpha_tested_frame(gpu_ctx):
    # a larger frame: 1920x1080, 40k triangles, every draw with all five fields
    w, h = 1920, 1080
    mk = lambda: with_states(make_scene(21, w, h, ndraws=4, ntri=10000, size=0.3), 6, w, h, what=list(FIELDS))
    rc, rz, frags = render_oracle(mk(), w, h)
    gc, gz, _ = render_gpu(gpu_ctx, mk(), w, h)
    assert_images_equal(gc, gz, rc, rz, "full HD")
