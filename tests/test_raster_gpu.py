"""GPU parity: CUDA rasteriser (through the C ABI) vs the CPU restatement of Tiny.
Gate (north_star): coverage and depth bit-exact, colour <= 1 LSB.  We assert colour EXACT."""
import numpy as np
import pytest

from datagen import clip_case, degenerate_soup, make_scene
from rasterlib import assert_images_equal, render_gpu, render_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h", [(257, 193), (64, 64), (640, 480), (33, 17)])
def test_random_scenes_exact(gpu_ctx, w, h):
    for seed in range(3):
        cases = make_scene(seed * 7 + w, w, h, ndraws=6, ntri=150)
        rc, rz, frags = render_oracle(cases, w, h)
        gc, gz, stats = render_gpu(gpu_ctx, cases, w, h)
        assert frags > 0
        assert_images_equal(gc, gz, rc, rz, "scene seed=%d %dx%d" % (seed, w, h))


def test_corner_cases_exact(gpu_ctx):
    w, h = 128, 96
    rng = np.random.default_rng(2)
    clip = degenerate_soup(rng, w, h)
    for kw in (dict(lighting=False), dict(lighting=True, blend_add=True), dict(depth_write=False), dict(depth_test=False)):
        cases = clip_case(clip, w, h, **kw)
        rc, rz, _ = render_oracle(cases, w, h)
        gc, gz, _ = render_gpu(gpu_ctx, cases, w, h)
        assert_images_equal(gc, gz, rc, rz, "corner cases %r" % (kw,))


def test_tiny_triangles_dense(gpu_ctx):
    # many sub-pixel triangles: Tiny draws rounding-noise pixels inside the bbox; must match bit for bit
    w, h = 320, 240
    cases = make_scene(5, w, h, ndraws=3, ntri=20000, size=0.02)
    rc, rz, frags = render_oracle(cases, w, h)
    gc, gz, stats = render_gpu(gpu_ctx, cases, w, h)
    assert_images_equal(gc, gz, rc, rz, "tiny triangles")


def test_no_clear_keeps_previous_contents(gpu_ctx):
    w, h = 100, 80
    rng = np.random.default_rng(9)
    init_c = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    init_z = rng.uniform(0.3, 1.0, size=(h, w)).astype(np.float32)
    cases = make_scene(3, w, h, ndraws=4, ntri=60)
    rc, rz, _ = render_oracle(cases, w, h, clear=None, init=(init_c, init_z))
    gc, gz, _ = render_gpu(gpu_ctx, cases, w, h, clear=None, init=(init_c, init_z))
    assert_images_equal(gc, gz, rc, rz, "no clear")


def test_clear_values(gpu_ctx):
    import ogre_b200
    rt = ogre_b200.RenderTarget(gpu_ctx, 50, 30)
    for col in ((0, 0, 0, 0), (0.1, 0.2, 0.3, 1), (1.0, 0.999, 0.5, 0), (-1, 2, 0.00390625, 1)):
        rt.clear(colour=col, depth=0.75)
        rt.flush()
        rc, rz, _ = render_oracle([], 50, 30, clear=col, depth=0.75)
        assert np.array_equal(rt.read_colour(), rc), col
        assert np.array_equal(rt.read_depth(), rz)
    assert rt.read_colour((5, 3, 20, 11)).shape == (8, 15, 3)
    with pytest.raises(ogre_b200.OgrecuError):
        rt.read_colour((0, 0, 51, 30))
    rt.destroy()


def test_split_frame_strips_compose(gpu_ctx):
    # config 5 mechanics on one GPU: strips rendered separately compose to the full frame
    import ogre_b200
    from rasterlib import GpuScene
    w, h = 256, 160
    cases = make_scene(12, w, h, ndraws=5, ntri=300)
    rc, rz, _ = render_oracle(cases, w, h)
    scene = GpuScene(gpu_ctx, cases)
    full_c = np.zeros((h, w, 3), np.uint8); full_z = np.zeros((h, w), np.float32)
    for (y0, y1) in ((0, 48), (48, 112), (112, 160)):
        rt = ogre_b200.RenderTarget(gpu_ctx, w, h)
        rt.set_strip(y0, y1)
        scene.render(rt)
        full_c[y0:y1] = rt.read_colour()[y0:y1]; full_z[y0:y1] = rt.read_depth()[y0:y1]
        rt.destroy()
    scene.free()
    assert_images_equal(full_c, full_z, rc, rz, "split frame")
