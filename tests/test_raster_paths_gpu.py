"""GPU parity of BOTH per-tile resolve paths of the rasteriser against the CPU oracle:
the order-independent visibility path (opaque, depth-tested, depth-writing tiles) and the in-order
path, plus every condition that must make a tile fall back (blend, depth-write off, discarding
textures, NaN depths from degenerate triangles, NaN / negative initial depth)."""
import numpy as np
import pytest

from datagen import clip_case, degenerate_soup, make_scene, random_texture
from rasterlib import assert_images_equal, render_gpu, render_oracle

pytestmark = pytest.mark.gpu

OPAQUE = dict(depth_test=True, depth_write=True, blend_add=False)


def both_modes(ctx, cases, w, h, what, expect_visibility=None, **kw):
    rc, rz, frags = render_oracle(cases, w, h, **kw)
    out = {}
    for mode in (0, 1):
        for xf in (1, 0):       # vertex pre-transform on / off: same bits either way
            ctx.set_config_option("raster_mode", mode)
            ctx.set_config_option("raster_xf", xf)
            try:
                gc, gz, stats = render_gpu(ctx, cases, w, h, **kw)
            finally:
                ctx.set_config_option("raster_mode", 0)
                ctx.set_config_option("raster_xf", 1)
            assert_images_equal(gc, gz, rc, rz, "%s (raster_mode %d, raster_xf %d)" % (what, mode, xf))
        out[mode] = stats
    assert out[1]["tiles_visibility"] == 0 and out[1]["tiles_ordered"] == out[1]["tiles_active"] > 0
    if expect_visibility is True:
        assert out[0]["tiles_visibility"] == out[0]["tiles_active"] > 0, out[0]
    elif expect_visibility is False:
        assert out[0]["tiles_visibility"] == 0, out[0]
    return out[0]


@pytest.mark.parametrize("w,h", [(257, 193), (64, 64), (640, 480), (33, 17)])
@pytest.mark.parametrize("size", [1.0, 0.05])
def test_opaque_scenes_take_the_visibility_path(gpu_ctx, w, h, size):
    rng = np.random.default_rng(w)
    tex = random_texture(rng, 37, 19)          # no zero alpha: never discards
    for seed in range(2):
        cases = make_scene(seed * 5 + h, w, h, ndraws=5, ntri=400 if size == 1.0 else 6000, size=size, texture=tex, **OPAQUE)
        both_modes(gpu_ctx, cases, w, h, "opaque scene %dx%d size %g" % (w, h, size), expect_visibility=True)


def test_equal_depth_later_triangle_wins(gpu_ctx):
    # coplanar, overlapping, identical depth everywhere: submission order decides every pixel
    w, h = 96, 64

    def ndc(x, y, z):
        return [x / w * 2 - 1, y / h * 2 - 1, z, 1.0]
    tris = []
    for k in range(40):
        x0 = 4 + (k * 7) % 50
        tris.append([ndc(x0, 3, .25), ndc(x0 + 40, 5, .25), ndc(x0 + 10, 60, .25)])
        tris.append([ndc(x0, 3, .25), ndc(x0 + 10, 60, .25), ndc(x0 + 40, 5, .25)])     # other winding: one of the two survives the cull
    cases = clip_case(np.array(tris), w, h, lighting=True, **OPAQUE)
    st = both_modes(gpu_ctx, cases, w, h, "equal depth", expect_visibility=True)
    assert st["tile_pairs"] >= 40


def test_mixed_state_tiles_fall_back(gpu_ctx):
    w, h = 200, 150
    cases = make_scene(4, w, h, ndraws=6, ntri=300)       # blend-add, depth-write-off, zero-alpha texture draws mixed in
    st = both_modes(gpu_ctx, cases, w, h, "mixed state")
    assert st["tiles_ordered"] > 0


def test_nan_depth_triangles_fall_back_exactly(gpu_ctx):
    w, h = 128, 96
    rng = np.random.default_rng(2)
    clip = degenerate_soup(rng, w, h)                      # slivers, zero-area, w<0: NaN / inf depths
    cases = clip_case(clip, w, h, lighting=False, **OPAQUE)
    st = both_modes(gpu_ctx, cases, w, h, "degenerate opaque")
    assert st["tiles_ordered"] > 0                         # at least one tile saw a NaN and replayed in order


def test_bad_initial_depth_falls_back(gpu_ctx):
    w, h = 100, 80
    rng = np.random.default_rng(9)
    init_c = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    init_z = rng.uniform(0.3, 1.0, size=(h, w)).astype(np.float32)
    init_z[10:30, 20:60] = np.nan
    init_z[50:70, 5:40] = -0.5
    init_z[40, :] = 0.0
    cases = make_scene(3, w, h, ndraws=4, ntri=200, texture=None, **OPAQUE)
    st = both_modes(gpu_ctx, cases, w, h, "no clear, NaN / negative initial depth", clear=None, init=(init_c, init_z))
    assert st["tiles_ordered"] > 0 and st["tiles_visibility"] > 0


def test_many_triangles_per_tile(gpu_ctx):
    # far more than one 256-entry chunk per tile, small and large triangles mixed
    w, h = 64, 48
    cases = make_scene(8, w, h, ndraws=2, ntri=20000, size=0.3, texture=None, **OPAQUE)
    both_modes(gpu_ctx, cases, w, h, "deep tile lists", expect_visibility=True)
