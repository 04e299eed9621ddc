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


@pytest.mark.parametrize("w,h,band,world", [(256, 160, 16, 3), (250, 150, 32, 4), (64, 33, 16, 8)])
def test_split_frame_interleaved_bands_compose(gpu_ctx, w, h, band, world):
    # load-balanced split (ogrecu_rt_set_bands): rank r owns the bands b with b % world == r; the owned rows of all
    # ranks compose to the full frame, rows a rank does not own stay untouched (cleared value)
    import ogre_b200
    from ogre_b200.shard import band_row_indices
    from rasterlib import GpuScene
    cases = make_scene(21, w, h, ndraws=5, ntri=400)
    rc, rz, _ = render_oracle(cases, w, h)
    scene = GpuScene(gpu_ctx, cases)
    full_c = np.zeros((h, w, 3), np.uint8); full_z = np.zeros((h, w), np.float32)
    seen = np.zeros(h, int)
    for r in range(world):
        rt = ogre_b200.RenderTarget(gpu_ctx, w, h)
        rt.set_bands(band, r, world)
        scene.render(rt)
        rows = band_row_indices(h, r, world, band)
        seen[rows] += 1
        c, z = rt.read_colour(), rt.read_depth()
        full_c[rows] = c[rows]; full_z[rows] = z[rows]
        others = np.setdiff1d(np.arange(h), rows)
        assert (z[others] == 1.0).all(), "a rank wrote rows it does not own"
        rt.destroy()
    scene.free()
    assert (seen == 1).all()
    assert_images_equal(full_c, full_z, rc, rz, "interleaved split frame")


def test_draw_views_equals_one_draw_per_view(gpu_ctx):
    # config 4 mechanics: one pipeline pass over n_views x triangles with per-view mvp / mvit / light
    import ctypes as C
    import ogre_b200
    from ogre_b200 import capi, check, lib
    from datagen import DrawCase, look_at, perspective, random_soup, random_texture, viewport_matrix
    from rasterlib import GpuScene
    ctx = gpu_ctx
    w, h = 200, 152
    rng = np.random.default_rng(77)
    verts, idx = random_soup(rng, 3000, size=0.4)
    tex = random_texture(rng, 64, 32, zero_alpha=True)
    nviews = 5
    proj = perspective(50, w / h, 0.5, 100.0)
    mvps, mvits, lights, refs = [], [], [], []
    for k in range(nviews):
        view = look_at(rng.uniform(-2, 2, 3) + np.array([0, 0, 9.0]), rng.uniform(-1, 1, 3))
        mvps.append((proj @ view).astype(np.float32))
        mvits.append(np.ascontiguousarray(np.linalg.inv(view).T, dtype=np.float32))
        l = rng.normal(size=3); lights.append((l / np.linalg.norm(l)).astype(np.float32))
        c = DrawCase(verts, idx, index_type=4, texture=tex, mvp=mvps[k], mvit=mvits[k], viewport=viewport_matrix(w, h),
                     ambient=(0.2, 0.2, 0.25, 1.0))
        c.light_dir = lights[k]
        refs.append(render_oracle([c], w, h))
    mvps, mvits, lights = [np.ascontiguousarray(np.stack(a)) for a in (mvps, mvits, lights)]
    base = DrawCase(verts, idx, index_type=4, texture=tex, mvp=mvps[0], mvit=mvits[0], viewport=viewport_matrix(w, h),
                    ambient=(0.2, 0.2, 0.25, 1.0))
    scene = GpuScene(ctx, [base])
    rts = [ogre_b200.RenderTarget(ctx, w, h) for _ in range(nviews)]
    for r in rts:
        r.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
    hs = (C.c_void_p * nviews)(*[r.handle for r in rts])
    fp = capi.c_float_p
    check(lib().ogrecu_draw_views(hs, nviews, C.byref(scene.descs[0]), mvps.ctypes.data_as(fp), mvits.ctypes.data_as(fp),
                                  lights.ctypes.data_as(fp)))
    for k, r in enumerate(rts):
        assert_images_equal(r.read_colour(), r.read_depth(), refs[k][0], refs[k][1], "draw_views view %d" % k)
        r.destroy()
    scene.free()


@pytest.mark.parametrize("world", [2, 5])
def test_sort_last_composite_equals_whole_frame(gpu_ctx, world):
    """Sort-last split of one frame: slice r of the TRIANGLES (submission order) rendered into its own full-size target,
    targets merged per pixel with ogrecu_composite_depth_dev in rank order — equals the frame rendered at once, including
    pixels where fragments of different slices tie in depth (the later slice must win) and untouched pixels."""
    import ctypes as C
    import os
    import sys
    import ogre_b200
    from ogre_b200 import capi, check, lib
    from ogre_b200.shard import triangle_range
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import bench_configs as bc
    ctx = gpu_ctx
    w, h = 640, 480
    verts, idx = bc.terrain(200)
    # coplanar duplicates with other texture coordinates, appended at the end: they tie in depth with the original
    # triangles of the first slices and must win (later in submission order)
    tri = idx.reshape(-1, 3)
    dup = tri[1000:3000].copy()
    extra = verts[dup.reshape(-1)].copy()
    extra[:, 6:8] += 0.37
    base = len(verts)
    verts2 = np.concatenate([verts, extra]).astype(np.float32)
    idx2 = np.concatenate([idx, base + np.arange(len(extra), dtype=np.uint32)]).astype(np.uint32)
    ntri = len(idx2) // 3
    desc, keep = bc.upload_scene(ctx, verts2, idx2, bc.checker_texture(256), w, h)
    mvps, mvits, lights = bc.camera_ring(1, w, h)
    clear = (0.1, 0.2, 0.3, 1.0)

    def render(rt, lo, hi):
        d = ogre_b200.DrawDesc.from_buffer_copy(desc)
        d.indices = desc.indices + lo * 12
        d.draw_count = (hi - lo) * 3
        rt.clear(colour=clear, depth=1.0)
        hs = (C.c_void_p * 1)(rt.handle)
        check(lib().ogrecu_draw_views(hs, 1, C.byref(d), mvps.ctypes.data_as(capi.c_float_p), mvits.ctypes.data_as(capi.c_float_p),
                                      lights.ctypes.data_as(capi.c_float_p)))
        rt.flush()
    whole = ogre_b200.RenderTarget(ctx, w, h)
    render(whole, 0, ntri)
    wc, wz = whole.read_colour(), whole.read_depth()
    assert whole.stats()["tiles_ordered"] == 0
    parts = []
    for r in range(world):
        rt = ogre_b200.RenderTarget(ctx, w, h)
        render(rt, *triangle_range(ntri, r, world))
        parts.append(rt)
    ptr = lambda rt, f: (lambda p: (check(f(rt.handle, C.byref(p))), p.value)[1])(C.c_void_p())
    c0, z0 = ptr(parts[0], lib().ogrecu_rt_colour_device_ptr), ptr(parts[0], lib().ogrecu_rt_depth_device_ptr)
    cl = (C.c_float * 4)(*clear)
    if world == 2:
        check(lib().ogrecu_composite_depth_dev(ctx.handle, c0, z0, ptr(parts[1], lib().ogrecu_rt_colour_device_ptr),
                                               ptr(parts[1], lib().ogrecu_rt_depth_device_ptr), w * h, 1, 0, cl, 1.0))
    else:
        # several sources in ONE call: stacked planes of ranks 1 .. G-1, in rank order
        sc = ctx.alloc((world - 1) * w * h * 3); sz = ctx.alloc((world - 1) * w * h * 4)
        for r in range(1, world):
            check(lib().ogrecu_memcpy_d2d(ctx.handle, sc.ptr + (r - 1) * w * h * 3, ptr(parts[r], lib().ogrecu_rt_colour_device_ptr), w * h * 3))
            check(lib().ogrecu_memcpy_d2d(ctx.handle, sz.ptr + (r - 1) * w * h * 4, ptr(parts[r], lib().ogrecu_rt_depth_device_ptr), w * h * 4))
        check(lib().ogrecu_composite_depth_dev(ctx.handle, c0, z0, sc.ptr, sz.ptr, w * h, world - 1, w * h, cl, 1.0))
    ctx.sync()
    gc, gz = parts[0].read_colour(), parts[0].read_depth()
    assert np.array_equal(gz.view(np.uint32), wz.view(np.uint32))
    assert np.array_equal(gc, wc)
    # the duplicates really decided some pixels (otherwise the tie rule was not exercised)
    only_first = ogre_b200.RenderTarget(ctx, w, h)
    render(only_first, 0, len(idx) // 3)
    assert (only_first.read_colour() != wc).any()
    for rt in parts + [whole, only_first]:
        rt.destroy()


def test_readback_packed32_equals_host_shuffle(gpu_ctx):
    # ogrecu_rt_read_colour_packed32: RGB8 plane -> r << rs | g << gs | b << bs | fill per pixel, whole target and sub-box
    import ogre_b200
    from rasterlib import GpuScene
    w, h = 250, 130
    cases = make_scene(31, w, h, ndraws=4, ntri=200)
    scene = GpuScene(gpu_ctx, cases)
    rt = ogre_b200.RenderTarget(gpu_ctx, w, h)
    scene.render(rt, (0.1, 0.2, 0.3, 1.0), 1.0)
    rgb = rt.read_colour().astype(np.uint32)
    for rs, gs, bs, fill in [(16, 8, 0, 0xFF000000), (0, 8, 16, 0xFF000000), (24, 16, 8, 0xFF), (8, 16, 24, 0xFF)]:
        want = (rgb[..., 0] << rs) | (rgb[..., 1] << gs) | (rgb[..., 2] << bs) | fill
        assert np.array_equal(rt.read_colour_packed32(rs, gs, bs, fill), want)
        assert np.array_equal(rt.read_colour_packed32(rs, gs, bs, fill, box=(7, 3, 201, 99)), want[3:99, 7:201])
    pinned = ogre_b200.pinned_empty((h, w), np.uint32)
    assert np.array_equal(rt.read_colour_packed32(16, 8, 0, 0xFF000000, out=pinned), (rgb[..., 0] << 16) | (rgb[..., 1] << 8) | rgb[..., 2] | 0xFF000000)
    scene.free(); rt.destroy()


def test_scaled_readback_equals_image_scale(gpu_ctx):
    """ogrecu_rt_read_colour_scaled: the blitToMemory read-back into a destination of another size,
    resampled on the device, vs the restated Image::scale(FILTER_BILINEAR) and vs the real one's outputs in
    tests/golden/image_scale.npz — RGB8 with padded rows, and packed 32-bit destinations."""
    import os
    import ogre_b200
    import oracle
    from datagen import PACKED32, SCALE_CASES, pack32, scale_image
    from goldenlib import GOLDEN
    z = np.load(os.path.join(GOLDEN, "image_scale.npz"))
    for k, (seed, h, w, box, dw, dh) in enumerate(SCALE_CASES):
        src = scale_image(seed, h, w)
        rt = ogre_b200.RenderTarget(gpu_ctx, w, h)
        rt.write_colour(src)
        got = rt.read_colour_scaled(dw, dh, box=box, pad=k % 3)
        assert np.array_equal(got, oracle.scale_bilinear(src, dw, dh, box=box, dst_pad=k % 3)), k
        assert np.array_equal(got, z["c%d_rgb" % k]), k
        for name in ("PF_A8R8G8B8", "PF_A8B8G8R8"):
            words = rt.read_colour_scaled(dw, dh, box=box, packed=PACKED32[name], pad=2)
            assert np.array_equal(words[:, :dw], pack32(oracle.scale_bilinear(src, dw, dh, box=box), *PACKED32[name])), (k, name)
            assert (words[:, dw:] == 0xEEEEEEEE).all()
        assert np.array_equal(rt.read_colour_scaled(dw, dh, box=box, packed=PACKED32["PF_A8R8G8B8"]).view(np.uint8).reshape(dh, dw, 4),
                              z["c%d_PF_A8R8G8B8" % k]), k
        rt.destroy()
    # a rendered frame, full size: 1280x720 -> 640x360 and -> 1920x1080
    from rasterlib import GpuScene
    w, h = 1280, 720
    scene = GpuScene(gpu_ctx, make_scene(77, w, h, ndraws=4, ntri=300))
    rt = ogre_b200.RenderTarget(gpu_ctx, w, h)
    scene.render(rt, (0.1, 0.2, 0.3, 1.0), 1.0)
    rgb = rt.read_colour()
    for dw, dh in ((640, 360), (1920, 1080), (333, 719)):
        assert np.array_equal(rt.read_colour_scaled(dw, dh), oracle.scale_bilinear(rgb, dw, dh)), (dw, dh)
    with pytest.raises(ogre_b200.OgrecuError):
        rt.read_colour_scaled(10, 10, box=(0, 0, w + 1, h))
    scene.free(); rt.destroy()


@pytest.mark.parametrize("index_type,blend", [(4, False), (2, False), (0, False), (4, True)], ids=["u32", "u16", "nonindexed", "u32_blend_add"])
def test_instanced_draw_equals_expanded_indices(gpu_ctx, index_type, blend):
    """ogrecu_draw_desc.instance_count: instance i draws the same indices shifted by i * instance_vertex_count.  Equal, bit for
    bit (colour and depth; with additive blending the order matters too), to ONE draw over the concatenated vertex stream with
    the index list written out per instance — on the GPU and in the oracle (= Tiny, which knows no instancing)."""
    import ctypes as C
    import ogre_b200
    from datagen import DrawCase, look_at, perspective, random_soup, random_texture, viewport_matrix
    from rasterlib import GpuScene
    ctx = gpu_ctx
    w, h = 200, 152
    rng = np.random.default_rng(40 + index_type)
    ninst, ntri = 37, 60
    verts, idx = random_soup(rng, ntri, size=0.7, strip=False, indexed=index_type != 0)
    nv = len(verts)
    allv = np.concatenate([verts + np.array([rng.uniform(-3, 3), rng.uniform(-2, 2), rng.uniform(-2, 2), 0, 0, 0, 0, 0], np.float32) for _ in range(ninst)])
    if index_type:
        expanded = np.concatenate([idx.astype(np.uint32) + i * nv for i in range(ninst)])
        per_draw = len(idx)
    else:
        assert nv % 3 == 0
        expanded, per_draw = None, nv
    proj = perspective(50, w / h, 0.5, 100.0)
    mv = look_at(np.array([0.5, 0.3, 9.0]), (0, 0, 0))
    kw = dict(texture=random_texture(rng, 37, 19), lighting=True, blend_add=blend, depth_write=not blend, mvp=(proj @ mv).astype(np.float32),
              mvit=np.linalg.inv(mv).T.astype(np.float32), viewport=viewport_matrix(w, h), light_dir=(0.2, 0.4, 0.9))
    whole = DrawCase(allv, expanded, index_type=4, **kw)
    rc, rz, frags = render_oracle([whole], w, h)
    gc, gz, _ = render_gpu(ctx, [whole], w, h)
    assert frags > 0
    assert_images_equal(gc, gz, rc, rz, "expanded indices")
    # the instanced form: the base index list once, instance_count copies
    base = DrawCase(allv, idx if index_type else None, index_type=index_type if index_type else 4, **kw)
    scene = GpuScene(ctx, [base])
    d = scene.descs[0]
    d.draw_count, d.vertex_count = per_draw, ninst * nv
    d.instance_count, d.instance_vertex_count = ninst, nv
    rt = ogre_b200.RenderTarget(ctx, w, h)
    scene.render(rt)
    ic, iz = rt.read_colour(), rt.read_depth()
    assert_images_equal(ic, iz, rc, rz, "instanced draw")
    # and without the vertex pre-transform (vertex_count unknown)
    d.vertex_count = 0
    scene.render(rt)
    assert_images_equal(rt.read_colour(), rt.read_depth(), rc, rz, "instanced draw, per-index transform")
    scene.free(); rt.destroy()


def test_draw_views_clear_batching_keeps_per_target_clears(gpu_ctx):
    """ogrecu_draw_views runs the targets' pending clears as one launch when they all hold the same single clear; targets cleared to
    DIFFERENT values, a target with a depth-only clear, and a target with an earlier draw recorded take the per-target path — every
    view still equals the oracle on its own initial contents"""
    import ctypes as C
    import ogre_b200
    from ogre_b200 import capi, check, lib
    from datagen import DrawCase, look_at, perspective, random_soup, viewport_matrix
    from rasterlib import GpuScene
    ctx = gpu_ctx
    w, h = 160, 96
    rng = np.random.default_rng(5)
    verts, idx = random_soup(rng, 800, size=0.5)
    proj = perspective(50, w / h, 0.5, 100.0)
    nviews = 5
    mvps, mvits, lights = [], [], []
    for k in range(nviews):
        view = look_at(rng.uniform(-2, 2, 3) + np.array([0, 0, 9.0]), rng.uniform(-1, 1, 3))
        mvps.append((proj @ view).astype(np.float32)); mvits.append(np.ascontiguousarray(np.linalg.inv(view).T, dtype=np.float32))
        l = rng.normal(size=3); lights.append((l / np.linalg.norm(l)).astype(np.float32))
    cases = []
    for k in range(nviews):
        c = DrawCase(verts, idx, index_type=4, mvp=mvps[k], mvit=mvits[k], viewport=viewport_matrix(w, h), ambient=(0.2, 0.2, 0.25, 1.0), has_uv=False)
        c.light_dir = lights[k]
        cases.append(c)
    mv, mi, li = [np.ascontiguousarray(np.stack(a)) for a in (mvps, mvits, lights)]
    scene = GpuScene(ctx, [cases[0]])
    fp = capi.c_float_p
    for variant in ("same", "different colours", "one depth-only", "one with a draw before"):
        rts = [ogre_b200.RenderTarget(ctx, w, h) for _ in range(nviews)]
        refs = []
        for k, r in enumerate(rts):
            colour = (0.1, 0.2, 0.3, 1.0) if variant != "different colours" else (0.1 * k, 0.2, 0.3, 1.0)
            if variant == "one depth-only" and k == 2:
                r.clear(colour=(0.5, 0.5, 0.5, 1.0), depth=1.0); r.flush()
                r.clear(buffers=ogre_b200.CLEAR_DEPTH, depth=0.97)
                refs.append(render_oracle([cases[k]], w, h, clear=None, init=(np.full((h, w, 3), 128, np.uint8), np.full((h, w), 0.97, np.float32))))
            elif variant == "one with a draw before" and k == 1:
                r.clear(colour=colour, depth=1.0)
                r.draw(scene.descs[0])
                refs.append(render_oracle([cases[0], cases[k]], w, h, clear=colour))
            else:
                r.clear(colour=colour, depth=1.0)
                refs.append(render_oracle([cases[k]], w, h, clear=colour))
        hs = (C.c_void_p * nviews)(*[r.handle for r in rts])
        check(lib().ogrecu_draw_views(hs, nviews, C.byref(scene.descs[0]), mv.ctypes.data_as(fp), mi.ctypes.data_as(fp), li.ctypes.data_as(fp)))
        for k, r in enumerate(rts):
            assert_images_equal(r.read_colour(), r.read_depth(), refs[k][0], refs[k][1], "%s, view %d" % (variant, k))
            r.destroy()
    scene.free()


def test_named_buffers_are_tracked_like_looked_up_ones(gpu_ctx):
    """ogrecu_draw_refs: the caller names the ogrecu_bufs a draw reads; a host write to one of them before the flush must still run the
    recorded draw first (the same hazard rule as for pointers the library looks up itself)"""
    import ctypes as C
    import ogre_b200
    from ogre_b200 import check, lib
    ctx = gpu_ctx
    w, h = 128, 96
    cases = make_scene(8, w, h, ndraws=1, ntri=300)
    case = cases[0]
    case.texture = None
    rc, rz, _ = render_oracle([case], w, h)
    vb = ogre_b200.HardwareBuffer(ctx, case.verts.nbytes); vb.write(case.verts)
    ib = ogre_b200.HardwareBuffer(ctx, case.indices.nbytes); ib.write(case.indices)
    d = ogre_b200.DrawDesc()
    d.positions, d.pos_stride = vb.device_ptr, 32
    d.normals, d.normal_stride = (vb.device_ptr + 12, 32) if case.has_normals else (None, 0)
    d.indices, d.index_type, d.draw_count, d.vertex_count = ib.device_ptr, case.index_type, case.draw_count, len(case.verts)
    d.prim_type = ogre_b200.PRIM_TRIANGLE_LIST
    d.flags = ((ogre_b200.DRAW_DEPTH_TEST if case.depth_test else 0) | (ogre_b200.DRAW_DEPTH_WRITE if case.depth_write else 0) |
               (ogre_b200.DRAW_BLEND_ADD if case.blend_add else 0) | (ogre_b200.DRAW_LIGHTING if case.lighting else 0))
    d.mvp[:] = case.mvp.reshape(16); d.tex_matrix[:] = case.tex_matrix.reshape(16); d.mvit[:] = case.mvit.reshape(16)
    d.viewport[:] = case.viewport.reshape(16); d.light_dir[:] = case.light_dir; d.ambient[:] = case.ambient
    rt = ogre_b200.RenderTarget(ctx, w, h)
    rt.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
    refs = (C.c_void_p * 2)(vb.handle, ib.handle)
    check(lib().ogrecu_draw_refs(rt.handle, C.byref(d), None, refs, 2))
    vb.write(np.zeros_like(case.verts))                  # recorded, not flushed: the write has to wait for the draw
    assert_images_equal(rt.read_colour(), rt.read_depth(), rc, rz, "draw recorded with named buffers, vertices overwritten afterwards")
    rt.destroy(); vb.destroy(); ib.destroy()
