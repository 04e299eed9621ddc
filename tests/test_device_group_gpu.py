"""GPU (two or more devices): the C ABI's device group (ogrecu_init_devices, include/ogrecu.h) — one process, one context, several
B200s.  Everything a group renders must equal what one device renders, which the other GPU tests tie to the oracle:
 * SPLIT targets (BASELINE config 5): the frame's tile rows dealt out in interleaved bands, every device rasterising its bands
   straight into the root's planes over NVLink; in-order content, clears between draws, textures, vertex colours included;
 * PLACED targets (config 4): batched views (ogrecu_draw_views) and ordinary draws on targets that live on the peers;
 * replicas: geometry in HardwareBuffers is copied to the peers when referenced and re-copied (only the changed range) after a
   host write or an OptimisedUtil result; raw device pointers are read in place through peer access.
Skipped on a box with one GPU (run with `gpurun --gpus 2`; the log is kept under profiles/)."""
import ctypes as C

import numpy as np
import pytest

import ogre_b200
from datagen import DrawCase, SkinCase, look_at, make_scene, perspective, random_soup, random_texture, viewport_matrix, with_states
from rasterlib import GpuScene, assert_images_equal, render_oracle

pytestmark = pytest.mark.gpu


def device_count():
    n = C.c_int(0)
    ogre_b200.lib().ogrecu_device_count(C.byref(n))
    return n.value


@pytest.fixture(scope="module")
def group():
    n = device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    ctx = ogre_b200.Context(devices=list(range(min(n, 4))))
    yield ctx
    ctx.close()


def test_group_reports_its_devices_and_rejects_bad_lists(group):
    n = C.c_int(0)
    assert ogre_b200.lib().ogrecu_num_devices(group.handle, C.byref(n)) == 0 and n.value == group.num_devices
    h = C.c_void_p()
    twice = (C.c_int * 2)(0, 0)
    assert ogre_b200.lib().ogrecu_init_devices(2, twice, C.byref(h)) != 0 and not h.value
    far = (C.c_int * 2)(0, 99)
    assert ogre_b200.lib().ogrecu_init_devices(2, far, C.byref(h)) != 0 and not h.value
    rt = C.c_void_p()
    assert ogre_b200.lib().ogrecu_rt_create_placed(group.handle, 64, 64, group.num_devices, C.byref(rt)) != 0


@pytest.mark.parametrize("w,h,band_tiles,seed", [(256, 160, 1, 3), (250, 150, 2, 4), (640, 360, 4, 5), (64, 33, 1, 6)])
def test_split_target_equals_the_oracle(group, w, h, band_tiles, seed):
    """a scene with opaque, blended, depth-write-off and textured draws (both raster paths), three segments with clears between"""
    group.set_config_option("split_band_tiles", band_tiles)
    cases = make_scene(seed, w, h, ndraws=6, ntri=500)
    rc, rz, _ = render_oracle(cases, w, h)
    rt = ogre_b200.RenderTarget(group, w, h, placement=ogre_b200.PLACE_SPLIT)
    scene = GpuScene(group, cases)
    scene.render(rt)
    assert_images_equal(rt.read_colour(), rt.read_depth(), rc, rz, "split target")
    # a second frame into the same target, with a depth-only clear in the middle of the list (ordered against the peers' passes)
    rt.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
    for d, st in list(zip(scene.descs, scene.states))[:3]:
        rt.draw(d, st)
    rt.clear(buffers=ogre_b200.CLEAR_DEPTH, depth=1.0)
    for d, st in list(zip(scene.descs, scene.states))[3:]:
        rt.draw(d, st)
    rt.flush()
    oc = np.zeros((h, w, 3), np.uint8); oz = np.zeros((h, w), np.float32)
    import oracle
    ol = oracle.lib()
    col = np.asarray((0.1, 0.2, 0.3, 1.0), np.float32)
    ol.oracle_tiny_clear(oc.ctypes.data, oz.ctypes.data, w, h, 3, col.ctypes.data, 1.0)
    for c in cases[:3]:
        d = c.oracle_desc(); ol.oracle_tiny_draw(C.byref(d), oc.ctypes.data, oz.ctypes.data, w, h)
    ol.oracle_tiny_clear(oc.ctypes.data, oz.ctypes.data, w, h, 2, col.ctypes.data, 1.0)
    for c in cases[3:]:
        d = c.oracle_desc(); ol.oracle_tiny_draw(C.byref(d), oc.ctypes.data, oz.ctypes.data, w, h)
    assert_images_equal(rt.read_colour(), rt.read_depth(), oc, oz, "split target, clear between segments")
    scene.free(); rt.destroy()
    group.set_config_option("split_band_tiles", 4)


def test_split_target_with_the_state_tiny_ignores(group):
    w, h = 320, 200
    cases = with_states(make_scene(9, w, h, ndraws=5, ntri=400), 31, w, h)
    rc, rz, _ = render_oracle(cases, w, h)
    rt = ogre_b200.RenderTarget(group, w, h, placement=ogre_b200.PLACE_SPLIT)
    scene = GpuScene(group, cases)
    scene.render(rt)
    assert_images_equal(rt.read_colour(), rt.read_depth(), rc, rz, "split target, beyond-Tiny state")
    scene.free(); rt.destroy()


class BufferScene:
    """DrawCases whose vertex / index data live in HardwareBuffers (ogrecu_buf): the peers read replicas"""

    def __init__(self, ctx, cases):
        self.bufs, self.descs = [], []
        for case in cases:
            vb = ogre_b200.HardwareBuffer(ctx, case.verts.nbytes); vb.write(case.verts)
            ib = None
            if case.indices is not None:
                ib = ogre_b200.HardwareBuffer(ctx, case.indices.nbytes); ib.write(case.indices)
            self.bufs.append((vb, ib))
            d = ogre_b200.DrawDesc()
            d.positions, d.pos_stride = vb.device_ptr, 32
            d.normals, d.normal_stride = (vb.device_ptr + 12, 32) if case.has_normals else (None, 0)
            d.uvs, d.uv_stride = (vb.device_ptr + 24, 32) if case.has_uv else (None, 0)
            d.indices = ib.device_ptr if ib is not None else None
            d.index_type, d.draw_count, d.vertex_count = case.index_type, case.draw_count, len(case.verts)
            d.prim_type = ogre_b200.PRIM_TRIANGLE_STRIP if case.strip else ogre_b200.PRIM_TRIANGLE_LIST
            d.flags = ((ogre_b200.DRAW_DEPTH_TEST if case.depth_test else 0) | (ogre_b200.DRAW_DEPTH_WRITE if case.depth_write else 0) |
                       (ogre_b200.DRAW_BLEND_ADD if case.blend_add else 0) | (ogre_b200.DRAW_LIGHTING if case.lighting else 0))
            d.mvp[:] = case.mvp.reshape(16); d.tex_matrix[:] = case.tex_matrix.reshape(16)
            d.mvit[:] = case.mvit.reshape(16); d.viewport[:] = case.viewport.reshape(16)
            d.light_dir[:] = case.light_dir; d.ambient[:] = case.ambient
            self.descs.append(d)

    def render(self, rt):
        rt.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
        for d in self.descs:
            rt.draw(d)
        rt.flush()

    def free(self):
        for vb, ib in self.bufs:
            vb.destroy()
            if ib is not None:
                ib.destroy()


def untextured(cases):
    for c in cases:
        c.texture = None
    return cases


def test_replicas_follow_host_writes(group):
    """frame 1 from buffers, then the host rewrites PART of one vertex buffer (a locked sub-range): the peers' replicas pick up
    exactly that, frame 2 equals the oracle on the new data; on a split target and on a target placed on the last peer"""
    w, h = 256, 192
    cases = untextured(make_scene(17, w, h, ndraws=4, ntri=300))
    for placement in (ogre_b200.PLACE_SPLIT, group.num_devices - 1):
        scene = BufferScene(group, cases)
        rt = ogre_b200.RenderTarget(group, w, h, placement=placement)
        scene.render(rt)
        rc, rz, _ = render_oracle(cases, w, h)
        assert_images_equal(rt.read_colour(), rt.read_depth(), rc, rz, "frame 1 (placement %d)" % placement)
        # move the second half of draw 1's vertices
        moved = cases[1].verts.copy()
        half = len(moved) // 2
        moved[half:, 0:3] += np.float32(0.35)
        vb = scene.bufs[1][0]
        vb.write(moved[half:], offset=half * 32)
        old = cases[1].verts
        cases[1].verts = moved
        scene.render(rt)
        rc, rz, _ = render_oracle(cases, w, h)
        assert_images_equal(rt.read_colour(), rt.read_depth(), rc, rz, "frame 2 after a partial host write (placement %d)" % placement)
        cases[1].verts = old
        scene.free(); rt.destroy()


def test_replicas_follow_optimised_util_results(group):
    """softwareVertexSkinning writes a HardwareBuffer on the root (host entry point with shadow pointers, the way OGRE calls it);
    a split target then draws from that buffer: the peers must see the skinned vertices, also after the next frame's call"""
    import oracle
    w, h = 256, 192
    n = 3000
    case = SkinCase(seed=5, n=n, nw=2, layout="shared", nbones=12)
    src = ogre_b200.HardwareBuffer(group, case.src.size); src.write(case.src)
    blend = ogre_b200.HardwareBuffer(group, case.blend.size); blend.write(case.blend)
    dst0, _ = case.new_dst()
    dst = ogre_b200.HardwareBuffer(group, dst0.size); dst.write(dst0)
    ou = ogre_b200.OptimisedUtil(group)
    proj = perspective(50, w / h, 0.5, 100.0)
    view = look_at(np.array([0, 0, 40.0]), np.zeros(3))
    mvp = (proj @ view).astype(np.float32)
    mvit = np.ascontiguousarray(np.linalg.inv(view).T, dtype=np.float32)
    rt = ogre_b200.RenderTarget(group, w, h, placement=ogre_b200.PLACE_SPLIT)
    for frame in range(2):
        mats = case.mats if frame == 0 else np.ascontiguousarray(case.mats[::-1])
        sp, dp = src.shadow_address(), dst.shadow_address()
        b = blend.shadow_address()
        dst.lock(ogre_b200.MAP_WRITE | ogre_b200.MAP_DISCARD)
        ou.softwareVertexSkinning(sp, dp, sp + 12, dp + 12, b + case.bw_off, b + case.bi_off, list(mats), *case.strides(), case.nw, n)
        dst.unlock()
        ref = dst0.copy()
        rsp, rdp, rsn, rdn, rbw, rbi = case.pointers(ref, None)
        oracle.skin("sse", rsp, rdp, rsn, rdn, rbw, rbi, mats, case.strides(), case.nw, n)
        d = ogre_b200.DrawDesc()
        d.positions, d.pos_stride = dst.device_ptr, 24
        d.normals, d.normal_stride = dst.device_ptr + 12, 24
        d.index_type, d.draw_count, d.vertex_count = ogre_b200.INDEX_NONE, n - n % 3, n
        d.prim_type = ogre_b200.PRIM_TRIANGLE_LIST
        d.flags = ogre_b200.DRAW_DEPTH_TEST | ogre_b200.DRAW_DEPTH_WRITE | ogre_b200.DRAW_LIGHTING
        d.mvp[:] = mvp.reshape(16); d.tex_matrix[:] = np.eye(4, dtype=np.float32).reshape(16)
        d.mvit[:] = mvit.reshape(16); d.viewport[:] = viewport_matrix(w, h).reshape(16)
        light = np.asarray((0.3, 0.5, 0.8), np.float64)
        d.light_dir[:] = (light / np.linalg.norm(light)).astype(np.float32); d.ambient[:] = (0.2, 0.2, 0.2, 1.0)      # as DrawCase normalises it
        rt.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
        rt.draw(d); rt.flush()
        got_c, got_z = rt.read_colour(), rt.read_depth()
        # what the root holds (positions bit-identical to the SSE oracle), drawn by the oracle rasteriser: the peers must have seen exactly that
        have = dst.read()
        assert np.array_equal(have[:n * 24].view(np.float32).reshape(n, 6)[:, :3], ref[:n * 24].view(np.float32).reshape(n, 6)[:, :3])
        verts = np.zeros((n, 8), np.float32)
        verts[:, 0:6] = have[:n * 24].view(np.float32).reshape(n, 6)
        oc = DrawCase(verts[:n - n % 3], None, mvp=mvp, mvit=mvit, viewport=viewport_matrix(w, h), light_dir=(0.3, 0.5, 0.8),
                      ambient=(0.2, 0.2, 0.2, 1.0), has_uv=False)
        rc, rz, _ = render_oracle([oc], w, h)
        assert_images_equal(got_c, got_z, rc, rz, "frame %d drawn from the skinned buffer" % frame)
        assert (got_z != 1.0).mean() > 0.02, "nothing was drawn"
    rt.destroy()
    for bb in (src, blend, dst):
        bb.destroy()


def test_views_on_placed_targets_equal_one_device(group):
    """config 4 mechanics on a group: 7 views, target i on device i mod G, one ogrecu_draw_views call"""
    from ogre_b200 import capi, check, lib
    w, h = 200, 152
    rng = np.random.default_rng(77)
    verts, idx = random_soup(rng, 3000, size=0.4)
    tex = random_texture(rng, 64, 32, zero_alpha=True)
    nviews = 7
    proj = perspective(50, w / h, 0.5, 100.0)
    mvps, mvits, lights, refs = [], [], [], []
    for k in range(nviews):
        view = look_at(rng.uniform(-2, 2, 3) + np.array([0, 0, 9.0]), rng.uniform(-1, 1, 3))
        mvps.append((proj @ view).astype(np.float32))
        mvits.append(np.ascontiguousarray(np.linalg.inv(view).T, dtype=np.float32))
        l = rng.normal(size=3); lights.append((l / np.linalg.norm(l)).astype(np.float32))
        c = DrawCase(verts, idx, index_type=4, texture=tex, mvp=mvps[k], mvit=mvits[k], viewport=viewport_matrix(w, h), ambient=(0.2, 0.2, 0.25, 1.0))
        c.light_dir = lights[k]
        refs.append(render_oracle([c], w, h))
    mvps, mvits, lights = [np.ascontiguousarray(np.stack(a)) for a in (mvps, mvits, lights)]
    base = DrawCase(verts, idx, index_type=4, texture=tex, mvp=mvps[0], mvit=mvits[0], viewport=viewport_matrix(w, h), ambient=(0.2, 0.2, 0.25, 1.0))
    scene = GpuScene(group, [base])
    rts = [ogre_b200.RenderTarget(group, w, h, placement=k % group.num_devices) for k in range(nviews)]
    for r in rts:
        r.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
    hs = (C.c_void_p * nviews)(*[r.handle for r in rts])
    fp = capi.c_float_p
    for rep in range(2):            # twice: the second call finds the replicas current
        check(lib().ogrecu_draw_views(hs, nviews, C.byref(scene.descs[0]), mvps.ctypes.data_as(fp), mvits.ctypes.data_as(fp), lights.ctypes.data_as(fp)))
        for k, r in enumerate(rts):
            assert_images_equal(r.read_colour(), r.read_depth(), refs[k][0], refs[k][1], "placed view %d (call %d)" % (k, rep))
            r.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
    for r in rts:
        r.destroy()
    scene.free()


def test_render_texture_on_a_peer_can_be_sampled_by_a_split_target(group):
    """a target placed on a peer is resolved into a texture (ogrecu_tex_copy_from_rt writes the root's texels from the peer), the
    texture is then sampled by every device of a split target"""
    w, h = 128, 96
    cases = untextured(make_scene(23, w, h, ndraws=3, ntri=200))
    rc, rz, _ = render_oracle(cases, w, h)
    src_rt = ogre_b200.RenderTarget(group, w, h, placement=group.num_devices - 1)
    scene = GpuScene(group, cases)
    scene.render(src_rt)
    tex = ogre_b200.Texture(group, np.zeros((h, w, 4), np.uint8))
    ogre_b200.check(ogre_b200.lib().ogrecu_tex_copy_from_rt(tex.handle, src_rt.handle))
    rgba = np.dstack([rc, np.full((h, w, 1), 255, np.uint8)])
    W, H = 320, 240
    quad = make_scene(41, W, H, ndraws=2, ntri=150)
    for c in quad:
        c.texture = rgba
    qc, qz, _ = render_oracle(quad, W, H)
    dst = ogre_b200.RenderTarget(group, W, H, placement=ogre_b200.PLACE_SPLIT)
    qs = GpuScene(group, quad)
    for d in qs.descs:
        d.texture = tex.handle
    qs.render(dst)
    assert_images_equal(dst.read_colour(), dst.read_depth(), qc, qz, "split target sampling a peer's render texture")
    qs.free(); scene.free(); tex.destroy(); dst.destroy(); src_rt.destroy()


@pytest.mark.parametrize("mode", [0, 1])
def test_split_target_sort_last_form_equals_the_oracle(group, mode):
    """a split target with a big opaque, depth-tested draw: "split_mode" 0 takes the sort-last form (every device draws its slice
    of the TRIANGLES into private planes, merged through peer memory), 1 keeps the bands; both equal the oracle, including
    pixels where coplanar duplicates of different slices tie in depth (the later slice wins), pixels nothing touches, prior
    content of the target (a second frame drawn over the first without a clear), and a blended draw that follows (bands)"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import bench_configs as bc
    w, h = 640, 480
    verts, idx = bc.terrain(380)
    tri = idx.reshape(-1, 3)
    dup = tri[1000:4000].copy()
    extra = verts[dup.reshape(-1)].copy()
    extra[:, 6:8] += 0.37
    verts2 = np.concatenate([verts, extra]).astype(np.float32)
    idx2 = np.concatenate([idx, len(verts) + np.arange(len(extra), dtype=np.uint32)]).astype(np.uint32)
    assert len(idx2) // 3 >= 262144
    mvps, mvits, lights = bc.camera_ring(3, w, h)
    tex = bc.checker_texture(256)
    big = [DrawCase(verts2, idx2, index_type=4, texture=tex, mvp=mvps[k], mvit=mvits[k], viewport=viewport_matrix(w, h), ambient=(0.2, 0.2, 0.25, 1.0))
           for k in range(2)]
    for k in range(2):
        big[k].light_dir = lights[k]
    small = make_scene(3, w, h, ndraws=3, ntri=300)          # blended / depth-write-off draws among them
    group.set_config_option("split_mode", mode)
    try:
        rt = ogre_b200.RenderTarget(group, w, h, placement=ogre_b200.PLACE_SPLIT)
        scene = GpuScene(group, big + small)
        # frame: clear, big draw (view 0); then WITHOUT a clear the big draw from view 1 over it, then the small mixed draws
        rt.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
        rt.draw(scene.descs[0]); rt.flush()
        rc, rz, _ = render_oracle(big[:1], w, h)
        assert_images_equal(rt.read_colour(), rt.read_depth(), rc, rz, "sort-last form, first frame (mode %d)" % mode)
        rt.draw(scene.descs[1])
        for d, st in list(zip(scene.descs, scene.states))[2:]:
            rt.draw(d, st)
        rt.flush()
        rc, rz, _ = render_oracle(big[1:] + small, w, h, clear=None, init=(rc, rz))
        assert_images_equal(rt.read_colour(), rt.read_depth(), rc, rz, "sort-last form over prior content + bands after it (mode %d)" % mode)
        scene.free(); rt.destroy()
    finally:
        group.set_config_option("split_mode", 0)
