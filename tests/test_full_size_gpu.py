"""GPU: the BASELINE.json configurations at their FULL sizes, checked through size-independent properties (the oracle
takes minutes to hours at these sizes) plus oracle spot checks on random samples:

  config 2  Sinbad x 10 000 instances (55.34 M vertices per step): the fused morph+skin pass equals softwareVertexMorph
            followed by softwareVertexSkinning bit for bit over the whole 1.33 GB output; instances given the same pose and
            morph weight produce identical bytes wherever they sit in the batch; 12 random instances equal the oracle.
  config 3  2 M triangles, 8 lights: the fused 8-light light-facing / extrusion launches equal eight single-light calls;
            20 000 random face normals equal the oracle.
  config 5  3840x2160, 5 M triangles (bounded for test time; the bench uses 20 M): four contiguous strips and eight
            interleaved bands compose to exactly the whole frame rendered at once (colour and depth).
  configs 4, 5 at FULL size against the CPU: two views of the 1 M-triangle scene at 640x480 and the 20 M-triangle 3840x2160
            frame equal the unmodified TinyRenderSystem::_render (oracle/_ref; about 5 s of CPU for the 4K frame) — colour
            exact, depth bit-exact.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import oracle
from datagen import bits

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config2_full_batch_properties(gpu_ctx):
    import torch
    import ogre_b200
    from ogre_b200 import SkinJob, check, lib
    sys.path.insert(0, ROOT)
    import bench
    ctx = gpu_ctx
    dev = torch.device("cuda", 0)
    subs, palettes = bench.load_sinbad()
    ninst = 10000
    pal_h, t_h = bench.instance_inputs(ninst, palettes, 0)
    # instances 0 and 9 999, 1 234 and 5 678 get identical inputs: their outputs must be identical bytes
    twins = [(0, 9999), (1234, 5678)]
    for a, b in twins:
        pal_h[b] = pal_h[a]; t_h[b] = t_h[a]
    nb = pal_h.shape[1]
    d_pal, d_t = torch.from_numpy(pal_h).to(dev), torch.from_numpy(t_h).to(dev)
    rng = np.random.default_rng(3)
    sample = sorted(set(int(i) for i in rng.integers(0, ninst, size=12)) | {0, ninst - 1})
    L = lib()
    for s in subs:
        n, nw = s["nverts"], s["nw"]
        k1, k2 = torch.from_numpy(s["key1"]).to(dev), torch.from_numpy(s["key2"]).to(dev)
        blend = torch.from_numpy(s["blend"]).to(dev)
        bmap = torch.from_numpy(s["bone_map"].view(np.int16)).to(dev)
        tmp = torch.empty(ninst, n, 6, device=dev); sep = torch.empty(ninst, n, 6, device=dev); fus = torch.empty(ninst, n, 6, device=dev)
        j = SkinJob()
        j.src_pos, j.src_norm = tmp.data_ptr(), tmp.data_ptr() + 12
        j.dst_pos, j.dst_norm = sep.data_ptr(), sep.data_ptr() + 12
        j.blend_weight, j.blend_index = blend.data_ptr() + 4, blend.data_ptr()
        j.palette, j.bone_map = d_pal.data_ptr(), bmap.data_ptr()
        j.src_pos_stride = j.dst_pos_stride = j.src_norm_stride = j.dst_norm_stride = 24
        j.blend_weight_stride = j.blend_index_stride = s["blend_stride"]
        j.num_weights, j.num_matrices, j.num_vertices = nw, len(s["bone_map"]), n
        torch.cuda.synchronize()
        check(L.ogrecu_morph_instanced_dev(ctx.handle, d_t.data_ptr(), k1.data_ptr(), k2.data_ptr(), tmp.data_ptr(), 24, 24, 24, n, 1, ninst, n * 24))
        check(L.ogrecu_skin_instanced_dev(ctx.handle, C.byref(j), ninst, nb * 16, n * 24, n * 24))
        jf = SkinJob.from_buffer_copy(j)
        jf.src_pos, jf.src_norm = k1.data_ptr(), k1.data_ptr() + 12
        jf.dst_pos, jf.dst_norm = fus.data_ptr(), fus.data_ptr() + 12
        check(L.ogrecu_anim_instanced_dev(ctx.handle, C.byref(jf), k2.data_ptr(), d_t.data_ptr(), None, 0, ninst, nb * 16, n * 24))
        ctx.sync()
        # (1) fused == separate over the whole batch, as bit patterns
        assert torch.equal(fus.view(torch.int32), sep.view(torch.int32)), "fused pass differs from morph-then-skin (nw=%d)" % nw
        # (2) twins
        for a, b in twins:
            assert torch.equal(fus[a].view(torch.int32), fus[b].view(torch.int32))
        # (3) oracle spot checks
        got = fus[sample].cpu().numpy()
        for r, i in enumerate(sample):
            mats = np.ascontiguousarray(pal_h[i][s["bone_map"]])
            keep, pp = oracle.palette_ptrs(mats)
            src = np.zeros((n, 6), np.float32); ref = np.zeros((n, 6), np.float32)
            oracle.lib().oracle_morph(float(t_h[i]), s["key1"].ctypes.data, s["key2"].ctypes.data, src.ctypes.data, 24, 24, 24, n, 1)
            bl = s["blend"].ctypes.data
            oracle.lib().oracle_skin_sse(src.ctypes.data, ref.ctypes.data, src.ctypes.data + 12, ref.ctypes.data + 12, bl + 4, bl, pp,
                                         24, 24, 24, 24, s["blend_stride"], s["blend_stride"], nw, n)
            assert np.array_equal(bits(got[r]), bits(ref)), (nw, i)
        del tmp, sep, fus


def test_config3_fused_lights_equal_single_calls(gpu_ctx):
    from ogre_b200 import addr, capi, check, lib
    ctx = gpu_ctx
    nu = 1000
    nv, nt = nu * nu, 2 * nu * nu
    u, v = np.meshgrid(np.linspace(0, 2 * np.pi, nu, endpoint=False), np.linspace(0, 2 * np.pi, nu, endpoint=False), indexing="ij")
    pos = np.stack([(3 + np.cos(v)) * np.cos(u), np.sin(v), (3 + np.cos(v)) * np.sin(u)], -1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(nu), np.arange(nu), indexing="ij")
    a = i * nu + j; b = ((i + 1) % nu) * nu + j; c = i * nu + (j + 1) % nu; d = ((i + 1) % nu) * nu + (j + 1) % nu
    tri_idx = np.stack([np.stack([a, b, c], -1), np.stack([b, d, c], -1)], 2).reshape(-1, 3)
    tris = np.zeros((nt, 8), np.uint32)
    tris[:, 2:5] = tri_idx; tris[:, 5:8] = tri_idx
    lights = np.array([[12 * np.cos(k * np.pi / 4), 6.0, 12 * np.sin(k * np.pi / 4), float(k != 3)] for k in range(8)], np.float32)
    L = lib()
    lp = lights.ctypes.data_as(capi.c_float_p)
    dpos, dtri = ctx.to_device(pos), ctx.to_device(tris)
    dfn = ctx.alloc(nt * 16)
    check(L.ogrecu_face_normals_dev(ctx.handle, dpos.ptr, dtri.ptr, dfn.ptr, nt))
    fn = dfn.download(np.float32).reshape(nt, 4)
    rng = np.random.default_rng(1)
    pick = rng.integers(0, nt, size=20000)
    ref = np.zeros((len(pick), 4), np.float32)
    sub = np.ascontiguousarray(tris[pick])
    oracle.lib().oracle_face_normals(addr(pos), addr(sub), addr(ref), len(pick))
    assert np.array_equal(bits(fn[pick]), bits(ref))
    # light facing: 8 fused == 8 single
    dfa, dfb = ctx.alloc(8 * nt), ctx.alloc(8 * nt)
    check(L.ogrecu_light_facing_multi_dev(ctx.handle, lp, 8, dfn.ptr, dfa.ptr, nt, nt))
    for k in range(8):
        check(L.ogrecu_light_facing_dev(ctx.handle, C.cast(lights[k].ctypes.data, capi.c_float_p), dfn.ptr, dfb.ptr + k * nt, nt))
    fa, fb = dfa.download(np.int8), dfb.download(np.int8)
    assert np.array_equal(fa, fb)
    assert 0.2 < fa.mean() < 0.8          # a closed mesh lit from outside: roughly half the faces face each light
    # extrusion: 8 fused == 8 single, and a sample equals the oracle
    dea, deb = ctx.alloc(8 * nv * 12), ctx.alloc(8 * nv * 12)
    check(L.ogrecu_extrude_multi_dev(ctx.handle, lp, 8, 40.0, dpos.ptr, dea.ptr, nv * 3, nv))
    for k in range(8):
        check(L.ogrecu_extrude_dev(ctx.handle, C.cast(lights[k].ctypes.data, capi.c_float_p), 40.0, dpos.ptr, deb.ptr + k * nv * 12, nv))
    ea, eb = dea.download(np.float32).reshape(8, nv, 3), deb.download(np.float32).reshape(8, nv, 3)
    assert np.array_equal(bits(ea), bits(eb))
    for k in (0, 3, 7):
        ref = np.zeros((5000, 3), np.float32)
        l = np.ascontiguousarray(lights[k])
        oracle.lib().oracle_extrude(addr(l), 40.0, addr(pos), addr(ref), 5000)
        assert np.array_equal(bits(ea[k, :5000]), bits(ref))
    for b_ in (dpos, dtri, dfn, dfa, dfb, dea, deb):
        b_.free()


def test_config5_split_frames_compose_at_4k(gpu_ctx):
    import ogre_b200
    from ogre_b200 import capi, check, lib
    from ogre_b200.shard import band_row_indices, strip_rows
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs as bc
    ctx = gpu_ctx
    w, h = 3840, 2160
    n = int(round((5000000 / 2) ** 0.5))
    verts, idx = bc.terrain(n)
    desc, keep = bc.upload_scene(ctx, verts, idx, bc.checker_texture(1024), w, h)
    mvps, mvits, lights = bc.camera_ring(1, w, h)
    rt = ogre_b200.RenderTarget(ctx, w, h)
    handles = (C.c_void_p * 1)(rt.handle)

    def render():
        rt.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
        check(lib().ogrecu_draw_views(handles, 1, C.byref(desc), mvps.ctypes.data_as(capi.c_float_p), mvits.ctypes.data_as(capi.c_float_p),
                                      lights.ctypes.data_as(capi.c_float_p)))
    render()
    whole_c, whole_z = rt.read_colour(), rt.read_depth()
    assert (whole_z != 1.0).mean() > 0.2
    comp_c, comp_z = np.zeros_like(whole_c), np.zeros_like(whole_z)
    for r in range(4):                                    # contiguous strips
        y0, y1 = strip_rows(h, r, 4)
        rt.set_strip(y0, y1)
        render()
        comp_c[y0:y1] = rt.read_colour((0, y0, w, y1)); comp_z[y0:y1] = rt.read_depth((0, y0, w, y1))
    assert np.array_equal(comp_c, whole_c) and np.array_equal(bits(comp_z), bits(whole_z))
    comp_c[:] = 0; comp_z[:] = 0
    for r in range(8):                                    # interleaved bands of one tile row
        rt.set_bands(16, r, 8)
        render()
        rows = band_row_indices(h, r, 8, 16)
        c, z = rt.read_colour(), rt.read_depth()
        comp_c[rows] = c[rows]; comp_z[rows] = z[rows]
    assert np.array_equal(comp_c, whole_c) and np.array_equal(bits(comp_z), bits(whole_z))
    rt.destroy()


def _cpu_frame(case, w, h):
    """colour, depth of one DrawCase on the CPU: the unmodified TinyRenderSystem::_render (oracle/_ref) where it is built,
    else the oracle's C restatement (pinned against it by tests/test_oracle_vs_ref.py)"""
    from oracle import ref
    if ref.available():
        c, z = ref.tiny_render([case], w, h)[:2]
        return c, z, "reference"
    from rasterlib import render_oracle
    c, z, _ = render_oracle([case], w, h)
    return c, z, "port"


@pytest.mark.parametrize("name,ntri,w,h,view", [("config4 view", 1000000, 640, 480, 0), ("config4 view", 1000000, 640, 480, 517),
                                                ("config5 frame", 20000000, 3840, 2160, 0)],
                         ids=["c4_view0", "c4_view517", "c5_frame"])
def test_full_size_raster_configs_equal_the_cpu_reference(gpu_ctx, name, ntri, w, h, view):
    """BASELINE configs 4 and 5 at their full sizes (1 M triangles at 640x480; 20 M triangles at 3840x2160, the scenes
    bench.py renders) against the CPU: colour exact, depth bit-exact."""
    import ogre_b200
    from ogre_b200 import capi, check, lib
    from datagen import DrawCase, viewport_matrix
    from rasterlib import assert_images_equal
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs as bc
    ctx = gpu_ctx
    n = int(round((ntri / 2) ** 0.5))
    verts, idx = bc.terrain(n)
    tex = bc.checker_texture(1024)
    mvps, mvits, lights = bc.camera_ring(1024 if view else 1, w, h)
    mvp, mvit, light = (np.ascontiguousarray(a[view:view + 1]) for a in (mvps, mvits, lights))
    desc, keep = bc.upload_scene(ctx, verts, idx, tex, w, h)
    rt = ogre_b200.RenderTarget(ctx, w, h)
    rt.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
    handles = (C.c_void_p * 1)(rt.handle)
    check(lib().ogrecu_draw_views(handles, 1, C.byref(desc), mvp.ctypes.data_as(capi.c_float_p), mvit.ctypes.data_as(capi.c_float_p),
                                  light.ctypes.data_as(capi.c_float_p)))
    got_c, got_z = rt.read_colour(), rt.read_depth()
    rt.destroy()
    for k in keep[:2]:
        k.free()
    case = DrawCase(verts, idx, index_type=4, texture=tex, mvp=mvp[0], mvit=mvit[0], viewport=viewport_matrix(w, h), light_dir=(1, 0, 0),
                    ambient=(0.2, 0.2, 0.25, 1.0))
    case.light_dir = light[0]
    ref_c, ref_z, kind = _cpu_frame(case, w, h)
    assert (ref_z != 1.0).mean() > 0.2, "the scene must cover the frame"
    assert_images_equal(got_c, got_z, ref_c, ref_z, "%s vs the CPU %s" % (name, kind))
