"""GPU: the CUDA path (through the C ABI) vs golden fixtures dumped from the unmodified
reference — config 1 (ogrehead image-identical to Tiny) and config 2 (Sinbad / jaiqua skinned
exactly as OGRE's own Entity pipeline skins them)."""
import ctypes as C

import numpy as np
import pytest

from datagen import bits
from goldenlib import SkelFixture, load_scene
from rasterlib import assert_images_equal, render_gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ogrehead_320x180.npz", "ogrehead_1280x720.npz"])
def test_ogrehead_image_identical_to_tiny(gpu_ctx, name):
    cases, colour, depth, (w, h) = load_scene(name)
    gc, gz, stats = render_gpu(gpu_ctx, cases, w, h, clear=(0.1, 0.2, 0.3, 1.0), depth=1.0)
    assert stats["triangles_in"] == 4336
    # north_star: coverage + depth bit-exact, colour <= 1 LSB; we require colour exact too
    assert_images_equal(gc, gz, colour, depth, name)


@pytest.mark.parametrize("name", ["sinbad_skin.npz", "jaiqua_skin.npz"])
def test_skinned_entities_match_ogre(gpu_ctx, name):
    import ogre_b200
    from ogre_b200 import SkinJob, check, lib
    ctx = gpu_ctx
    F = SkelFixture(name)
    times = (3, 17)
    pals = np.ascontiguousarray(F.palettes[list(times)])            # [2, nbones, 16]
    dpal = ctx.to_device(pals)
    for s in F.subs:
        n, nw = s["nverts"], s["nw"]
        dsrc, dbl, dmap = ctx.to_device(s["src"]), ctx.to_device(s["blend"]), ctx.to_device(s["bone_map"])
        ddst = ctx.alloc(len(times) * n * 24)
        j = SkinJob()
        j.src_pos, j.src_norm = dsrc.ptr, dsrc.ptr + s["norm_off"]
        j.dst_pos, j.dst_norm = ddst.ptr, ddst.ptr + 12
        j.blend_weight, j.blend_index = dbl.ptr + s["weight_off"], dbl.ptr + s["index_off"]
        j.palette, j.bone_map = dpal.ptr, dmap.ptr
        j.src_pos_stride = j.src_norm_stride = s["src_stride"]
        j.dst_pos_stride = j.dst_norm_stride = 24
        j.blend_weight_stride = j.blend_index_stride = s["blend_stride"]
        j.num_weights, j.num_matrices, j.num_vertices = nw, len(s["bone_map"]), n
        # two "instances" = the two poses, one launch (the config-2 instanced path)
        check(lib().ogrecu_skin_instanced_dev(ctx.handle, C.byref(j), len(times), F.nbones * 16, 0, n * 24))
        got = ddst.download(np.float32).reshape(len(times), n, 6)
        for k, t in enumerate(times):
            ref = s["skinned"][t].view(np.float32).reshape(n, 6)
            assert np.array_equal(bits(got[k][:, :3]), bits(ref[:, :3])), (name, t, nw)     # positions: bit-exact vs SSE
            assert np.abs(got[k][:, 3:] - ref[:, 3:]).max() <= 4e-4                          # normals: SSE rsqrt estimate
            rel = np.abs(got[k][:, :3] - ref[:, :3]).max() / max(1.0, np.abs(ref[:, :3]).max())
            assert rel <= 1e-5                                                              # north_star tolerance
        for b in (dsrc, dbl, dmap, ddst):
            b.free()
    dpal.free()
