import sys, os, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import ogre_b200
from ogre_b200 import capi, check, lib
import bench_configs as bc
from datagen import DrawCase, viewport_matrix
from rasterlib import render_oracle, render_gpu

ctx = ogre_b200.Context(0)
w, h = 640, 480
verts, idx = bc.terrain(60)
tex = bc.checker_texture(256)
mvps, mvits, lights = bc.camera_ring(4, w, h)
refs = []
for k in range(3):
    c = DrawCase(verts, idx, index_type=4, texture=tex, mvp=mvps[k], mvit=mvits[k], viewport=viewport_matrix(w, h), light_dir=(1, 0, 0), ambient=(0.2, 0.2, 0.25, 1.0))
    c.light_dir = lights[k]
    refs.append(render_oracle([c], w, h)[0])
desc, keep = bc.upload_scene(ctx, verts, idx, tex, w, h)
fp = capi.c_float_p

def views(sel, use_desc=False):
    rts = [ogre_b200.RenderTarget(ctx, w, h) for _ in sel]
    for r in rts: r.clear(colour=(0.1, 0.2, 0.3, 1.0), depth=1.0)
    hs = (C.c_void_p * len(sel))(*[r.handle for r in rts])
    m = np.ascontiguousarray(mvps[sel]); mi = np.ascontiguousarray(mvits[sel]); li = np.ascontiguousarray(lights[sel])
    if use_desc:
        desc.mvp[:] = m[0].reshape(16); desc.mvit[:] = mi[0].reshape(16); desc.light_dir[:] = li[0]
        check(lib().ogrecu_draw_views(hs, len(sel), C.byref(desc), m.ctypes.data_as(fp), None, None))
    else:
        check(lib().ogrecu_draw_views(hs, len(sel), C.byref(desc), m.ctypes.data_as(fp), mi.ctypes.data_as(fp), li.ctypes.data_as(fp)))
    out = [r.read_colour() for r in rts]
    for r in rts: r.destroy()
    return out

for sel, ud in (([0], False), ([0], True), ([1], False), ([0, 1], False), ([1, 0], False), ([0, 1, 2], False)):
    out = views(sel, ud)
    print(sel, "desc-uniforms" if ud else "per-view", [int((o != refs[k]).any(axis=2).sum()) for o, k in zip(out, sel)])
