"""Helpers that run a list of DrawCase through the CPU oracle and through the CUDA C ABI."""
import ctypes as C

import numpy as np


def render_oracle(cases, w, h, clear=(0.1, 0.2, 0.3, 1.0), depth=1.0, init=None):
    import oracle
    lib = oracle.lib()
    colour = np.zeros((h, w, 3), dtype=np.uint8)
    zbuf = np.zeros((h, w), dtype=np.float32)
    if init is not None:
        colour[:], zbuf[:] = init
    if clear is not None:
        c = np.asarray(clear, dtype=np.float32)
        lib.oracle_tiny_clear(colour.ctypes.data, zbuf.ctypes.data, w, h, 3, c.ctypes.data, depth)
    frags = 0
    for case in cases:
        d = case.oracle_desc()
        st = case.oracle_state()
        if st is None:
            frags += lib.oracle_tiny_draw(C.byref(d), colour.ctypes.data, zbuf.ctypes.data, w, h)
        else:
            frags += lib.oracle_tiny_draw_state(C.byref(d), C.byref(st), colour.ctypes.data, zbuf.ctypes.data, w, h)
    return colour, zbuf, frags


class GpuScene:
    """Uploads DrawCases once and issues them through ogrecu_draw."""

    def __init__(self, ctx, cases):
        import ogre_b200
        self.ctx, self.cases = ctx, cases
        self.keep, self.descs, self.states, self.textures = [], [], [], {}
        for case in cases:
            vb = ctx.to_device(case.verts)
            ib = ctx.to_device(case.indices) if case.indices is not None else None
            self.keep += [vb, ib]
            d = ogre_b200.DrawDesc()
            d.positions, d.pos_stride = vb.ptr, 32
            d.normals, d.normal_stride = (vb.ptr + 12, 32) if case.has_normals else (None, 0)
            d.uvs, d.uv_stride = (vb.ptr + 24, 32) if case.has_uv else (None, 0)
            d.indices = ib.ptr if ib is not None else None
            d.index_type, d.draw_count = case.index_type, case.draw_count
            d.vertex_count = len(case.verts) if getattr(case, "give_vertex_count", True) else 0
            d.prim_type = ogre_b200.PRIM_TRIANGLE_STRIP if case.strip else ogre_b200.PRIM_TRIANGLE_LIST
            d.flags = ((ogre_b200.DRAW_DEPTH_TEST if case.depth_test else 0) | (ogre_b200.DRAW_DEPTH_WRITE if case.depth_write else 0) |
                       (ogre_b200.DRAW_BLEND_ADD if case.blend_add else 0) | (ogre_b200.DRAW_LIGHTING if case.lighting else 0))
            if case.texture is not None:
                key = id(case.texture)
                if key not in self.textures:
                    self.textures[key] = ogre_b200.Texture(ctx, case.texture)
                d.texture = self.textures[key].handle
            d.mvp[:] = case.mvp.reshape(16); d.tex_matrix[:] = case.tex_matrix.reshape(16)
            d.mvit[:] = case.mvit.reshape(16); d.viewport[:] = case.viewport.reshape(16)
            d.light_dir[:] = case.light_dir; d.ambient[:] = case.ambient
            self.descs.append(d)
            st = None
            if case.state:
                cb = ctx.to_device(case.state["colours"]) if "colours" in case.state else None
                self.keep.append(cb)
                st = case.fill_state(ogre_b200.DrawState(), cb.ptr if cb is not None else None)
            self.states.append(st)

    def render(self, rt, clear=(0.1, 0.2, 0.3, 1.0), depth=1.0):
        if clear is not None:
            rt.clear(colour=clear, depth=depth)
        for d, st in zip(self.descs, self.states):
            rt.draw(d, st)
        rt.flush()

    def free(self):
        for t in self.textures.values():
            t.destroy()
        for b in self.keep:
            if b is not None:
                b.free()


def render_gpu(ctx, cases, w, h, clear=(0.1, 0.2, 0.3, 1.0), depth=1.0, init=None):
    import ogre_b200
    rt = ogre_b200.RenderTarget(ctx, w, h)
    if init is not None:
        rt.write_colour(init[0]); rt.write_depth(init[1])
    scene = GpuScene(ctx, cases)
    scene.render(rt, clear, depth)
    colour, zbuf = rt.read_colour(), rt.read_depth()
    stats = rt.stats()
    scene.free(); rt.destroy()
    return colour, zbuf, stats


def assert_images_equal(got_c, got_z, ref_c, ref_z, what=""):
    # bit patterns; NaN payload/sign is canonicalised (x86 produces 0xFFC00000 for 0*inf, the GPU
    # 0x7FFFFFFF - both are "NaN", which is all the depth test can observe)
    zb_g, zb_r = got_z.view(np.uint32).copy(), ref_z.view(np.uint32).copy()
    zb_g[np.isnan(got_z)] = 0x7FC00000
    zb_r[np.isnan(ref_z)] = 0x7FC00000
    nz = int((zb_g != zb_r).sum())
    nc = int((got_c != ref_c).any(axis=2).sum())
    if nz or nc:
        ys, xs = np.nonzero((zb_g != zb_r) | (got_c != ref_c).any(axis=2))
        raise AssertionError("%s: %d depth-bit and %d colour mismatches; first at (x=%d,y=%d): got z=%r c=%s, ref z=%r c=%s" % (
            what, nz, nc, xs[0], ys[0], got_z[ys[0], xs[0]], got_c[ys[0], xs[0]], ref_z[ys[0], xs[0]], ref_c[ys[0], xs[0]]))
