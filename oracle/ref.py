"""ctypes access to oracle/_ref/lib/libref_driver.so — the UNMODIFIED reference (real
OptimisedUtilSSE/General, real TinyRenderSystem, real Root/SceneManager scenes).
TEST INFRASTRUCTURE ONLY; see oracle/ref_driver.cpp."""
import ctypes as C
import os

import numpy as np

from . import OracleDraw

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "lib", "libref_driver.so")
MEDIA = os.path.join(_HERE, "_ref", "media")
_lib = None
vp, sz, c_f = C.c_void_p, C.c_size_t, C.c_float

SSE, GENERAL = 0, 1


class SkelSubInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("src_stride", "norm_off", "blend_stride", "weight_off", "index_off", "nw",
                                          "nverts", "has_norm", "dst_stride", "nmap")] + \
               [(n, sz) for n in ("src_size", "blend_size", "skinned_size")]


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(LIB, mode=C.RTLD_GLOBAL)
        l.ref_skin.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, C.POINTER(vp), sz, sz, sz, sz, sz, sz, sz, sz]
        l.ref_morph.argtypes = [C.c_int, c_f, vp, vp, vp, sz, sz, sz, sz, C.c_int]
        l.ref_concat.argtypes = [C.c_int, vp, vp, vp, sz]
        l.ref_face_normals.argtypes = [C.c_int, vp, vp, vp, sz]
        l.ref_light_facing.argtypes = [C.c_int, vp, vp, vp, sz]
        l.ref_extrude.argtypes = [C.c_int, vp, c_f, vp, vp, sz]
        l.ref_sizeof_edge_triangle.restype = sz
        l.ref_tiny_open.argtypes = [C.c_uint32, C.c_uint32]
        l.ref_tiny_clear.argtypes = [C.c_uint, vp, c_f]
        l.ref_tiny_write.argtypes = [vp, vp]
        l.ref_tiny_read.argtypes = [vp, vp]
        l.ref_tiny_draw.argtypes = [C.POINTER(OracleDraw), sz]
        l.ref_set_media_dir.argtypes = [C.c_char_p]
        l.ref_scene_open.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, c_f]
        l.ref_scene_render.argtypes = [C.c_int]
        l.ref_scene_triangles.restype = sz
        l.ref_scene_batches.restype = sz
        l.ref_capture_count.restype = sz
        l.ref_capture_draw.argtypes = [sz]
        l.ref_capture_draw.restype = C.POINTER(OracleDraw)
        l.ref_capture_vbuf.argtypes = [sz, C.c_int, C.POINTER(sz), C.POINTER(sz)]
        l.ref_capture_vbuf.restype = vp
        l.ref_capture_ibuf.argtypes = [sz, C.POINTER(sz), C.POINTER(sz)]
        l.ref_capture_ibuf.restype = vp
        l.ref_capture_tex.argtypes = [sz]
        l.ref_capture_tex_count.restype = sz
        l.ref_capture_tex_data.argtypes = [sz, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        l.ref_capture_tex_data.restype = vp
        l.ref_scene_time.argtypes = [C.c_int]
        l.ref_scene_time.restype = C.c_double
        l.ref_skel_open.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32]
        l.ref_skel_pose.argtypes = [c_f]
        l.ref_skel_num_bones.restype = sz
        l.ref_skel_bones.restype = vp
        l.ref_skel_num_subs.restype = sz
        l.ref_skel_sub_info.argtypes = [sz, C.POINTER(SkelSubInfo)]
        for n in ("src", "blend", "map", "skinned"):
            f = getattr(l, "ref_skel_sub_" + n)
            f.argtypes = [sz]
            f.restype = vp
        l.ref_bench_skin_instances.argtypes = [vp, vp, vp, vp, sz, sz, sz, sz, sz, sz, vp, sz, sz, vp, sz, C.c_int, C.c_int]
        l.ref_bench_skin_instances.restype = C.c_double
        l.ref_bench_anim_sub.argtypes = [vp, vp, vp, vp, vp, sz, sz, sz, sz, vp, sz, vp, sz, sz, vp, vp, C.c_int, C.c_int]
        l.ref_bench_anim_sub.restype = C.c_double
        l.ref_bench_morph.argtypes = [c_f, vp, vp, vp, sz, sz, C.c_int, C.c_int, C.c_int]
        l.ref_bench_morph.restype = C.c_double
        l.ref_skel_def_sizes.argtypes = [vp, vp, vp]
        l.ref_skel_set_blend_mode.argtypes = [C.c_uint32]
        l.ref_skel_set_blend_mode.restype = None
        l.ref_skel_def_bones.argtypes = [vp, vp, vp, vp]
        l.ref_skel_def_bones.restype = None
        l.ref_skel_anim_info.argtypes = [C.c_uint32, vp, vp, vp, vp, vp, C.c_char_p]
        l.ref_skel_anim_data.argtypes = [C.c_uint32, vp, vp, vp, vp]
        l.ref_skel_anim_data.restype = None
        l.ref_skel_set_interpolation.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        l.ref_skel_pose_states.argtypes = [C.c_uint32, vp, vp, vp, vp]
        l.ref_skel_pose_states_ex.argtypes = [C.c_uint32, vp, vp, vp, vp, C.c_uint32, vp, vp, vp]
        l.ref_bench_palettes.argtypes = [C.c_uint32, vp, sz, vp]
        l.ref_bench_palettes.restype = C.c_double
        l.ref_pose_blend.argtypes = [c_f, C.c_uint32, vp, vp, C.c_uint32, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_int]
        l.ref_pose_entity.argtypes = [C.c_uint32, vp, C.c_uint32, vp, vp, vp, vp, vp, C.c_int, vp]
        l.ref_shadow_open.argtypes = [C.c_char_p, vp, vp, vp]
        l.ref_shadow_edges.argtypes = [vp, vp, vp]
        l.ref_shadow_edges.restype = None
        l.ref_shadow_volume.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]
        l.ref_shadow_volume.restype = C.c_long
        l.ref_shadow_indices.restype = vp
        l.ref_shadow_ranges.restype = vp
        l.ref_image_scale.argtypes = [vp, C.c_uint32, C.c_uint32, sz, C.c_char_p, vp, C.c_uint32, C.c_uint32, sz, C.c_char_p]
        l.ref_pixel_format_bytes.argtypes = [C.c_char_p]
        l.ref_pixel_format_bytes.restype = sz
        l.ref_set_media_dir(MEDIA.encode())
        _lib = l
    return _lib


def _bytes(ptr, n):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,)).copy() if n else np.zeros(0, np.uint8)


def tiny_render(cases, w, h, clear=(0.1, 0.2, 0.3, 1.0), depth=1.0, init=None):
    """Run DrawCases (tests/datagen.py) through the real TinyRenderSystem::_render."""
    l = lib()
    assert l.ref_tiny_open(w, h) == 0
    if init is not None:
        c0 = np.ascontiguousarray(init[0]); z0 = np.ascontiguousarray(init[1])
        l.ref_tiny_write(c0.ctypes.data, z0.ctypes.data)
    else:
        zero = np.zeros(4, np.float32)
        l.ref_tiny_clear(3, zero.ctypes.data, 0.0)
    if clear is not None:
        c = np.asarray(clear, dtype=np.float32)
        l.ref_tiny_clear(3, c.ctypes.data, depth)
    for case in cases:
        d = case.oracle_desc()
        assert l.ref_tiny_draw(C.byref(d), len(case.verts)) == 0
    colour = np.zeros((h, w, 3), np.uint8); z = np.zeros((h, w), np.float32)
    l.ref_tiny_read(colour.ctypes.data, z.ctypes.data)
    return colour, z


def capture_scene(mesh, w, h, cam_z=140.0):
    """Render `mesh` with the real scene pipeline + Tiny; returns (colour, depth, draws, textures, stats)
    where draws are dicts of numpy arrays describing every captured _render call."""
    l = lib()
    assert l.ref_scene_open(mesh.encode(), w, h, cam_z) == 0
    assert l.ref_scene_render(1) == 0
    colour = np.zeros((h, w, 3), np.uint8); z = np.zeros((h, w), np.float32)
    l.ref_tiny_read(colour.ctypes.data, z.ctypes.data)
    textures = []
    for t in range(l.ref_capture_tex_count()):
        tw, th = C.c_uint32(), C.c_uint32()
        p = l.ref_capture_tex_data(t, C.byref(tw), C.byref(th))
        textures.append(_bytes(p, tw.value * th.value * 4).reshape(th.value, tw.value, 4))
    draws = []
    for i in range(l.ref_capture_count()):
        d = l.ref_capture_draw(i).contents
        rec = {"tex": l.ref_capture_tex(i)}
        for k, name in enumerate(("pos", "uv", "nrm")):
            size, off = sz(), sz()
            p = l.ref_capture_vbuf(i, k, C.byref(size), C.byref(off))
            rec[name + "_buf"] = _bytes(p, size.value)
            rec[name + "_off"] = off.value
        size, off = sz(), sz()
        p = l.ref_capture_ibuf(i, C.byref(size), C.byref(off))
        rec["idx_buf"], rec["idx_off"] = _bytes(p, size.value), off.value
        rec["meta"] = np.array([d.pos_stride, d.uv_stride, d.normal_stride, d.index_type, d.draw_count, d.is_strip,
                                d.depth_test, d.depth_write, d.blend_add, d.lighting], dtype=np.uint32)
        rec["uniforms"] = np.concatenate([np.array(d.mvp), np.array(d.tex_matrix), np.array(d.mvit), np.array(d.viewport),
                                          np.array(d.light_dir), np.array(d.ambient)]).astype(np.float32)
        draws.append(rec)
    stats = {"triangles": l.ref_scene_triangles(), "batches": l.ref_scene_batches()}
    return colour, z, draws, textures, stats


def skeleton_pose(mesh, time, w=64, h=64, reopen=True):
    """Pose the skeletal `mesh` at `time` with the real Entity pipeline; returns (bones[nb,16], subs)."""
    l = lib()
    if reopen:
        rc = l.ref_skel_open(mesh.encode(), w, h)
        assert rc == 0, rc
    assert l.ref_skel_pose(time) == 0
    nb = l.ref_skel_num_bones()
    bones = np.ctypeslib.as_array(C.cast(l.ref_skel_bones(), C.POINTER(C.c_float)), shape=(nb, 16)).copy()
    subs = []
    for i in range(l.ref_skel_num_subs()):
        info = SkelSubInfo()
        l.ref_skel_sub_info(i, C.byref(info))
        s = {n: getattr(info, n) for n, _ in SkelSubInfo._fields_}
        s["src"] = _bytes(l.ref_skel_sub_src(i), info.src_size)
        s["blend"] = _bytes(l.ref_skel_sub_blend(i), info.blend_size)
        s["skinned"] = _bytes(l.ref_skel_sub_skinned(i), info.skinned_size)
        s["bone_map"] = np.ctypeslib.as_array(C.cast(l.ref_skel_sub_map(i), C.POINTER(C.c_uint16)), shape=(info.nmap,)).copy()
        subs.append(s)
    return bones, subs


def skeleton_def(mesh=None, w=64, h=64):
    """The skeleton of `mesh` (or of the mesh already open) as flat arrays, straight from the real SkeletonInstance:
    {num_bones, blend_mode, parent[n], bind[n,10], inherit[n,2], bind_inverse[n,10], anims: [{name, length,
    interpolation, rotation_interpolation, track_bone, track_shortest, key_offset, keys[k,11]}]}."""
    l = lib()
    if mesh is not None:
        rc = l.ref_skel_open(mesh.encode(), w, h)
        assert rc == 0, rc
    nb, na, bm = C.c_uint32(), C.c_uint32(), C.c_uint32()
    assert l.ref_skel_def_sizes(C.addressof(nb), C.addressof(na), C.addressof(bm)) == 0
    n = nb.value
    d = {"num_bones": n, "blend_mode": bm.value, "parent": np.zeros(n, np.int32), "bind": np.zeros((n, 10), np.float32),
         "inherit": np.zeros((n, 2), np.uint8), "bind_inverse": np.zeros((n, 10), np.float32), "anims": []}
    l.ref_skel_def_bones(d["parent"].ctypes.data, d["bind"].ctypes.data, d["inherit"].ctypes.data, d["bind_inverse"].ctypes.data)
    for a in range(na.value):
        length = C.c_float(); im, rim, nt, nk = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        name = C.create_string_buffer(64)
        assert l.ref_skel_anim_info(a, C.addressof(length), C.addressof(im), C.addressof(rim), C.addressof(nt), C.addressof(nk), name) == 0
        an = {"name": name.value.decode(), "length": length.value, "interpolation": im.value, "rotation_interpolation": rim.value,
              "track_bone": np.zeros(nt.value, np.int32), "track_shortest": np.zeros(nt.value, np.uint8),
              "key_offset": np.zeros(nt.value + 1, np.uint32), "keys": np.zeros((nk.value, 11), np.float32)}
        l.ref_skel_anim_data(a, an["track_bone"].ctypes.data, an["track_shortest"].ctypes.data, an["key_offset"].ctypes.data,
                             an["keys"].ctypes.data)
        d["anims"].append(an)
    return d


def skeleton_pose_states(num_bones, anim, time, weight):
    """Real SkeletonInstance::setAnimationState + _getBoneMatrices with exactly these states enabled, in this order."""
    anim = np.ascontiguousarray(anim, np.uint32); time = np.ascontiguousarray(time, np.float32)
    weight = np.ascontiguousarray(weight, np.float32)
    out = np.zeros((num_bones, 16), np.float32)
    rc = lib().ref_skel_pose_states(len(anim), anim.ctypes.data, time.ctypes.data, weight.ctypes.data, out.ctypes.data)
    assert rc == 0, rc
    return out


def skeleton_pose_states_ex(num_bones, anim, time, weight, mask=None, manual_handles=(), manual=None):
    """... with AnimationState blend masks ([states, bones]) and manually controlled bones (handles + [k, 10] local transforms)."""
    anim = np.ascontiguousarray(anim, np.uint32); time = np.ascontiguousarray(time, np.float32)
    weight = np.ascontiguousarray(weight, np.float32)
    mk = None if mask is None else np.ascontiguousarray(mask, np.float32)
    mh = np.ascontiguousarray(manual_handles, np.uint32)
    mx = np.ascontiguousarray(manual if manual is not None else np.zeros((0, 10)), np.float32)
    out = np.zeros((num_bones, 16), np.float32)
    rc = lib().ref_skel_pose_states_ex(len(anim), anim.ctypes.data, time.ctypes.data, weight.ctypes.data, None if mk is None else mk.ctypes.data,
                                       len(mh), mh.ctypes.data, mx.ctypes.data, out.ctypes.data)
    assert rc == 0, rc
    return out


def shadow_open(mesh):
    """Edge list of `mesh` (Mesh::buildEdgeList) as flat arrays {groups[g,5], edges[e,5], tris[t,3]}."""
    l = lib()
    ng, ne, nt = C.c_uint32(), C.c_uint32(), C.c_uint32()
    rc = l.ref_shadow_open(mesh.encode(), C.addressof(ng), C.addressof(ne), C.addressof(nt))
    assert rc == 0, rc
    ed = {"groups": np.zeros((ng.value, 5), np.uint32), "edges": np.zeros((ne.value, 5), np.uint32), "tris": np.zeros((nt.value, 3), np.uint32)}
    l.ref_shadow_edges(ed["groups"].ctypes.data, ed["edges"].ctypes.data, ed["tris"].ctypes.data)
    return ed


def shadow_volume(ed, light4, bounds6, flags, separate_light_cap, buffer_indices, used_in):
    """The real ShadowCaster::generateShadowVolume on the open mesh's EdgeData (light facings from the real
    EdgeData::updateTriangleLightFacing) -> (indices u16 [buffer_indices], ranges [g,4], used_out, facings[t])."""
    l = lib()
    light4 = np.ascontiguousarray(light4, np.float32); bounds6 = np.ascontiguousarray(bounds6, np.float32)
    facings = np.zeros(len(ed["tris"]), np.int8)
    used = C.c_uint32()
    n = l.ref_shadow_volume(light4.ctypes.data, bounds6.ctypes.data, flags, int(separate_light_cap), buffer_indices, used_in,
                            facings.ctypes.data, C.addressof(used))
    assert n == buffer_indices, n
    idx = np.ctypeslib.as_array(C.cast(l.ref_shadow_indices(), C.POINTER(C.c_uint16)), shape=(buffer_indices,)).copy()
    rng = np.ctypeslib.as_array(C.cast(l.ref_shadow_ranges(), C.POINTER(C.c_uint32)), shape=(len(ed["groups"]), 4)).copy()
    return idx, rng, used.value, facings


def image_scale(src, dst_w, dst_h, src_fmt="PF_BYTE_RGB", dst_fmt=None, box=None, dst_pad=0):
    """The real Image::scale(FILTER_BILINEAR) of `src` ((h, w, bytes) uint8, or its sub-box (l, t, r, b)) to dst_w x dst_h in
    `dst_fmt` (default: the source format); the destination rows are `dst_pad` pixels wider than dst_w (padding keeps 0xEE)."""
    dst_fmt = dst_fmt or src_fmt
    src = np.ascontiguousarray(src, np.uint8)
    h, w, c = src.shape
    assert c == lib().ref_pixel_format_bytes(src_fmt.encode())
    l, t, r, b = box or (0, 0, w, h)
    dc = lib().ref_pixel_format_bytes(dst_fmt.encode())
    out = np.full((dst_h, dst_w + dst_pad, dc), 0xEE, np.uint8)
    rc = lib().ref_image_scale(src.ctypes.data + (t * w + l) * c, r - l, b - t, w, src_fmt.encode(), out.ctypes.data, dst_w, dst_h,
                               dst_w + dst_pad, dst_fmt.encode())
    assert rc == 0
    return out
