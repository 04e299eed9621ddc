"""oracle — CPU restatement of the reference algorithms (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product path (ogre_b200/, libogrecu.so, the plugin) never
does.  See oracle/ogre_oracle.h for the reference file:line each function restates and
for the pinning status.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libogre_oracle.so")
_lib = None

c_f = C.c_float
vp = C.c_void_p
sz = C.c_size_t


class OracleDraw(C.Structure):
    """oracle_draw (oracle/ogre_oracle.h)"""
    _fields_ = [
        ("positions", vp), ("pos_stride", C.c_uint32),
        ("uvs", vp), ("uv_stride", C.c_uint32),
        ("normals", vp), ("normal_stride", C.c_uint32),
        ("indices", vp), ("index_type", C.c_uint32),
        ("draw_count", C.c_uint32), ("is_strip", C.c_uint32),
        ("depth_test", C.c_uint32), ("depth_write", C.c_uint32), ("blend_add", C.c_uint32), ("lighting", C.c_uint32),
        ("tex_rgba", vp), ("tex_w", C.c_uint32), ("tex_h", C.c_uint32),
        ("mvp", c_f * 16), ("tex_matrix", c_f * 16), ("mvit", c_f * 16), ("viewport", c_f * 16),
        ("light_dir", c_f * 3), ("ambient", c_f * 4),
    ]


class OracleDrawState(C.Structure):
    """oracle_draw_state (oracle/ogre_oracle.h): the opt-in "beyond Tiny" render state"""
    _fields_ = [("enable", C.c_uint32), ("depth_func", C.c_uint32), ("cull_mode", C.c_uint32), ("alpha_func", C.c_uint32),
                ("alpha_ref", C.c_uint32), ("scissor", C.c_uint32 * 4), ("colour_stride", C.c_uint32), ("colours", vp)]


class OracleSkeleton(C.Structure):
    """oracle_skeleton (oracle/ogre_oracle.h)"""
    _fields_ = [("num_bones", C.c_uint32), ("blend_mode", C.c_uint32), ("parent", vp), ("bind", vp), ("inherit", vp)]


class OracleAnimation(C.Structure):
    """oracle_animation (oracle/ogre_oracle.h)"""
    _fields_ = [("length", c_f), ("interpolation", C.c_uint32), ("rotation_interpolation", C.c_uint32), ("num_tracks", C.c_uint32),
                ("track_bone", vp), ("track_shortest", vp), ("key_offset", vp), ("keys", vp)]


class Skeleton:
    """Oracle-side skeleton built from the flat dict of oracle.ref.skeleton_def / tests.goldenlib.load_skeleton."""

    def __init__(self, d):
        self.d = d
        self.keep = [np.ascontiguousarray(d["parent"], np.int32), np.ascontiguousarray(d["bind"], np.float32),
                     np.ascontiguousarray(d["inherit"], np.uint8)]
        self.sk = OracleSkeleton(d["num_bones"], d["blend_mode"], *[a.ctypes.data for a in self.keep])
        self.anims = (OracleAnimation * max(1, len(d["anims"])))()
        for i, a in enumerate(d["anims"]):
            arrs = [np.ascontiguousarray(a["track_bone"], np.int32), np.ascontiguousarray(a["track_shortest"], np.uint8),
                    np.ascontiguousarray(a["key_offset"], np.uint32), np.ascontiguousarray(a["keys"], np.float32)]
            self.keep += arrs
            self.anims[i] = OracleAnimation(a["length"], a["interpolation"], a["rotation_interpolation"], len(arrs[0]),
                                            *[x.ctypes.data for x in arrs])
        self.bind_inverse = np.zeros((d["num_bones"], 10), np.float32)
        lib().oracle_skeleton_bind_inverse(C.byref(self.sk), self.bind_inverse.ctypes.data)

    def palette(self, anim, time, weight, bind_inverse=None, blend_mode=None, mask=None, manual_slot=None, manual=None):
        """mask: [states, bones] blend masks; manual_slot [bones] (row of `manual` or -1), manual [k, 10] local transforms"""
        self.sk.blend_mode = self.d["blend_mode"] if blend_mode is None else int(blend_mode)
        anim = np.ascontiguousarray(anim, np.int32); time = np.ascontiguousarray(time, np.float32)
        weight = np.ascontiguousarray(weight, np.float32)
        bi = self.bind_inverse if bind_inverse is None else np.ascontiguousarray(bind_inverse, np.float32)
        out = np.zeros((self.d["num_bones"], 16), np.float32)
        mk = None if mask is None else np.ascontiguousarray(mask, np.float32)
        ms = None if manual_slot is None else np.ascontiguousarray(manual_slot, np.int32)
        mx = None if manual is None else np.ascontiguousarray(manual, np.float32)
        rc = lib().oracle_skeleton_palette_ex(C.byref(self.sk), bi.ctypes.data, self.anims, len(anim), anim.ctypes.data,
                                              time.ctypes.data, weight.ctypes.data, None if mk is None else mk.ctypes.data,
                                              None if ms is None else ms.ctypes.data, None if mx is None else mx.ctypes.data, out.ctypes.data)
        if rc:
            raise ValueError("interpolation mode not restated by the oracle")
        return out


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        l = C.CDLL(_LIB)
        skin_args = [vp, vp, vp, vp, vp, vp, C.POINTER(vp), sz, sz, sz, sz, sz, sz, sz, sz]
        l.oracle_skin_general.argtypes = skin_args
        l.oracle_skin_sse.argtypes = skin_args
        l.oracle_morph.argtypes = [c_f, vp, vp, vp, sz, sz, sz, sz, C.c_int]
        l.oracle_concat_affine_sse.argtypes = [vp, vp, vp, sz]
        l.oracle_concat_affine_general.argtypes = [vp, vp, vp, sz]
        l.oracle_face_normals.argtypes = [vp, vp, vp, sz]
        l.oracle_light_facing_sse.argtypes = [vp, vp, vp, sz]
        l.oracle_light_facing_general.argtypes = [vp, vp, vp, sz]
        l.oracle_extrude.argtypes = [vp, c_f, vp, vp, sz]
        l.oracle_tiny_clear.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint, vp, c_f]
        l.oracle_tiny_draw.argtypes = [C.POINTER(OracleDraw), vp, vp, C.c_uint32, C.c_uint32]
        l.oracle_tiny_draw.restype = C.c_uint64
        l.oracle_tiny_draw_state.argtypes = [C.POINTER(OracleDraw), C.POINTER(OracleDrawState), vp, vp, C.c_uint32, C.c_uint32]
        l.oracle_tiny_draw_state.restype = C.c_uint64
        l.oracle_pose_blend.argtypes = [c_f, vp, vp, C.c_uint32, vp, C.c_uint32]
        l.oracle_pose_blend.restype = None
        l.oracle_finalise_pose_normals.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32]
        l.oracle_finalise_pose_normals.restype = None
        l.oracle_scale_bilinear_bytes.argtypes = [vp, C.c_uint32, C.c_uint32, sz, C.c_uint32, vp, C.c_uint32, C.c_uint32, sz]
        l.oracle_scale_bilinear_bytes.restype = None
        l.oracle_shadow_use_mcguire.argtypes = [C.c_uint32, vp, vp]
        l.oracle_shadow_volume.argtypes = [C.c_uint32, vp, vp, vp, vp, C.c_int, C.c_int, C.c_uint32, C.c_int, vp, C.c_uint32, vp, vp]
        l.oracle_shadow_volume.restype = C.c_long
        l.oracle_skeleton_bind_inverse.argtypes = [C.POINTER(OracleSkeleton), vp]
        l.oracle_skeleton_bind_inverse.restype = None
        l.oracle_skeleton_palette.argtypes = [C.POINTER(OracleSkeleton), vp, C.POINTER(OracleAnimation), C.c_uint32, vp, vp, vp, vp]
        l.oracle_skeleton_palette_ex.argtypes = [C.POINTER(OracleSkeleton), vp, C.POINTER(OracleAnimation), C.c_uint32, vp, vp, vp, vp, vp, vp, vp]
        for n in ("oracle_skin_general", "oracle_skin_sse", "oracle_morph", "oracle_concat_affine_sse",
                  "oracle_concat_affine_general", "oracle_face_normals", "oracle_light_facing_sse",
                  "oracle_light_facing_general", "oracle_extrude", "oracle_tiny_clear"):
            getattr(l, n).restype = None
        _lib = l
    return _lib


def _p(a, off=0):
    return None if a is None else a.ctypes.data + off


def palette_ptrs(mats):
    """(N,16) float32 -> ctypes array of row pointers (the reference's pointer palette)."""
    mats = np.ascontiguousarray(mats, dtype=np.float32).reshape(-1, 16)
    arr = (vp * len(mats))(*[mats.ctypes.data + 64 * i for i in range(len(mats))])
    return mats, arr


def skin(kind, src_pos, dst_pos, src_norm, dst_norm, weight, index, mats, strides, nw, n):
    """kind: 'sse' | 'general'.  Pointers are integer addresses, strides a 6-tuple
    (srcPos, dstPos, srcNorm, dstNorm, weight, index)."""
    keep, pal = palette_ptrs(mats)
    fn = lib().oracle_skin_sse if kind == "sse" else lib().oracle_skin_general
    fn(src_pos, dst_pos, src_norm, dst_norm, weight, index, pal, *strides, nw, n)
    return keep


def shadow_volume(ed, facings, light4, bounds6, flags, separate_light_cap, buffer_indices, used_in, fill=0xCDCD):
    """generateShadowVolume on the flat edge data `ed` = {groups[g,5], edges[e,5], tris[t,3]} -> (indices u16
    [buffer_indices], ranges [g,4], used_out, emitted, use_mcguire)."""
    light4 = np.ascontiguousarray(light4, np.float32); bounds6 = np.ascontiguousarray(bounds6, np.float32)
    mc = lib().oracle_shadow_use_mcguire(len(ed["groups"]), bounds6.ctypes.data, light4.ctypes.data)
    out = np.full(buffer_indices, fill, np.uint16)
    ranges = np.zeros((len(ed["groups"]), 4), np.uint32)
    used = C.c_uint32(used_in)
    f = np.ascontiguousarray(facings, np.int8)
    n = lib().oracle_shadow_volume(len(ed["groups"]), ed["groups"].ctypes.data, ed["edges"].ctypes.data, ed["tris"].ctypes.data, f.ctypes.data,
                                   int(light4[3] == 0), mc, flags, int(separate_light_cap), out.ctypes.data, buffer_indices,
                                   C.addressof(used), ranges.ctypes.data)
    return out, ranges, used.value, n, mc


def scale_bilinear(src, dst_w, dst_h, box=None, dst_pad=0):
    """LinearResampler_Byte on a (h, w, channels) uint8 image (or its sub-box (l, t, r, b)) -> (dst_h, dst_w + dst_pad, channels);
    padding pixels keep 0xEE."""
    src = np.ascontiguousarray(src, np.uint8)
    h, w, c = src.shape
    l, t, r, b = box or (0, 0, w, h)
    out = np.full((dst_h, dst_w + dst_pad, c), 0xEE, np.uint8)
    lib().oracle_scale_bilinear_bytes(src.ctypes.data + (t * w + l) * c, r - l, b - t, w, c, out.ctypes.data, dst_w, dst_h, dst_w + dst_pad)
    return out
