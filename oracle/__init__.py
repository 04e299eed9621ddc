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
