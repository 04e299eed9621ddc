"""Seeded synthetic inputs shared by the CPU and GPU parity tests (SURVEY §8c/§8d shapes)."""
import numpy as np


def rigid_matrices(rng, n, trans=5.0):
    """n random rigid Affine3 (rotation + translation), float32 [n,16], row 3 = 0,0,0,1."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    M = np.zeros((n, 4, 4), dtype=np.float32)
    M[:, :3, :3] = R
    M[:, :3, 3] = rng.uniform(-trans, trans, size=(n, 3))
    M[:, 3, 3] = 1
    return np.ascontiguousarray(M.reshape(n, 16), dtype=np.float32)


class SkinCase:
    """One softwareVertexSkinning call laid out in raw byte buffers.

    layout: 'shared' (pos+normal interleaved, stride 24), 'separate' (two packed streams,
    stride 12), 'posonly' (stride 12, no normals), 'fat' (pos+normal+uv interleaved stride 32,
    destination stride 32 with untouched gaps).
    head: byte offset of the first vertex inside the source allocation (misaligned heads).
    """

    def __init__(self, seed, n, nw, layout="shared", nbones=64, head=0, blend_interleaved=True):
        rng = np.random.default_rng(seed)
        self.n, self.nw, self.layout = n, nw, layout
        self.mats = rigid_matrices(rng, nbones)
        pos = rng.uniform(-10, 10, size=(n, 3)).astype(np.float32)
        nrm = rng.normal(size=(n, 3))
        nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
        w = rng.uniform(0.05, 1, size=(n, nw))
        w = (w / w.sum(axis=1, keepdims=True)).astype(np.float32)
        idx = rng.integers(0, nbones, size=(n, nw)).astype(np.uint8)
        self.has_norm = layout != "posonly"
        if layout == "shared":
            s, self.norm_off = 24, 12
        elif layout == "fat":
            s, self.norm_off = 32, 12
        else:
            s, self.norm_off = 12, None
        self.sp_stride = self.dp_stride = s
        self.head = head
        self.src = np.zeros(head + n * s + 16, dtype=np.uint8)
        sv = self.src[head:head + n * s].reshape(n, s)
        sv[:, 0:12] = pos.view(np.uint8).reshape(n, 12)
        if layout in ("shared", "fat"):
            sv[:, 12:24] = nrm.view(np.uint8).reshape(n, 12)
            self.sn_stride = self.dn_stride = s
            self.src_n = None
        elif layout == "separate":
            self.src_n = np.ascontiguousarray(nrm).view(np.uint8).reshape(-1).copy()
            self.sn_stride = self.dn_stride = 12
        else:
            self.src_n = None
            self.sn_stride = self.dn_stride = 0
        # blend stream: UBYTE4 indices @0 + nw float weights @4, stride 4+4nw (OgreMesh.cpp:775-800)
        if blend_interleaved:
            bs = 4 + 4 * nw
            self.blend = np.zeros(n * bs, dtype=np.uint8)
            bv = self.blend.reshape(n, bs)
            bv[:, 0:nw] = idx
            bv[:, 4:4 + 4 * nw] = w.view(np.uint8).reshape(n, 4 * nw)
            self.bw_off, self.bi_off, self.bw_stride, self.bi_stride = 4, 0, bs, bs
            self.blend_i = None
        else:
            self.blend = np.ascontiguousarray(w).view(np.uint8).reshape(-1).copy()
            self.blend_i = np.ascontiguousarray(idx).reshape(-1).copy()
            self.bw_off, self.bi_off, self.bw_stride, self.bi_stride = 0, 0, 4 * nw, nw

    def new_dst(self, fill=0xCD):
        d = np.full(self.head + self.n * self.dp_stride + 16, fill, dtype=np.uint8)
        dn = np.full(self.n * 12 + 16, fill, dtype=np.uint8) if self.layout == "separate" else None
        return d, dn

    def strides(self):
        return (self.sp_stride, self.dp_stride, self.sn_stride, self.dn_stride, self.bw_stride, self.bi_stride)

    def pointers(self, dst, dst_n):
        """(srcPos, dstPos, srcNorm, dstNorm, weight, index) integer addresses."""
        sp = self.src.ctypes.data + self.head
        dp = dst.ctypes.data + self.head
        if self.layout in ("shared", "fat"):
            sn, dn = sp + 12, dp + 12
        elif self.layout == "separate":
            sn, dn = self.src_n.ctypes.data, dst_n.ctypes.data
        else:
            sn, dn = None, None
        bw = self.blend.ctypes.data + self.bw_off
        bi = (self.blend_i.ctypes.data if self.blend_i is not None else self.blend.ctypes.data) + self.bi_off
        return sp, dp, sn, dn, bw, bi

    def dst_views(self, dst, dst_n):
        """float32 [n,3] positions and normals extracted from the destination buffers."""
        dv = dst[self.head:self.head + self.n * self.dp_stride].reshape(self.n, self.dp_stride)
        pos = dv[:, 0:12].copy().view(np.float32).reshape(self.n, 3)
        if self.layout in ("shared", "fat"):
            nrm = dv[:, 12:24].copy().view(np.float32).reshape(self.n, 3)
        elif self.layout == "separate":
            nrm = dst_n[:self.n * 12].copy().view(np.float32).reshape(self.n, 3)
        else:
            nrm = None
        return pos, nrm


def edge_triangles(rng, ntri, nvert):
    """EdgeData::Triangle array (32-byte records, vertIndex[3] at byte 8) with random indices."""
    t = np.zeros((ntri, 8), dtype=np.uint32)
    t[:, 0] = rng.integers(0, 4, size=ntri)
    t[:, 1] = rng.integers(0, 4, size=ntri)
    t[:, 2:5] = rng.integers(0, nvert, size=(ntri, 3))
    t[:, 5:8] = rng.integers(0, nvert, size=(ntri, 3))
    return t


def bits(a):
    """float32 array -> uint32 bit patterns with -0.0 canonicalised to +0.0."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = a.view(np.uint32).copy()
    b[b == 0x80000000] = 0
    return b
