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


# ------------------------------------------------------------------ raster scenes ------

def perspective(fovy_deg, aspect, near, far):
    """OpenGL-style projection, depth range [-1,1] (TinyRenderSystem keeps OGRE's GL convention,
    include/OgreTinyRenderSystem.h:90-92)."""
    f = 1.0 / np.tan(np.radians(fovy_deg) / 2)
    m = np.zeros((4, 4), dtype=np.float64)
    m[0, 0] = f / aspect; m[1, 1] = f
    m[2, 2] = -(far + near) / (far - near); m[2, 3] = -2 * far * near / (far - near)
    m[3, 2] = -1
    return m


def look_at(eye, target, up=(0, 1, 0)):
    eye = np.asarray(eye, float); target = np.asarray(target, float); up = np.asarray(up, float)
    f = target - eye; f /= np.linalg.norm(f)
    s = np.cross(f, up); s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[:3, 3] = -m[:3, :3] @ eye
    return m


def viewport_matrix(w, h, left=0, top=0):
    """mVP of TinyRenderSystem::_setViewport for a flipping target (OgreTinyRenderSystem.cpp:257-258)."""
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, 0] = np.float32(w) / np.float32(2); m[0, 3] = np.float32(left) + np.float32(w) / np.float32(2)
    m[1, 1] = np.float32(h) / np.float32(2); m[1, 3] = np.float32(top) + np.float32(h) / np.float32(2)
    m[2, 2] = 0.5; m[2, 3] = 0.5
    m[3, 3] = 1
    return m


def random_texture(rng, w, h, zero_alpha=False):
    t = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    t[..., 3] = 255
    if zero_alpha:
        hole = rng.random((h, w)) < 0.3
        t[hole, 3] = 0
    return t


class DrawCase:
    """One _render call: interleaved vertex buffer (pos f3 @0, normal f3 @12, uv f2 @24, stride 32 —
    ogrehead's layout) + indices + state + uniforms."""

    def __init__(self, verts, indices=None, index_type=4, strip=False, texture=None, lighting=True,
                 depth_test=True, depth_write=True, blend_add=False, mvp=None, mvit=None, tex_matrix=None,
                 viewport=None, light_dir=(0.3, 0.5, 0.8), ambient=(0.2, 0.25, 0.3, 1.0),
                 has_uv=True, has_normals=True, state=None):
        # state: the opt-in "beyond Tiny" render state, a dict with any of depth_func, cull_mode (Ogre enum values),
        # scissor (l, t, r, b), alpha (func, ref), colours ((n, 4) uint8 r,g,b,a per vertex)
        self.state = dict(state) if state else None
        if self.state and "colours" in self.state:
            self.state["colours"] = np.ascontiguousarray(self.state["colours"], dtype=np.uint8)
            assert self.state["colours"].shape == (len(verts), 4)
        self.verts = np.ascontiguousarray(verts, dtype=np.float32)      # [n, 8]
        if indices is None:
            self.index_type, self.indices = 0, None
            self.draw_count = len(self.verts)
        else:
            self.index_type = index_type
            self.indices = np.ascontiguousarray(indices, dtype=np.uint16 if index_type == 2 else np.uint32)
            self.draw_count = len(self.indices)
        self.strip, self.texture = strip, texture
        self.lighting, self.depth_test, self.depth_write, self.blend_add = lighting, depth_test, depth_write, blend_add
        self.mvp = np.ascontiguousarray(mvp, dtype=np.float32)
        self.mvit = np.ascontiguousarray(np.eye(4) if mvit is None else mvit, dtype=np.float32)
        self.tex_matrix = np.ascontiguousarray(np.eye(4) if tex_matrix is None else tex_matrix, dtype=np.float32)
        self.viewport = np.ascontiguousarray(viewport, dtype=np.float32)
        l = np.asarray(light_dir, dtype=np.float64)
        self.light_dir = (l / np.linalg.norm(l)).astype(np.float32)
        self.ambient = np.asarray(ambient, dtype=np.float32)
        self.has_uv, self.has_normals = has_uv, has_normals

    def oracle_desc(self):
        import oracle
        d = oracle.OracleDraw()
        base = self.verts.ctypes.data
        d.positions, d.pos_stride = base, 32
        d.normals, d.normal_stride = (base + 12, 32) if self.has_normals else (None, 0)
        d.uvs, d.uv_stride = (base + 24, 32) if self.has_uv else (None, 0)
        d.indices = self.indices.ctypes.data if self.indices is not None else None
        d.index_type, d.draw_count, d.is_strip = self.index_type, self.draw_count, int(self.strip)
        d.depth_test, d.depth_write, d.blend_add = int(self.depth_test), int(self.depth_write), int(self.blend_add)
        d.lighting = int(self.lighting and self.has_normals)
        if self.texture is not None:
            d.tex_rgba, d.tex_w, d.tex_h = self.texture.ctypes.data, self.texture.shape[1], self.texture.shape[0]
        d.mvp[:] = self.mvp.reshape(16); d.tex_matrix[:] = self.tex_matrix.reshape(16)
        d.mvit[:] = self.mvit.reshape(16); d.viewport[:] = self.viewport.reshape(16)
        d.light_dir[:] = self.light_dir; d.ambient[:] = self.ambient
        return d

    def fill_state(self, st, colours_ptr):
        """Fill an OracleDrawState / ogre_b200.DrawState (same fields) from self.state; colours_ptr = address of the colours."""
        s = self.state
        st.enable = 0
        if "depth_func" in s:
            st.enable |= 1; st.depth_func = int(s["depth_func"])
        if "cull_mode" in s:
            st.enable |= 2; st.cull_mode = int(s["cull_mode"])
        if "scissor" in s:
            st.enable |= 4; st.scissor[:] = [int(v) for v in s["scissor"]]
        if "alpha" in s:
            st.enable |= 8; st.alpha_func, st.alpha_ref = int(s["alpha"][0]), int(s["alpha"][1])
        if "colours" in s:
            st.enable |= 16; st.colours, st.colour_stride = colours_ptr, 4
        return st

    def oracle_state(self):
        import oracle
        if not self.state:
            return None
        return self.fill_state(oracle.OracleDrawState(), self.state["colours"].ctypes.data if "colours" in self.state else None)


def random_soup(rng, ntri, spread=3.0, size=1.0, strip=False, indexed=True):
    """Random triangles around the origin; returns (verts[n,8], indices or None)."""
    if strip:
        nv = ntri + 2
        p = np.cumsum(rng.normal(scale=size * 0.4, size=(nv, 3)), axis=0) + rng.uniform(-spread, spread, size=3)
        idx = np.arange(nv)
    else:
        c = rng.uniform(-spread, spread, size=(ntri, 1, 3))
        p = (c + rng.normal(scale=size, size=(ntri, 3, 3))).reshape(-1, 3)
        nv = len(p)
        idx = np.arange(nv)
        if indexed:   # share some vertices between triangles
            share = rng.random(nv) < 0.3
            idx[share] = rng.integers(0, nv, size=share.sum())
    n = rng.normal(size=(nv, 3)); n /= np.linalg.norm(n, axis=1, keepdims=True)
    uv = rng.uniform(-1.5, 2.5, size=(nv, 2))
    verts = np.concatenate([p, n, uv], axis=1).astype(np.float32)
    return verts, (idx if indexed else None)


def make_scene(seed, w, h, ndraws=4, ntri=200, **kw):
    """A list of DrawCase with assorted state covering every Tiny code path."""
    rng = np.random.default_rng(seed)
    proj = perspective(50, w / h, 0.5, 100.0)
    view = look_at(rng.uniform(-2, 2, 3) + np.array([0, 0, 9.0]), rng.uniform(-1, 1, 3))
    vp = viewport_matrix(w, h)
    texs = [random_texture(rng, 37, 19), random_texture(rng, 64, 64, zero_alpha=True), None]
    cases = []
    for k in range(ndraws):
        strip = kw.get("strip", k % 4 == 3)
        index_type = kw.get("index_type", 2 if k % 2 else 4)
        verts, idx = random_soup(rng, ntri, size=kw.get("size", 1.0), strip=strip, indexed=kw.get("indexed", k % 3 != 2))
        if idx is not None and index_type == 2 and len(verts) > 32767:   # the reference reads 16-bit indices as SIGNED (UB >= 32768)
            index_type = 4
        model = np.eye(4); model[:3, 3] = rng.uniform(-1, 1, 3)
        mv = view @ model
        mvp = (proj @ mv).astype(np.float32)
        mvit = np.linalg.inv(mv).T.astype(np.float32)
        texm = np.eye(4, dtype=np.float32)
        if k % 2:
            texm[0, 0], texm[1, 1], texm[0, 3] = 5.0, 5.0, 0.25      # Ogre/Tusks-style "scale 0.2 0.2"
        cases.append(DrawCase(
            verts, idx, index_type=index_type, strip=strip,
            texture=kw.get("texture", texs[k % 3]), lighting=kw.get("lighting", k % 3 != 1),
            depth_test=kw.get("depth_test", k != 2 or ndraws < 3), depth_write=kw.get("depth_write", k % 5 != 4),
            blend_add=kw.get("blend_add", k % 4 == 1), mvp=mvp, mvit=mvit, tex_matrix=texm, viewport=vp,
            light_dir=rng.normal(size=3), ambient=np.append(rng.uniform(0, 0.5, 3), 1.0)))
    return cases


def degenerate_soup(rng, w, h):
    """Clip-space triangles [n,3,4] hitting Tiny's corner cases: slivers with |det| <= 1e-6,
    coincident vertices, shared edges, equal depth, w < 0, partly/fully off-screen, axis-aligned
    zero-area lines through integer pixels, negative depth."""
    tris = []

    def ndc(x, y, z=0.0, wv=1.0):
        return [(x / w * 2 - 1) * wv, (y / h * 2 - 1) * wv, z * wv, wv]
    tris += [[ndc(10, 10, .2), ndc(60, 10, .2), ndc(10, 60, .2)], [ndc(60, 10, .2), ndc(60, 60, .2), ndc(10, 60, .2)]]
    tris += [[ndc(10, 10, .2), ndc(10, 60, .2), ndc(60, 10, .2)]]
    tris += [[ndc(30, 30), ndc(30, 30), ndc(30, 30)]]
    tris += [[ndc(5, 40), ndc(50, 40), ndc(50, 40)], [ndc(70, 5), ndc(70, 5), ndc(70, 90)], [ndc(0, 0), ndc(20, 20), ndc(40, 40)]]
    for k in range(40):
        x, y = rng.uniform(0, w), rng.uniform(0, h)
        d = rng.normal(size=2) * rng.uniform(1, 80)
        e = rng.normal(size=2) * 1e-4
        tris += [[ndc(x, y, .5), ndc(x + d[0], y + d[1], .5), ndc(x + d[0] / 2 + e[0], y + d[1] / 2 + e[1], .5)]]
    tris += [[ndc(20, 20, .1, -1.0), ndc(90, 25, .1, -1.0), ndc(40, 80, .1, -1.0)],
             [ndc(20, 20, .1, 1.0), ndc(90, 25, .1, -0.5), ndc(40, 80, .1, 2.0)]]
    tris += [[ndc(-50, -50), ndc(-10, -50), ndc(-50, -10)], [ndc(-30, 20, .3), ndc(w + 40, 50, .3), ndc(40, h + 60, .3)],
             [ndc(-0.5, 10), ndc(-0.2, 30), ndc(-0.9, 50)]]
    tris += [[ndc(50, 50, -1.5), ndc(90, 50, -1.5), ndc(50, 90, -1.5)], [ndc(55, 55, 1.5), ndc(95, 55, 1.5), ndc(55, 95, 1.5)]]
    return np.array(tris, dtype=np.float64)


def clip_case(clip_tris, w, h, **kw):
    """DrawCases from clip-space triangles [n,3,4].  A vertex is stored as (xc, yc, wc) and drawn with an
    MVP that yields clip = (x, y, ratio*w, w); consecutive triangles with the same depth ratio zc/wc
    share a draw, so submission order is kept."""
    a = np.asarray(clip_tris, dtype=np.float64)
    ratios = np.round(a[:, 0, 2] / a[:, 0, 3], 6)
    spans, start = [], 0
    for i in range(1, len(a) + 1):
        if i == len(a) or ratios[i] != ratios[start]:
            spans.append((start, i)); start = i
    vp = viewport_matrix(w, h)
    rng = np.random.default_rng(11)
    cases = []
    for (s, e) in spans:
        t = a[s:e]
        nv = (e - s) * 3
        pos = np.stack([t[..., 0].reshape(-1), t[..., 1].reshape(-1), t[..., 3].reshape(-1)], axis=1)
        n = rng.normal(size=(nv, 3)); n /= np.linalg.norm(n, axis=1, keepdims=True)
        uv = rng.uniform(0, 1, size=(nv, 2))
        verts = np.concatenate([pos, n, uv], axis=1).astype(np.float32)
        mvp = np.zeros((4, 4), dtype=np.float32)
        mvp[0, 0] = 1; mvp[1, 1] = 1; mvp[2, 2] = ratios[s]; mvp[3, 2] = 1
        cases.append(DrawCase(verts, None, mvp=mvp, viewport=vp, **kw))
    return cases


def pose_case(seed, nverts=4000, frac=0.3, with_normals=True):
    """A pose (sparse vertex offsets, optionally normals: Pose::getVertexOffsets / getNormals, std::map order = ascending
    vertex index) and a pos[+normal] vertex buffer to blend it into."""
    rng = np.random.default_rng(seed)
    fpv = 6 if with_normals else 3
    buf = rng.uniform(-5, 5, size=(nverts, fpv)).astype(np.float32)
    special = np.array([0.0, -0.0, 1e-39, -1e-45, 3e38], np.float32)
    m = rng.random(buf.shape) < 0.02
    buf[m] = special[rng.integers(0, len(special), size=int(m.sum()))]
    k = max(1, int(nverts * frac))
    off_idx = np.sort(rng.choice(nverts, k, replace=False)).astype(np.uint32)
    off = rng.uniform(-1, 1, size=(k, 3)).astype(np.float32)
    off[rng.random(off.shape) < 0.05] = np.float32(-0.0)
    kn = max(1, k // 2) if with_normals else 0
    nrm_idx = np.sort(rng.choice(nverts, kn, replace=False)).astype(np.uint32) if kn else np.zeros(0, np.uint32)
    nrm = rng.uniform(-1, 1, size=(kn, 3)).astype(np.float32)
    return buf, off_idx, off, nrm_idx, nrm


def pose_entity_case(seed, nv, nposes):
    """Inputs of a whole pose-animated frame: base mesh pos+normal [nv,6], poses (flattened maps with a prefix table),
    offsets and (sub-unit-length) normals per entry, influences per pose (one of them 0)."""
    rng = np.random.default_rng(seed)
    base = rng.uniform(-5, 5, size=(nv, 6)).astype(np.float32)
    nn = rng.normal(size=(nv, 3))
    base[:, 3:] = (nn / np.linalg.norm(nn, axis=1, keepdims=True)).astype(np.float32)
    offs, idx, off, nrm = [0], [], [], []
    for p in range(nposes):
        k = max(1, int(nv * rng.uniform(0.05, 0.6)))
        idx.append(np.sort(rng.choice(nv, k, replace=False)).astype(np.uint32))
        off.append(rng.uniform(-1, 1, size=(k, 3)).astype(np.float32))
        n3 = rng.normal(size=(k, 3)); n3 /= np.linalg.norm(n3, axis=1, keepdims=True); n3 *= rng.uniform(0.2, 1.0, size=(k, 1))
        nrm.append(n3.astype(np.float32))
        offs.append(offs[-1] + k)
    infl = rng.uniform(0.1, 1.0, nposes).astype(np.float32)
    if nposes > 2:
        infl[1] = 0.0
    return base, np.array(offs, np.uint32), np.concatenate(idx), np.concatenate(off), np.concatenate(nrm), infl


def pose_entity_oracle(oracle, base, offs, idx, off, nrm, infl, with_normals):
    """initialisePoseVertexData -> softwareVertexPoseBlend per pose -> finalisePoseNormals with the CPU oracle."""
    o = base.copy()
    if with_normals:
        o[:, 3:] = 0
    for p in range(len(infl)):
        a, b = int(offs[p]), int(offs[p + 1])
        ii, oo, n3 = np.ascontiguousarray(idx[a:b]), np.ascontiguousarray(off[a:b]), np.ascontiguousarray(nrm[a:b])
        oracle.lib().oracle_pose_blend(float(infl[p]), ii.ctypes.data, oo.ctypes.data, b - a, o.ctypes.data, 6)
        if with_normals:
            oracle.lib().oracle_pose_blend(float(infl[p]), ii.ctypes.data, n3.ctypes.data, b - a, o.ctypes.data + 12, 6)
    if with_normals:
        oracle.lib().oracle_finalise_pose_normals(o.ctypes.data + 12, base.ctypes.data + 12, len(base), 6, 6)
    return o


# ---- scaled read-back (Image::scale FILTER_BILINEAR on byte images) ----
SCALE_CASES = [  # (seed, src h, src w, box or None, dst w, dst h)
    (0, 36, 64, None, 32, 18), (1, 36, 64, None, 129, 75), (2, 36, 64, (3, 5, 57, 34), 97, 33), (3, 20, 20, (10, 10, 11, 11), 8, 8),
    (4, 36, 64, None, 1, 1), (5, 1, 64, None, 64, 3), (6, 48, 1, None, 5, 48), (7, 50, 70, None, 70, 49), (8, 50, 70, None, 69, 50),
    (9, 7, 9, None, 200, 150), (10, 120, 160, None, 7, 5), (11, 33, 65, (1, 0, 65, 33), 64, 32),
]


def scale_image(seed, h, w, channels=3):
    """Seeded source image: smooth gradients plus noise, so both interpolation and rounding are exercised."""
    rng = np.random.default_rng(1000 + seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([(x * 255 // max(1, w - 1)), (y * 255 // max(1, h - 1)), ((x + y) * 7) % 256][:channels] +
                   [np.full((h, w), 255)] * max(0, channels - 3), axis=-1).astype(np.int64)
    img = (img + rng.integers(-40, 41, img.shape)).clip(0, 255)
    img[rng.random((h, w)) < 0.1] = rng.integers(0, 256, channels)
    return img.astype(np.uint8)


# destination formats of the 32-bit 8-bit-per-channel family: name -> (r_shift, g_shift, b_shift, fill) as PixelUtil lays them
# out on a little-endian host (OgrePixelFormatDescriptions.h)
PACKED32 = {"PF_A8R8G8B8": (16, 8, 0, 0xFF000000), "PF_A8B8G8R8": (0, 8, 16, 0xFF000000), "PF_B8G8R8A8": (8, 16, 24, 0x000000FF),
            "PF_R8G8B8A8": (24, 16, 8, 0x000000FF), "PF_X8R8G8B8": (16, 8, 0, 0xFF000000), "PF_X8B8G8R8": (0, 8, 16, 0xFF000000)}


def pack32(rgb, r_shift, g_shift, b_shift, fill):
    c = rgb.astype(np.uint32)
    return (c[..., 0] << r_shift) | (c[..., 1] << g_shift) | (c[..., 2] << b_shift) | np.uint32(fill)


# ---- opt-in "beyond Tiny" render state ----
def random_state(rng, nverts, w, h, what=("depth_func", "cull_mode", "scissor", "alpha", "colours")):
    """A random ogrecu_draw_state / oracle_draw_state dict (see DrawCase) using the fields named in `what`."""
    s = {}
    if "depth_func" in what:
        s["depth_func"] = int(rng.integers(0, 8))
    if "cull_mode" in what:
        s["cull_mode"] = int(rng.integers(1, 4))
    if "scissor" in what:
        l, r = sorted(int(v) for v in rng.integers(0, w + 1, 2)); t, b = sorted(int(v) for v in rng.integers(0, h + 1, 2))
        s["scissor"] = (l, t, r, b)
    if "alpha" in what:
        s["alpha"] = (int(rng.integers(0, 8)), int(rng.integers(0, 256)))
    if "colours" in what:
        s["colours"] = rng.integers(0, 256, (nverts, 4), dtype=np.uint8)
    return s


def with_states(cases, seed, w, h, what=None):
    """The same draws, each with a random beyond-Tiny state made of a random subset of the fields (or of `what`)."""
    rng = np.random.default_rng(seed)
    names = ("depth_func", "cull_mode", "scissor", "alpha", "colours")
    for c in cases:
        pick = what if what is not None else [n for n in names if rng.random() < 0.5]
        c.state = random_state(rng, len(c.verts), w, h, pick) or None
    return cases
