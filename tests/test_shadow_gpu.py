"""GPU parity of the device shadow-volume index generation (ogrecu_shadow_volume_dev) against the
index buffers the real ShadowCaster::generateShadowVolume wrote (tests/golden/*_shadow.npz) and against the CPU oracle on
a large synthetic mesh.  Integer work: identical bytes."""
import numpy as np
import pytest

import oracle
import ogre_b200
from goldenlib import load_shadow

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ogre_b200.Context(0)
    yield c
    c.close()


def run_gpu(ctx, el, facings, directional, mc, flags, separate, index_type, capacity, used_in):
    dt = np.uint16 if index_type == 2 else np.uint32
    fill = 0xCDCD if index_type == 2 else 0xCDCDCDCD
    dfac = ctx.to_device(np.ascontiguousarray(facings, np.int8))
    didx = ctx.to_device(np.full(capacity, fill, dt))
    ranges, used, emitted = el.shadow_volume(dfac.ptr, directional, mc, flags, separate, index_type, didx.ptr, capacity, used_in)
    idx = didx.download(dt)
    dfac.free(); didx.free()
    return idx, ranges, used, emitted


@pytest.mark.parametrize("name", ["jaiqua_shadow.npz", "sinbad_shadow.npz"])
def test_golden_cases_identical(ctx, name):
    ed, cases = load_shadow(name)
    el = ogre_b200.EdgeList(ctx, ed)
    for k, c in enumerate(cases):
        directional = c["light"][3] == 0
        mc = el.use_mcguire(c["bounds"], c["light"])
        assert mc == c["mcguire"], (name, k)
        idx, ranges, used, emitted = run_gpu(ctx, el, c["facings"], directional, mc, c["flags"], c["separate"], 2, c["capacity"], c["used_in"])
        assert (emitted, used) == (c["emitted"], c["used_out"]), (name, k)
        assert np.array_equal(idx[used - emitted:used], c["indices"]), (name, k)
        assert (idx[:used - emitted] == 0xCDCD).all() and (idx[used:] == 0xCDCD).all(), "wrote outside its range"
        cols = 4 if (c["separate"] and c["flags"] & 1) else 2
        assert np.array_equal(ranges[:, :cols], c["ranges"][:, :cols]), (name, k)
        # 32-bit indices: the same values (these meshes stay below the 16-bit limit)
        idx32, r32, u32, e32 = run_gpu(ctx, el, c["facings"], directional, mc, c["flags"], c["separate"], 4, c["capacity"], c["used_in"])
        assert np.array_equal(idx32[u32 - e32:u32], c["indices"].astype(np.uint32)) and u32 == used
    el.destroy()


def torus_edges(nu):
    """Closed torus grid: V = nu*nu, T = 2*nu*nu, E = 3*nu*nu manifold edges, one edge group."""
    i, j = np.meshgrid(np.arange(nu), np.arange(nu), indexing="ij")
    a = i * nu + j; b = ((i + 1) % nu) * nu + j; c = i * nu + (j + 1) % nu; d = ((i + 1) % nu) * nu + (j + 1) % nu
    tris = np.stack([np.stack([a, b, c], -1), np.stack([b, d, c], -1)], 2).reshape(-1, 3).astype(np.uint32)
    nt = len(tris)
    first = {}
    edges = []
    for t in range(nt):
        for k in range(3):
            v0, v1 = int(tris[t, k]), int(tris[t, (k + 1) % 3])
            key = (v1, v0)
            if key in first:
                e = first.pop(key)
                edges[e][1] = t
                edges[e][4] = 0
            else:
                first[(v0, v1)] = len(edges)
                edges.append([t, 0, v0, v1, 1])
    edges = np.array(edges, np.uint32)
    groups = np.array([[0, len(edges), 0, nt, nu * nu]], np.uint32)
    return {"groups": groups, "edges": edges, "tris": tris}


def test_large_mesh_32_bit_indices_equal_the_oracle(ctx):
    """180 000 triangles, 90 000 vertices: above the reference's 16-bit limit (v + vertex_count >= 65 536), so the oracle's
    16-bit output is compared with the low halves and the 32-bit output with the untruncated values recomputed here."""
    ed = torus_edges(300)
    nt = len(ed["tris"])
    rng = np.random.default_rng(3)
    facings = (rng.random(nt) < 0.5).astype(np.int8)
    el = ogre_b200.EdgeList(ctx, ed)
    for flags, mc, directional, sep in [(3, 0, False, False), (7, 1, True, True), (2, 1, False, False), (1, 0, True, True), (0, 0, False, False)]:
        light = np.array([0, 1, 0, 0 if directional else 1], np.float32)
        out = np.full(3000000, 0xCDCD, np.uint16)
        ranges = np.zeros((1, 4), np.uint32)
        import ctypes as C
        used = C.c_uint32(12)
        n = oracle.lib().oracle_shadow_volume(1, ed["groups"].ctypes.data, ed["edges"].ctypes.data, ed["tris"].ctypes.data, facings.ctypes.data,
                                              int(directional), mc, flags, int(sep), out.ctypes.data, len(out), C.addressof(used), ranges.ctypes.data)
        idx16, r16, u16, e16 = run_gpu(ctx, el, facings, directional, mc, flags, sep, 2, len(out), 12)
        assert (e16, u16) == (n, used.value)
        assert np.array_equal(idx16, out)
        assert np.array_equal(r16[:, :2], ranges[:, :2])
        idx32, r32, u32, e32 = run_gpu(ctx, el, facings, directional, mc, flags, sep, 4, len(out), 12)
        assert np.array_equal((idx32[12:u32] & 0xffff).astype(np.uint16), out[12:used.value])
        assert idx32[12:u32].max() < 2 * 90000 and (idx32[12:u32] >= 65536).any()
    el.destroy()


def test_buffer_too_small_is_an_error(ctx):
    ed, cases = load_shadow("jaiqua_shadow.npz")
    el = ogre_b200.EdgeList(ctx, ed)
    c = cases[1]
    with pytest.raises(ogre_b200.OgrecuError):
        run_gpu(ctx, el, c["facings"], False, 0, 3, False, 2, 100, 0)
    el.destroy()


def test_multi_light_call_equals_single_light_calls(ctx):
    """ogrecu_shadow_volumes_multi_dev: up to 8 lights, one submission, no host round trip — every light's region holds the bytes
    the single-light call writes, the ranges are the single call's shifted to the region, and a light whose indices do not fit
    reports so and writes nothing.  Golden edge lists (several groups, McGuire fans) and the large torus."""
    combos = [(3, 0, False), (7, 1, True), (2, 1, False), (1, 0, True), (0, 0, False), (3, 1, False), (6, 0, True), (5, 1, True)]
    for name in ("sinbad_shadow.npz", "jaiqua_shadow.npz", "torus"):
        if name == "torus":
            ed = torus_edges(120)
            nt = len(ed["tris"])
            rng = np.random.default_rng(11)
            facing_sets = [(rng.random(nt) < p).astype(np.int8) for p in (0.5, 0.1, 0.9, 0.5, 0.0, 1.0, 0.3, 0.7)]
        else:
            ed, cases = load_shadow(name)
            nt = len(ed["tris"])
            facing_sets = [np.ascontiguousarray(cases[k % len(cases)]["facings"], np.int8) for k in range(8)]
        el = ogre_b200.EdgeList(ctx, ed)
        for index_type, sep in ((2, False), (4, True)):
            dt = np.uint16 if index_type == 2 else np.uint32
            fill = 0xCDCD if index_type == 2 else 0xCDCDCDCD
            singles = []
            for (flags, mc, directional), fac in zip(combos, facing_sets):
                idx, ranges, used, emitted = run_gpu(ctx, el, fac, directional, mc, flags, sep, index_type, 9 * nt + 64, 0)
                singles.append((idx[:emitted], ranges, emitted))
            region = max(e for _, _, e in singles) + 5
            stride = (nt + 15) // 16 * 16
            fac_all = np.zeros((8, stride), np.int8)
            for l, fac in enumerate(facing_sets):
                fac_all[l, :nt] = fac
            dfac = ctx.to_device(fac_all)
            didx = ctx.to_device(np.full(8 * region, fill, dt))
            ranges, result = el.shadow_volumes_multi(dfac.ptr, stride, [c[2] for c in combos], [c[1] for c in combos], [c[0] for c in combos],
                                                    sep, index_type, didx.ptr, region)
            got = didx.download(dt).reshape(8, region)
            for l, (ref_idx, ref_ranges, emitted) in enumerate(singles):
                assert tuple(result[l]) == (emitted, 0), (name, l)
                assert np.array_equal(got[l, :emitted], ref_idx), (name, index_type, l)
                assert (got[l, emitted:] == fill).all(), "wrote past its indices"
                shifted = ref_ranges.copy().astype(np.int64)
                shifted[:, 0] += l * region
                if sep and (combos[l][0] & 1):
                    shifted[:, 2] += l * region
                cols = 4 if (sep and combos[l][0] & 1) else 2
                assert np.array_equal(ranges[l][:, :cols].astype(np.int64), shifted[:, :cols]), (name, l)
            # regions too small for the biggest light: that light reports and writes nothing, the others are unaffected
            small = max(e for _, _, e in singles) - 1
            didx2 = ctx.to_device(np.full(8 * small, fill, dt))
            _, result2 = el.shadow_volumes_multi(dfac.ptr, stride, [c[2] for c in combos], [c[1] for c in combos], [c[0] for c in combos],
                                                 sep, index_type, didx2.ptr, small)
            got2 = didx2.download(dt).reshape(8, small)
            for l, (ref_idx, _, emitted) in enumerate(singles):
                if emitted > small:
                    assert tuple(result2[l]) == (emitted, 1) and (got2[l] == fill).all(), (name, l)
                else:
                    assert tuple(result2[l]) == (emitted, 0) and np.array_equal(got2[l, :emitted], ref_idx), (name, l)
            dfac.free(); didx.free(); didx2.free()
        el.destroy()
