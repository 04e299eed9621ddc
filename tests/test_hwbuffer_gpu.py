"""GPU: OptimisedUtil calls whose pointers lie inside HardwareBuffer shadows (the way OGRE calls them: lock the buffers, pass
the locked pointers, OgreMesh.cpp:1827-1896).  The streams are resolved to their device copies — nothing is staged except the
bone palette — and by default the call does not wait for the device ("optutil_sync" = 0: the next lock is the sync point).
Results must equal the oracle bit for bit in both modes, also for long runs of back-to-back calls that cycle the pinned
palette ring several times and for a morph whose output feeds a skinning call without the host in between."""
import numpy as np
import pytest

import ogre_b200
import oracle
from datagen import SkinCase

pytestmark = pytest.mark.gpu


class Resident:
    """A SkinCase whose source, blend and destination streams live in ogrecu_bufs."""

    def __init__(self, ctx, case):
        self.case = case
        dst, dstn = case.new_dst()
        assert dstn is None and case.blend_i is None      # 'shared' / 'posonly' layouts with an interleaved blend stream
        self.src = ogre_b200.HardwareBuffer(ctx, case.src.size)
        self.blend = ogre_b200.HardwareBuffer(ctx, case.blend.size)
        self.dst = ogre_b200.HardwareBuffer(ctx, dst.size)
        self.src.write(case.src); self.blend.write(case.blend); self.dst.write(dst)
        self.dst_init = dst

    def pointers(self):
        c = self.case
        sp = self.src.shadow_address() + c.head
        dp = self.dst.shadow_address() + c.head
        sn, dn = (sp + 12, dp + 12) if c.has_norm else (None, None)
        b = self.blend.shadow_address()
        return sp, dp, sn, dn, b + c.bw_off, b + c.bi_off

    def expected(self):
        c = self.case
        dst = self.dst_init.copy()
        sp, dp, sn, dn, bw, bi = c.pointers(dst, None)
        oracle.skin("sse", sp, dp, sn, dn, bw, bi, c.mats, c.strides(), c.nw, c.n)
        return dst

    def free(self):
        for b in (self.src, self.blend, self.dst):
            b.destroy()


@pytest.mark.parametrize("sync", [0, 1])
def test_resident_skinning_calls_back_to_back(gpu_ctx, sync):
    ou = ogre_b200.OptimisedUtil(gpu_ctx)
    gpu_ctx.set_config_option("optutil_sync", sync)
    try:
        cases = [Resident(gpu_ctx, SkinCase(seed=300 + k, n=[1, 17, 1000, 4099, 20000][k % 5], nw=1 + k % 4,
                                            layout="shared" if k % 3 else "posonly", head=[0, 8, 4, 12][k % 4])) for k in range(27)]
        for r in cases:             # 27 calls without looking at any result: the 8-slot palette ring wraps three times
            c = r.case
            # OGRE locks the destination, calls OptimisedUtil with the locked pointer, unlocks (OgreMesh.cpp:1827-1896)
            ptrs = r.pointers()
            r.dst.lock(ogre_b200.MAP_WRITE | ogre_b200.MAP_DISCARD)
            ou.softwareVertexSkinning(*ptrs, list(c.mats), *c.strides(), c.nw, c.n)
            r.dst.unlock()
        for k, r in enumerate(cases):
            assert np.array_equal(r.dst.read(), r.expected()), (sync, k)
    finally:
        gpu_ctx.set_config_option("optutil_sync", 0)
        for r in cases:
            r.free()


def test_resident_morph_feeds_skinning_without_the_host(gpu_ctx):
    # Entity::updateAnimation's chain: softwareVertexMorph into the entity's working copy, then softwareVertexSkinning from it
    ou = ogre_b200.OptimisedUtil(gpu_ctx)
    n = 5000
    case = SkinCase(seed=77, n=n, nw=4, layout="shared")
    rng = np.random.default_rng(5)
    key2 = case.src.copy()
    k2 = key2[:n * 24].view(np.float32).reshape(n, 6)
    k2[:, :3] += rng.normal(scale=0.3, size=(n, 3)).astype(np.float32)
    t = np.float32(0.37)
    b1 = ogre_b200.HardwareBuffer(gpu_ctx, case.src.size); b2 = ogre_b200.HardwareBuffer(gpu_ctx, key2.size)
    work = ogre_b200.HardwareBuffer(gpu_ctx, case.src.size); blend = ogre_b200.HardwareBuffer(gpu_ctx, case.blend.size)
    dst0, _ = case.new_dst()
    out = ogre_b200.HardwareBuffer(gpu_ctx, dst0.size)
    b1.write(case.src); b2.write(key2); blend.write(case.blend); out.write(dst0); work.write(np.zeros(case.src.size, np.uint8))
    a1, a2, aw, ab, ao = [b.shadow_address() for b in (b1, b2, work, blend, out)]
    ou.softwareVertexMorph(t, a1, a2, aw, 24, 24, 24, n, True)
    ou.softwareVertexSkinning(aw, ao, aw + 12, ao + 12, ab + case.bw_off, ab + case.bi_off, list(case.mats), *case.strides(), 4, n)
    # oracle chain on host copies
    morphed = np.zeros(case.src.size, np.uint8)
    oracle.lib().oracle_morph(t, case.src.ctypes.data, key2.ctypes.data, morphed.ctypes.data, 24, 24, 24, n, 1)
    want = dst0.copy()
    oracle.skin("sse", morphed.ctypes.data, want.ctypes.data, morphed.ctypes.data + 12, want.ctypes.data + 12,
                case.blend.ctypes.data + case.bw_off, case.blend.ctypes.data + case.bi_off, case.mats, case.strides(), 4, n)
    assert np.array_equal(work.read()[:n * 24], morphed[:n * 24])
    assert np.array_equal(out.read(), want)
    for b in (b1, b2, work, blend, out):
        b.destroy()


def test_partial_write_lock_after_a_pending_device_write_keeps_both(gpu_ctx):
    # the device result is still in flight when the host maps part of the buffer for writing: the map brings the device
    # write back first, the host's bytes land on top, nothing is lost
    ou = ogre_b200.OptimisedUtil(gpu_ctx)
    r = Resident(gpu_ctx, SkinCase(seed=9, n=3000, nw=3, layout="shared"))
    c = r.case
    ou.softwareVertexSkinning(*r.pointers(), list(c.mats), *c.strides(), c.nw, c.n)
    view = r.dst.lock(ogre_b200.MAP_WRITE, 240, 24)
    view[:] = 0xAB
    r.dst.unlock()
    want = r.expected()
    want[240:264] = 0xAB
    assert np.array_equal(r.dst.read(), want)
    r.free()


def test_profile_events_nest_and_reject_an_unmatched_end(gpu_ctx):
    """RenderSystem::beginProfileEvent / endProfileEvent / markProfileEvent (OgreRenderSystem.h:1104-1115) -> NVTX ranges: no tool
    attached here, so only the bookkeeping is visible — nesting works and an end without a begin is an error, not a crash."""
    from ogre_b200 import lib
    L, h = lib(), gpu_ctx.handle
    assert L.ogrecu_profile_begin(h, b"frame") == 0
    assert L.ogrecu_profile_begin(h, b"shadow pass") == 0
    assert L.ogrecu_profile_mark(h, b"marker") == 0
    assert L.ogrecu_profile_end(h) == 0
    assert L.ogrecu_profile_end(h) == 0
    assert L.ogrecu_profile_end(h) != 0
    assert b"without a matching begin" in L.ogrecu_last_error_string()
