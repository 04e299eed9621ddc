"""ogre_b200 — B200-native back end for OGRE's software-rendering hot path.

The product is libogrecu.so (hand-written sm_100a kernels behind the C ABI of
include/ogrecu.h) plus the C++ drop-in plugin in ogre_b200/plugin.  This Python package
is only the thin host-side mirror used by the tests and bench.py: it forwards to the C
ABI through ctypes and keeps the reference's names and argument meaning
(Ogre::OptimisedUtil, OgreMain/include/OgreOptimisedUtil.h:99-206;
TinyRenderSystem::_render, RenderSystems/Tiny/src/OgreTinyRenderSystem.cpp:349-414).
It never computes results itself and never touches oracle/.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import AnimJob, AnimationDesc, DrawDesc, DrawState, OgrecuError, RasterStats, SkeletonDesc, SkinJob, check, lib

__all__ = ["Context", "DeviceBuffer", "OptimisedUtil", "Skeleton", "EdgeList", "Texture", "RenderTarget", "HardwareBuffer", "DrawDesc", "DrawState", "SkinJob", "AnimJob",
           "OgrecuError", "addr", "pinned_empty"]

MAP_READ, MAP_WRITE, MAP_DISCARD = 1, 2, 4
CLEAR_COLOUR, CLEAR_DEPTH = 1, 2
PRIM_TRIANGLE_LIST, PRIM_TRIANGLE_STRIP = 4, 5
INDEX_NONE, INDEX_U16, INDEX_U32 = 0, 2, 4
DRAW_DEPTH_TEST, DRAW_DEPTH_WRITE, DRAW_BLEND_ADD, DRAW_LIGHTING = 1, 2, 4, 8
# opt-in "beyond Tiny" state (ogrecu_draw_state): Ogre::CompareFunction / Ogre::CullingMode values
EXT_DEPTH_FUNC, EXT_CULL_MODE, EXT_SCISSOR, EXT_ALPHA_REJECT, EXT_VERTEX_COLOUR = 1, 2, 4, 8, 16
CMPF_ALWAYS_FAIL, CMPF_ALWAYS_PASS, CMPF_LESS, CMPF_LESS_EQUAL, CMPF_EQUAL, CMPF_NOT_EQUAL, CMPF_GREATER_EQUAL, CMPF_GREATER = range(8)
CULL_NONE, CULL_CLOCKWISE, CULL_ANTICLOCKWISE = 1, 2, 3


class _Pinned:
    """Owner of one ogrecu_host_alloc block; freed when the last numpy view goes away."""

    def __init__(self, nbytes):
        p = C.c_void_p()
        check(lib().ogrecu_host_alloc(int(nbytes), C.byref(p)))
        self.ptr, self.nbytes = p.value, int(nbytes)

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().ogrecu_host_free(self.ptr)
            self.ptr = None


def pinned_empty(shape, dtype=np.uint8):
    """numpy array over pinned host memory (ogrecu_host_alloc): read-backs into it are direct DMA."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    own = _Pinned(max(n, 1))
    buf = (C.c_uint8 * max(n, 1)).from_address(own.ptr)
    buf._owner = own
    arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    return arr


def addr(arr, byte_offset=0):
    """Address of a numpy array's first byte (+offset); None stays None (NULL)."""
    if arr is None:
        return None
    return arr.ctypes.data + int(byte_offset)


class DeviceBuffer:
    """Raw device allocation owned by a Context."""

    def __init__(self, ctx, nbytes):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(lib().ogrecu_malloc(ctx.handle, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def upload(self, arr, byte_offset=0):
        a = np.ascontiguousarray(arr)
        assert byte_offset + a.nbytes <= self.nbytes
        check(lib().ogrecu_memcpy_h2d(self.ctx.handle, self.ptr + byte_offset, a.ctypes.data, a.nbytes))
        self.ctx.sync()   # the source numpy array may be pageable and short-lived
        return self

    def download(self, dtype=np.uint8, count=None, byte_offset=0):
        dt = np.dtype(dtype)
        n = (self.nbytes - byte_offset) // dt.itemsize if count is None else int(count)
        out = np.empty(n, dtype=dt)
        check(lib().ogrecu_memcpy_d2h(self.ctx.handle, out.ctypes.data, self.ptr + byte_offset, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            check(lib().ogrecu_free(self.ctx.handle, self.ptr))
            self.ptr = None


class Context:
    """One CUDA device (ogrecu_ctx).  Raises OgrecuError when no device is usable."""

    def __init__(self, device=0, devices=None):
        """device: one CUDA device; devices=[d0, d1, ...]: a device group driven from this process (ogrecu_init_devices: d0 is the
        root, the others take raster passes of placed / split targets)."""
        h = C.c_void_p()
        if devices is not None:
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            check(lib().ogrecu_init_devices(len(devices), devs, C.byref(h)))
            self.device = int(devices[0])
            self.num_devices = len(devices)
        else:
            check(lib().ogrecu_init(int(device), C.byref(h)))
            self.device = int(device)
            self.num_devices = 1
        self.handle = h.value

    def close(self):
        if self.handle:
            check(lib().ogrecu_shutdown(self.handle))
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def sync(self):
        check(lib().ogrecu_sync(self.handle))

    def set_stream(self, cuda_stream):
        check(lib().ogrecu_set_stream(self.handle, cuda_stream))

    def set_config_option(self, name, value):
        check(lib().ogrecu_set_config_option(self.handle, name.encode(), int(value)))

    @property
    def launch_count(self):
        return int(lib().ogrecu_launch_count(self.handle))

    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def to_device(self, arr):
        a = np.ascontiguousarray(arr)
        return DeviceBuffer(self, max(a.nbytes, 1)).upload(a)


def _light4(light):
    a = np.ascontiguousarray(light, dtype=np.float32).reshape(-1)
    return a, a.ctypes.data_as(capi.c_float_p)


MAP_READ, MAP_WRITE, MAP_DISCARD = 1, 2, 4


class HardwareBuffer:
    """ogrecu_buf — the HardwareBuffer delegate's storage (OgreMain/include/OgreHardwareBuffer.h:181-186,279-330): a device
    block with a pinned host shadow.  lock() returns the shadow as a numpy byte view (valid for the buffer's life); OptimisedUtil
    calls made with addresses inside it run on the device copy."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, int(nbytes)
        h = C.c_void_p()
        check(lib().ogrecu_buf_create(ctx.handle, self.nbytes, C.byref(h)))
        self.handle = h.value

    def lock(self, mode=MAP_READ | MAP_WRITE, offset=0, length=None):
        length = self.nbytes - offset if length is None else length
        p = C.c_void_p()
        check(lib().ogrecu_buf_map(self.handle, offset, length, mode, C.byref(p)))
        return np.ctypeslib.as_array(C.cast(p.value, C.POINTER(C.c_uint8)), shape=(length,))

    def unlock(self):
        check(lib().ogrecu_buf_unmap(self.handle))

    def write(self, array, offset=0):
        a = np.ascontiguousarray(array).view(np.uint8).reshape(-1)
        self.lock(MAP_WRITE, offset, a.size)[:] = a
        self.unlock()

    def read(self, offset=0, length=None):
        out = self.lock(MAP_READ, offset, length).copy()
        self.unlock()
        return out

    @property
    def device_ptr(self):
        """address of the device copy (draw descriptions, *_dev entry points)"""
        p = C.c_void_p()
        check(lib().ogrecu_buf_device_ptr(self.handle, C.byref(p)))
        return p.value

    def shadow_address(self):
        a = self.lock(MAP_READ, 0, 0 if self.nbytes == 0 else 1).ctypes.data
        self.unlock()
        return a

    def destroy(self):
        if self.handle:
            check(lib().ogrecu_buf_destroy(self.handle))
            self.handle = None


class OptimisedUtil:
    """Host-pointer mirror of Ogre::OptimisedUtil (same method names and argument order as
    OgreMain/include/OgreOptimisedUtil.h:99-206).  Pointer arguments are integer addresses
    (use addr()); the work runs on the GPU through the *_host C-ABI entry points."""

    def __init__(self, ctx):
        self.ctx = ctx

    def softwareVertexSkinning(self, srcPosPtr, destPosPtr, srcNormPtr, destNormPtr, blendWeightPtr, blendIndexPtr,
                               blendMatrices, srcPosStride, destPosStride, srcNormStride, destNormStride,
                               blendWeightStride, blendIndexStride, numWeightsPerVertex, numVertices):
        # blendMatrices: sequence of float32[16] arrays (the reference's `const Affine3* const*`)
        mats = [np.ascontiguousarray(m, dtype=np.float32).reshape(16) for m in blendMatrices]
        ptrs = (C.c_void_p * max(len(mats), 1))(*[m.ctypes.data for m in mats])
        check(lib().ogrecu_software_vertex_skinning_host(
            self.ctx.handle, srcPosPtr, destPosPtr, srcNormPtr, destNormPtr, blendWeightPtr, blendIndexPtr, ptrs,
            srcPosStride, destPosStride, srcNormStride, destNormStride, blendWeightStride, blendIndexStride,
            numWeightsPerVertex, numVertices))

    def softwareVertexMorph(self, t, srcPos1, srcPos2, dstPos, pos1VSize, pos2VSize, dstVSize, numVertices,
                            morphNormals):
        check(lib().ogrecu_software_vertex_morph_host(self.ctx.handle, float(t), srcPos1, srcPos2, dstPos,
                                                      pos1VSize, pos2VSize, dstVSize, numVertices,
                                                      1 if morphNormals else 0))

    def concatenateAffineMatrices(self, baseMatrix, srcMatrices, dstMatrices, numMatrices):
        b = np.ascontiguousarray(baseMatrix, dtype=np.float32).reshape(16)
        check(lib().ogrecu_concatenate_affine_matrices_host(self.ctx.handle, b.ctypes.data_as(capi.c_float_p),
                                                            srcMatrices, dstMatrices, numMatrices))

    def calculateFaceNormals(self, positions, triangles, faceNormals, numTriangles, numVertices=0):
        check(lib().ogrecu_calculate_face_normals_host(self.ctx.handle, positions, numVertices, triangles,
                                                       faceNormals, numTriangles))

    def calculateLightFacing(self, lightPos, faceNormals, lightFacings, numFaces):
        keep, lp = _light4(lightPos)
        check(lib().ogrecu_calculate_light_facing_host(self.ctx.handle, lp, faceNormals, lightFacings, numFaces))

    def extrudeVertices(self, lightPos, extrudeDist, srcPositions, destPositions, numVertices):
        keep, lp = _light4(lightPos)
        check(lib().ogrecu_extrude_vertices_host(self.ctx.handle, lp, float(extrudeDist), srcPositions,
                                                 destPositions, numVertices))


class Skeleton:
    """Device-side skeleton (ogrecu_skel): bone-palette production for many instances, the work of
    SkeletonInstance::setAnimationState + Skeleton::_getBoneMatrices (OgreMain/src/OgreSkeleton.cpp:162-217,402-424).
    `d` is the flat description {num_bones, blend_mode, parent, bind, inherit, [bind_inverse], anims: [{length,
    interpolation, rotation_interpolation, track_bone, track_shortest, key_offset, keys}]}."""

    def __init__(self, ctx, d, use_given_bind_inverse=False):
        self.ctx, self.num_bones = ctx, int(d["num_bones"])
        keep = [np.ascontiguousarray(d["parent"], np.int32), np.ascontiguousarray(d["bind"], np.float32),
                np.ascontiguousarray(d["inherit"], np.uint8)]
        desc = SkeletonDesc(self.num_bones, int(d["blend_mode"]), *[a.ctypes.data for a in keep], None)
        if use_given_bind_inverse:
            bi = np.ascontiguousarray(d["bind_inverse"], np.float32)
            desc.bind_inverse = bi.ctypes.data
        h = C.c_void_p()
        check(lib().ogrecu_skel_create(ctx.handle, C.byref(desc), C.byref(h)))
        self.handle = h.value
        for a in d["anims"]:
            self.add_animation(a)

    def add_animation(self, a):
        arrs = [np.ascontiguousarray(a["track_bone"], np.int32), np.ascontiguousarray(a["track_shortest"], np.uint8),
                np.ascontiguousarray(a["key_offset"], np.uint32), np.ascontiguousarray(a["keys"], np.float32)]
        desc = AnimationDesc(float(a["length"]), int(a["interpolation"]), int(a["rotation_interpolation"]), len(arrs[0]),
                             *[x.ctypes.data for x in arrs])
        idx = C.c_uint32()
        check(lib().ogrecu_skel_add_animation(self.handle, C.byref(desc), C.byref(idx)))
        return idx.value

    def set_blend_mode(self, mode):
        check(lib().ogrecu_skel_set_blend_mode(self.handle, int(mode)))

    def bind_inverse(self):
        out = np.zeros((self.num_bones, 10), np.float32)
        check(lib().ogrecu_skel_get_bind_inverse(self.handle, out.ctypes.data))
        return out

    def palettes_dev(self, num_instances, num_states, anim_ptr, time_ptr, weight_ptr, out_ptr, stride_floats=None):
        check(lib().ogrecu_skel_palettes_dev(self.handle, int(num_instances), int(num_states), anim_ptr, time_ptr, weight_ptr,
                                             out_ptr, int(stride_floats or self.num_bones * 16)))

    def palettes(self, anim, time, weight, mask=None, manual_bones=None, manual=None):
        """Host convenience (tests): anim/time/weight [num_instances, num_states] -> palettes [num_instances, num_bones, 16].
        mask: blend masks [num_instances, num_states, num_bones] or [num_states, num_bones] (shared); manual_bones: handles of the
        manually controlled bones, manual: their local transforms [num_instances, k, 10] (ogrecu_skel_palettes_ex_dev)."""
        anim = np.ascontiguousarray(anim, np.int32); time = np.ascontiguousarray(time, np.float32)
        weight = np.ascontiguousarray(weight, np.float32)
        ni, ns = anim.shape
        da, dt, dw = self.ctx.to_device(anim), self.ctx.to_device(time), self.ctx.to_device(weight)
        out = self.ctx.alloc(ni * self.num_bones * 64)
        keep = [da, dt, dw, out]
        if mask is None and manual_bones is None:
            self.palettes_dev(ni, ns, da.ptr, dt.ptr, dw.ptr, out.ptr)
        else:
            from .capi import PaletteOptions
            opt = PaletteOptions()
            if mask is not None:
                m = np.ascontiguousarray(mask, np.float32)
                dm = self.ctx.to_device(m); keep.append(dm)
                opt.mask_dev = dm.ptr
                opt.mask_instance_stride_floats = ns * self.num_bones if m.ndim == 3 else 0
            if manual_bones is not None and len(manual_bones):
                hb = (C.c_uint32 * len(manual_bones))(*[int(v) for v in manual_bones])
                dx = self.ctx.to_device(np.ascontiguousarray(manual, np.float32)); keep.append(dx)
                opt.num_manual = len(manual_bones); opt.manual_bones = hb; opt.manual_dev = dx.ptr
            check(lib().ogrecu_skel_palettes_ex_dev(self.handle, ni, ns, da.ptr, dt.ptr, dw.ptr, C.byref(opt), out.ptr, self.num_bones * 16))
        res = out.download(np.float32).reshape(ni, self.num_bones, 16)
        for b in keep:
            b.free()
        return res

    def destroy(self):
        if self.handle:
            check(lib().ogrecu_skel_destroy(self.handle))
            self.handle = None


class EdgeList:
    """Device-resident EdgeData (ogrecu_edges) + ShadowCaster::generateShadowVolume on it
    (OgreMain/src/OgreShadowCaster.cpp:191-515).  `ed` = {groups[g,5], edges[e,5], tris[t,3]} (uint32)."""

    def __init__(self, ctx, ed):
        self.ctx = ctx
        self.groups = np.ascontiguousarray(ed["groups"], np.uint32)
        edges = np.ascontiguousarray(ed["edges"], np.uint32); tris = np.ascontiguousarray(ed["tris"], np.uint32)
        self.num_tris = len(tris)
        h = C.c_void_p()
        check(lib().ogrecu_edges_create(ctx.handle, len(self.groups), self.groups.ctypes.data, edges.ctypes.data, tris.ctypes.data,
                                        len(tris), C.byref(h)))
        self.handle = h.value

    def use_mcguire(self, bounds6, light4):
        b = np.ascontiguousarray(bounds6, np.float32); l = np.ascontiguousarray(light4, np.float32)
        return int(lib().ogrecu_shadow_use_mcguire(len(self.groups), b.ctypes.data, l.ctypes.data))

    def shadow_volume(self, facings_ptr, directional, use_mcguire, flags, separate_light_cap, index_type, indices_ptr, buffer_indices,
                      used_in):
        used, emitted = C.c_uint32(int(used_in)), C.c_uint32()
        ranges = np.zeros((len(self.groups), 4), np.uint32)
        check(lib().ogrecu_shadow_volume_dev(self.handle, facings_ptr, int(directional), int(use_mcguire), int(flags), int(separate_light_cap),
                                             int(index_type), indices_ptr, int(buffer_indices), C.byref(used), ranges.ctypes.data,
                                             C.byref(emitted)))
        return ranges, used.value, emitted.value

    def shadow_volumes_multi(self, facings_ptr, facing_stride, directional, use_mcguire, flags, separate_light_cap, index_type, indices_ptr,
                             region_indices):
        """several lights in one submission (ogrecu_shadow_volumes_multi_dev): returns (ranges[L, G, 4], result[L, 2]) after a sync"""
        n = len(flags)
        d = (C.c_int * n)(*[int(v) for v in directional]); m = (C.c_int * n)(*[int(v) for v in use_mcguire])
        f = (C.c_uint32 * n)(*[int(v) for v in flags])
        ranges = pinned_empty((n, len(self.groups), 4), np.uint32); result = pinned_empty((n, 2), np.uint32)
        check(lib().ogrecu_shadow_volumes_multi_dev(self.handle, n, facings_ptr, int(facing_stride), d, m, f, int(separate_light_cap), int(index_type),
                                                    indices_ptr, int(region_indices), ranges.ctypes.data, result.ctypes.data))
        self.ctx.sync()
        return ranges.copy(), result.copy()

    def destroy(self):
        if self.handle:
            check(lib().ogrecu_edges_destroy(self.handle))
            self.handle = None


class Texture:
    """RGBA8 texture (TinyTexture, RenderSystems/Tiny/src/OgreTinyTexture.cpp:30-57)."""

    def __init__(self, ctx, rgba):
        a = np.ascontiguousarray(rgba, dtype=np.uint8)
        assert a.ndim == 3 and a.shape[2] == 4
        self.ctx = ctx
        self.height, self.width = a.shape[0], a.shape[1]
        h = C.c_void_p()
        check(lib().ogrecu_tex_create(ctx.handle, self.width, self.height, C.byref(h)))
        self.handle = h.value
        check(lib().ogrecu_tex_upload(self.handle, a.ctypes.data, self.width * 4))

    def destroy(self):
        if self.handle:
            check(lib().ogrecu_tex_destroy(self.handle))
            self.handle = None


PLACE_SPLIT = -1


class RenderTarget:
    """RGB8 colour + f32 depth planes (TinyWindow + TinyDepthBuffer)."""

    def __init__(self, ctx, width, height, placement=None):
        """placement (device groups): None / 0 = on the root, k = on device k of the group, PLACE_SPLIT = split over all devices"""
        self.ctx, self.width, self.height = ctx, int(width), int(height)
        h = C.c_void_p()
        if placement is None:
            check(lib().ogrecu_rt_create(ctx.handle, self.width, self.height, C.byref(h)))
        else:
            check(lib().ogrecu_rt_create_placed(ctx.handle, self.width, self.height, int(placement), C.byref(h)))
        self.handle = h.value

    def destroy(self):
        if self.handle:
            check(lib().ogrecu_rt_destroy(self.handle))
            self.handle = None

    def clear(self, buffers=CLEAR_COLOUR | CLEAR_DEPTH, colour=(0, 0, 0, 0), depth=1.0):
        c = (C.c_float * 4)(*[float(v) for v in colour])
        check(lib().ogrecu_rt_clear(self.handle, buffers, c, float(depth)))

    def draw(self, desc, state=None):
        if state is None:
            check(lib().ogrecu_draw(self.handle, C.byref(desc)))
        else:
            check(lib().ogrecu_draw_with_state(self.handle, C.byref(desc), C.byref(state)))

    def flush(self):
        check(lib().ogrecu_flush(self.handle))

    def set_strip(self, y0, y1):
        check(lib().ogrecu_rt_set_strip(self.handle, int(y0), int(y1)))

    def set_bands(self, band_rows, phase, period):
        check(lib().ogrecu_rt_set_bands(self.handle, int(band_rows), int(phase), int(period)))

    def read_colour(self, box=None, out=None):
        """copyContentsToMemory: `out` may be a preallocated (ideally pinned, see pinned_empty) array."""
        l, t, r, b = box if box else (0, 0, self.width, self.height)
        if out is None:
            out = np.empty((b - t, r - l, 3), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.shape == (b - t, r - l, 3) and out.flags.c_contiguous
        check(lib().ogrecu_rt_read_colour(self.handle, l, t, r, b, out.ctypes.data, (r - l) * 3))
        return out

    def read_colour_packed32(self, r_shift, g_shift, b_shift, fill, box=None, out=None):
        """copyContentsToMemory into a 32-bit 8-bit-per-channel format, converted on the device (one uint32 per pixel)."""
        l, t, r, b = box if box else (0, 0, self.width, self.height)
        if out is None:
            out = np.empty((b - t, r - l), dtype=np.uint32)
        assert out.dtype == np.uint32 and out.shape == (b - t, r - l) and out.flags.c_contiguous
        check(lib().ogrecu_rt_read_colour_packed32(self.handle, l, t, r, b, out.ctypes.data, (r - l) * 4, r_shift, g_shift, b_shift, fill))
        return out

    def read_colour_scaled(self, dst_w, dst_h, box=None, packed=None, pad=0):
        """blitToMemory into a destination of another size: Image::scale(FILTER_BILINEAR)'s byte arithmetic on the device.
        packed = (r_shift, g_shift, b_shift, fill) returns uint32 words, else RGB8; rows are `pad` pixels wider than dst_w
        (the padding keeps 0xEE)."""
        l, t, r, b = box if box else (0, 0, self.width, self.height)
        if packed is None:
            out = np.full((dst_h, dst_w + pad, 3), 0xEE, dtype=np.uint8)
            check(lib().ogrecu_rt_read_colour_scaled(self.handle, l, t, r, b, dst_w, dst_h, out.ctypes.data, (dst_w + pad) * 3, 3, 0, 0, 0, 0))
        else:
            out = np.full((dst_h, dst_w + pad), 0xEEEEEEEE, dtype=np.uint32)
            check(lib().ogrecu_rt_read_colour_scaled(self.handle, l, t, r, b, dst_w, dst_h, out.ctypes.data, (dst_w + pad) * 4, 4, *packed))
        return out

    def read_depth(self, box=None, out=None):
        l, t, r, b = box if box else (0, 0, self.width, self.height)
        if out is None:
            out = np.empty((b - t, r - l), dtype=np.float32)
        assert out.dtype == np.float32 and out.shape == (b - t, r - l) and out.flags.c_contiguous
        check(lib().ogrecu_rt_read_depth(self.handle, l, t, r, b, out.ctypes.data, (r - l) * 4))
        return out

    def write_colour(self, rgb):
        a = np.ascontiguousarray(rgb, dtype=np.uint8)
        assert a.shape == (self.height, self.width, 3)
        check(lib().ogrecu_rt_write_colour(self.handle, a.ctypes.data, self.width * 3))

    def write_depth(self, depth):
        a = np.ascontiguousarray(depth, dtype=np.float32)
        assert a.shape == (self.height, self.width)
        check(lib().ogrecu_rt_write_depth(self.handle, a.ctypes.data, self.width * 4))

    def stats(self):
        s = RasterStats()
        check(lib().ogrecu_rt_last_stats(self.handle, C.byref(s)))
        return {"triangles_in": s.triangles_in, "triangles_binned": s.triangles_binned,
                "tile_pairs": s.tile_pairs, "tiles_active": s.tiles_active,
                "tiles_visibility": s.tiles_visibility, "tiles_ordered": s.tiles_ordered}
