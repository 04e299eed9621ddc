"""ctypes view of the C ABI declared in include/ogrecu.h (libogrecu.so).

Nothing here computes anything: it declares argument types and loads the in-tree shared
library.  If the library is missing or no CUDA device is usable the calls FAIL — there is
no CPU fallback anywhere in the product path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# OGRECU_LIB: load another build of the same library (kernel tuning experiments); still no fallback
LIB_PATH = os.environ.get("OGRECU_LIB") or os.path.join(_HERE, "libogrecu.so")

c_void_pp = C.POINTER(C.c_void_p)
c_float_p = C.POINTER(C.c_float)


class OgrecuError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("ogrecu status %d: %s" % (status, message))
        self.status = status


class SkinJob(C.Structure):
    """ogrecu_skin_job (include/ogrecu.h)"""
    _fields_ = [
        ("src_pos", C.c_void_p), ("dst_pos", C.c_void_p),
        ("src_norm", C.c_void_p), ("dst_norm", C.c_void_p),
        ("blend_weight", C.c_void_p), ("blend_index", C.c_void_p),
        ("palette", C.c_void_p), ("bone_map", C.c_void_p),
        ("src_pos_stride", C.c_uint32), ("dst_pos_stride", C.c_uint32),
        ("src_norm_stride", C.c_uint32), ("dst_norm_stride", C.c_uint32),
        ("blend_weight_stride", C.c_uint32), ("blend_index_stride", C.c_uint32),
        ("num_weights", C.c_uint32), ("num_matrices", C.c_uint32),
        ("num_vertices", C.c_uint32), ("reserved", C.c_uint32),
    ]


class AnimJob(C.Structure):
    """ogrecu_anim_job (include/ogrecu.h)"""
    _fields_ = [("skin", SkinJob), ("key2", C.c_void_p), ("morph_out", C.c_void_p), ("morph_instance_stride_bytes", C.c_size_t),
                ("dst_instance_stride_bytes", C.c_size_t)]


class DrawDesc(C.Structure):
    """ogrecu_draw_desc (include/ogrecu.h)"""
    _fields_ = [
        ("positions", C.c_void_p), ("pos_stride", C.c_uint32),
        ("uvs", C.c_void_p), ("uv_stride", C.c_uint32),
        ("normals", C.c_void_p), ("normal_stride", C.c_uint32),
        ("indices", C.c_void_p), ("index_type", C.c_uint32),
        ("draw_count", C.c_uint32), ("prim_type", C.c_uint32), ("flags", C.c_uint32), ("vertex_count", C.c_uint32),
        ("texture", C.c_void_p),
        ("mvp", C.c_float * 16), ("tex_matrix", C.c_float * 16), ("mvit", C.c_float * 16),
        ("viewport", C.c_float * 16), ("light_dir", C.c_float * 3), ("ambient", C.c_float * 4),
        ("instance_count", C.c_uint32), ("instance_vertex_count", C.c_uint32),
    ]


class DrawState(C.Structure):
    """ogrecu_draw_state (include/ogrecu.h): the opt-in "beyond Tiny" render state"""
    _fields_ = [("enable", C.c_uint32), ("depth_func", C.c_uint32), ("cull_mode", C.c_uint32), ("alpha_func", C.c_uint32),
                ("alpha_ref", C.c_uint32), ("scissor", C.c_uint32 * 4), ("colour_stride", C.c_uint32), ("colours", C.c_void_p)]


class SkeletonDesc(C.Structure):
    """ogrecu_skeleton_desc (include/ogrecu.h)"""
    _fields_ = [("num_bones", C.c_uint32), ("blend_mode", C.c_uint32), ("parent", C.c_void_p), ("bind", C.c_void_p),
                ("inherit", C.c_void_p), ("bind_inverse", C.c_void_p)]


class PaletteOptions(C.Structure):
    """ogrecu_palette_options"""
    _fields_ = [("mask_dev", C.c_void_p), ("mask_instance_stride_floats", C.c_size_t), ("num_manual", C.c_uint32),
                ("manual_bones", C.POINTER(C.c_uint32)), ("manual_dev", C.c_void_p)]


class AnimationDesc(C.Structure):
    """ogrecu_animation_desc (include/ogrecu.h)"""
    _fields_ = [("length", C.c_float), ("interpolation", C.c_uint32), ("rotation_interpolation", C.c_uint32),
                ("num_tracks", C.c_uint32), ("track_bone", C.c_void_p), ("track_shortest", C.c_void_p),
                ("key_offset", C.c_void_p), ("keys", C.c_void_p)]


class RasterStats(C.Structure):
    _fields_ = [("triangles_in", C.c_uint64), ("triangles_binned", C.c_uint64),
                ("tile_pairs", C.c_uint64), ("tiles_active", C.c_uint64),
                ("tiles_visibility", C.c_uint64), ("tiles_ordered", C.c_uint64)]


# name -> (restype, argtypes); every symbol include/ogrecu.h declares
SIGNATURES = {
    "ogrecu_abi_version": (C.c_int, []),
    "ogrecu_last_error_string": (C.c_char_p, []),
    "ogrecu_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "ogrecu_init": (C.c_int, [C.c_int, c_void_pp]),
    "ogrecu_shutdown": (C.c_int, [C.c_void_p]),
    "ogrecu_init_devices": (C.c_int, [C.c_int, C.POINTER(C.c_int), c_void_pp]),
    "ogrecu_num_devices": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "ogrecu_rt_create_placed": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, c_void_pp]),
    "ogrecu_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "ogrecu_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ogrecu_get_stream": (C.c_void_p, [C.c_void_p]),
    "ogrecu_sync": (C.c_int, [C.c_void_p]),
    "ogrecu_launch_count": (C.c_uint64, [C.c_void_p]),
    "ogrecu_profile_begin": (C.c_int, [C.c_void_p, C.c_char_p]),
    "ogrecu_profile_end": (C.c_int, [C.c_void_p]),
    "ogrecu_profile_mark": (C.c_int, [C.c_void_p, C.c_char_p]),
    "ogrecu_malloc": (C.c_int, [C.c_void_p, C.c_size_t, c_void_pp]),
    "ogrecu_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ogrecu_host_alloc": (C.c_int, [C.c_size_t, c_void_pp]),
    "ogrecu_host_free": (C.c_int, [C.c_void_p]),
    "ogrecu_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_memcpy_d2h_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_memcpy_h2d_on": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_memcpy_d2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_buf_create": (C.c_int, [C.c_void_p, C.c_size_t, c_void_pp]),
    "ogrecu_buf_destroy": (C.c_int, [C.c_void_p]),
    "ogrecu_buf_map": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, c_void_pp]),
    "ogrecu_buf_unmap": (C.c_int, [C.c_void_p]),
    "ogrecu_buf_device_ptr": (C.c_int, [C.c_void_p, c_void_pp]),
    "ogrecu_buf_size": (C.c_size_t, [C.c_void_p]),
    "ogrecu_buf_mark_device_dirty": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "ogrecu_buf_resolve": (C.c_int, [C.c_void_p, C.c_void_p, c_void_pp, C.POINTER(C.c_size_t)]),
    "ogrecu_skin_dev": (C.c_int, [C.c_void_p, C.POINTER(SkinJob), C.c_uint32]),
    "ogrecu_anim_instanced_multi_dev": (C.c_int, [C.c_void_p, C.POINTER(AnimJob), C.c_uint32, C.c_void_p, C.c_uint32, C.c_size_t]),
    "ogrecu_skin_plan_create": (C.c_int, [C.c_void_p, C.POINTER(SkinJob), c_void_pp]),
    "ogrecu_skin_plan_destroy": (C.c_int, [C.c_void_p]),
    "ogrecu_skin_plan_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32)]),
    "ogrecu_anim_frame_dev": (C.c_int, [C.c_void_p, C.POINTER(AnimJob), C.POINTER(C.c_void_p), C.c_uint32, C.c_void_p, C.c_uint32, C.c_size_t]),
    "ogrecu_skel_create": (C.c_int, [C.c_void_p, C.POINTER(SkeletonDesc), c_void_pp]),
    "ogrecu_skel_destroy": (C.c_int, [C.c_void_p]),
    "ogrecu_skel_add_animation": (C.c_int, [C.c_void_p, C.POINTER(AnimationDesc), C.POINTER(C.c_uint32)]),
    "ogrecu_skel_set_blend_mode": (C.c_int, [C.c_void_p, C.c_uint32]),
    "ogrecu_skel_get_bind_inverse": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ogrecu_skel_palettes_dev": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_size_t]),
    "ogrecu_skel_palettes_ex_dev": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PaletteOptions),
                                              C.c_void_p, C.c_size_t]),
    "ogrecu_pose_blend_dev": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]),
    "ogrecu_pose_blend_instanced_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32,
                                                  C.c_size_t]),
    "ogrecu_pose_normals_reset_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t]),
    "ogrecu_pose_normals_finalise_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t]),
    "ogrecu_edges_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, c_void_pp]),
    "ogrecu_edges_destroy": (C.c_int, [C.c_void_p]),
    "ogrecu_shadow_use_mcguire": (C.c_int, [C.c_uint32, C.c_void_p, C.c_void_p]),
    "ogrecu_shadow_volumes_multi_dev": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                                  C.POINTER(C.c_uint32), C.c_int, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "ogrecu_shadow_volume_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_uint32,
                                           C.POINTER(C.c_uint32), C.c_void_p, C.POINTER(C.c_uint32)]),
    "ogrecu_skin_instanced_dev": (C.c_int, [C.c_void_p, C.POINTER(SkinJob), C.c_uint32, C.c_size_t, C.c_size_t, C.c_size_t]),
    "ogrecu_anim_instanced_dev": (C.c_int, [C.c_void_p, C.POINTER(SkinJob), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                            C.c_uint32, C.c_size_t, C.c_size_t]),
    "ogrecu_set_config_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "ogrecu_morph_instanced_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                             C.c_size_t, C.c_size_t, C.c_int, C.c_uint32, C.c_size_t]),
    "ogrecu_software_vertex_skinning_host": (C.c_int, [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.POINTER(C.c_void_p), C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
        C.c_size_t, C.c_size_t]),
    "ogrecu_optutil_flush": (C.c_int, [C.c_void_p]),
    "ogrecu_optutil_batch_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "ogrecu_morph_dev": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                   C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]),
    "ogrecu_software_vertex_morph_host": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]),
    "ogrecu_concat_affine_dev": (C.c_int, [C.c_void_p, c_float_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_concat_affine_instanced_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t]),
    "ogrecu_concatenate_affine_matrices_host": (C.c_int, [C.c_void_p, c_float_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_face_normals_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_calculate_face_normals_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_light_facing_dev": (C.c_int, [C.c_void_p, c_float_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_light_facing_multi_dev": (C.c_int, [C.c_void_p, c_float_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                                C.c_size_t, C.c_size_t]),
    "ogrecu_calculate_light_facing_host": (C.c_int, [C.c_void_p, c_float_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_extrude_dev": (C.c_int, [C.c_void_p, c_float_p, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_extrude_multi_dev": (C.c_int, [C.c_void_p, c_float_p, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "ogrecu_extrude_vertices_host": (C.c_int, [C.c_void_p, c_float_p, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_tex_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, c_void_pp]),
    "ogrecu_tex_destroy": (C.c_int, [C.c_void_p]),
    "ogrecu_tex_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_tex_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_tex_copy_from_rt": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ogrecu_tex_device_ptr": (C.c_int, [C.c_void_p, c_void_pp]),
    "ogrecu_rt_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, c_void_pp]),
    "ogrecu_rt_destroy": (C.c_int, [C.c_void_p]),
    "ogrecu_rt_size": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "ogrecu_rt_clear": (C.c_int, [C.c_void_p, C.c_uint, c_float_p, C.c_float]),
    "ogrecu_draw": (C.c_int, [C.c_void_p, C.POINTER(DrawDesc)]),
    "ogrecu_draw_with_state": (C.c_int, [C.c_void_p, C.POINTER(DrawDesc), C.POINTER(DrawState)]),
    "ogrecu_draw_refs": (C.c_int, [C.c_void_p, C.POINTER(DrawDesc), C.c_void_p, C.POINTER(C.c_void_p), C.c_uint32]),
    "ogrecu_flush": (C.c_int, [C.c_void_p]),
    "ogrecu_draw_views": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32, C.POINTER(DrawDesc), c_float_p, c_float_p, c_float_p]),
    "ogrecu_rt_set_strip": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "ogrecu_rt_set_bands": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "ogrecu_composite_depth_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_size_t, c_float_p,
                                             C.c_float]),
    "ogrecu_rt_read_colour": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]),
    "ogrecu_rt_read_colour_packed32": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.c_uint32,
                                                 C.c_uint32, C.c_uint32, C.c_uint32]),
    "ogrecu_rt_read_colour_scaled": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                               C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "ogrecu_rt_read_depth": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]),
    "ogrecu_rt_write_colour": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_rt_write_depth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ogrecu_rt_colour_device_ptr": (C.c_int, [C.c_void_p, c_void_pp]),
    "ogrecu_rt_depth_device_ptr": (C.c_int, [C.c_void_p, c_void_pp]),
    "ogrecu_rt_last_stats": (C.c_int, [C.c_void_p, C.POINTER(RasterStats)]),
}

_lib = None


def lib():
    """Load libogrecu.so (in-tree).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OgrecuError(-1, "libogrecu.so is not built (%s); run __graft_entry__.build()" % LIB_PATH)
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)   # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status):
    if status != 0:
        msg = lib().ogrecu_last_error_string()
        raise OgrecuError(status, msg.decode() if msg else "")
    return status
