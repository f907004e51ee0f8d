"""ctypes view of include/rtb.h and include/rtw_host.h.

The product path is the C ABI in ``_lib/librtb.so`` (CUDA, sm_100a) plus the host-side mirror of the
reference's scene/camera API in ``_lib/librtw_host.so``.  There is no Python or CPU fallback: if the
libraries are missing this module raises, and every compute entry point returns RTB_ERR_NO_DEVICE
when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RTB_LIB_DIR: development knob (tools/ experiments with variant builds, see the Makefile); unset everywhere else.
LIB_DIR = os.environ.get("RTB_LIB_DIR") or os.path.join(_HERE, "_lib")

RTB_ABI_VERSION = 1

# RtbStatus
RTB_OK = 0
RTB_ERR_INVALID_ARGUMENT = -1
RTB_ERR_CUDA = -2
RTB_ERR_NO_DEVICE = -3
RTB_ERR_OUT_OF_MEMORY = -4
RTB_ERR_CANCELLED = -5
RTB_ERR_UNSUPPORTED = -6

RTB_HITTABLE_SPHERE, RTB_HITTABLE_QUAD, RTB_HITTABLE_BOX, RTB_HITTABLE_CONSTANT_MEDIUM = 0, 1, 2, 3
RTB_HITTABLE_TRANSLATE, RTB_HITTABLE_ROTATE_Y, RTB_HITTABLE_LIST, RTB_HITTABLE_MEDIUM_OF = 4, 5, 6, 7
RTB_MAT_LAMBERTIAN, RTB_MAT_METAL, RTB_MAT_DIELECTRIC, RTB_MAT_DIFFUSE_LIGHT, RTB_MAT_ISOTROPIC = range(5)
RTB_TEX_SOLID, RTB_TEX_CHECKER, RTB_TEX_IMAGE, RTB_TEX_NOISE = range(4)
RTB_BACKGROUND_SOLID, RTB_BACKGROUND_SKY = 0, 1
RTB_INTEGRATOR_MEGAKERNEL, RTB_INTEGRATOR_WAVEFRONT = 0, 1
RTB_TRAVERSAL_REFERENCE, RTB_TRAVERSAL_ORDERED, RTB_TRAVERSAL_SAH, RTB_TRAVERSAL_SAH16 = 0, 1, 2, 3
RTB_FLAG_COUNT_WORK = 1

RTW_SCENE_BOOK1, RTW_SCENE_EARTH, RTW_SCENE_TWO_SPHERES, RTW_SCENE_TWO_PERLIN, RTW_SCENE_TEXTURED, \
    RTW_SCENE_RANDOM_SPHERES, RTW_SCENE_QUADS, RTW_SCENE_SIMPLE_LIGHT, RTW_SCENE_CORNELL_BOX, RTW_SCENE_CORNELL_SMOKE = range(10)
RTW_BOOK1_CHECKER_GROUND, RTW_BOOK1_EARTH_SPHERE, RTW_BOOK1_STATIC_SPHERES = 1, 2, 4

f32, u32, i32, u64, u16, u8 = C.c_float, C.c_uint32, C.c_int32, C.c_uint64, C.c_uint16, C.c_uint8


class RtbHittable(C.Structure):
    _fields_ = [("type", u32), ("material", u32), ("is_moving", u32), ("radius", f32),
                ("a", f32 * 3), ("b", f32 * 3), ("c", f32 * 3), ("sin_theta", f32), ("cos_theta", f32),
                ("child", u32)]


class RtbMaterial(C.Structure):
    _fields_ = [("type", u32), ("texture", u32), ("albedo", f32 * 3), ("fuzz", f32), ("ir", f32), ("reserved", u32)]


class RtbTexture(C.Structure):
    _fields_ = [("type", u32), ("index", u32), ("scale", f32), ("color", f32 * 3), ("color2", f32 * 3),
                ("reserved", u32 * 3)]


class RtbPerlin(C.Structure):
    _fields_ = [("ranvec", (f32 * 3) * 256), ("perm_x", u16 * 256), ("perm_y", u16 * 256), ("perm_z", u16 * 256)]


class RtbImage(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("bytes_per_row", u32), ("reserved", u32),
                ("data", C.POINTER(u8))]


class RtbBvhNode(C.Structure):
    _fields_ = [("bmin", f32 * 3), ("bmax", f32 * 3), ("left", i32), ("right", i32), ("leaf", i32), ("reserved", u32)]


class RtbSceneDesc(C.Structure):
    _fields_ = [("abi_version", u32), ("n_nodes", u32), ("n_hittables", u32), ("n_materials", u32),
                ("n_textures", u32), ("n_perlins", u32), ("n_images", u32), ("root", i32),
                ("nodes", C.POINTER(RtbBvhNode)), ("hittables", C.POINTER(RtbHittable)),
                ("materials", C.POINTER(RtbMaterial)), ("textures", C.POINTER(RtbTexture)),
                ("perlins", C.POINTER(RtbPerlin)), ("images", C.POINTER(RtbImage))]


class RtbCamera(C.Structure):
    _fields_ = [("image_width", u32), ("image_height", u32), ("samples_per_pixel", u32), ("max_depth", u32),
                ("center", f32 * 3), ("pixel00_loc", f32 * 3), ("pixel_delta_u", f32 * 3),
                ("pixel_delta_v", f32 * 3), ("defocus_disk_u", f32 * 3), ("defocus_disk_v", f32 * 3),
                ("defocus_angle", f32), ("background", f32 * 3), ("background_mode", u32), ("reserved", u32)]


class RtbRenderOptions(C.Structure):
    _fields_ = [("seed", u64), ("sample_begin", u32), ("sample_count", u32), ("pixel_begin", u32),
                ("pixel_count", u32), ("tile_rank", u32), ("tile_world", u32), ("integrator", u32),
                ("traversal", u32), ("flags", u32), ("samples_per_launch", u32)]


class RtbRenderStats(C.Structure):
    _fields_ = [("n_paths", u64), ("n_rays", u64), ("n_box_tests", u64), ("n_object_tests", u64), ("n_hits", u64),
                ("device_ms", C.c_double), ("n_launches", u32), ("reserved", u32)]


class RtbRay(C.Structure):
    _fields_ = [("origin", f32 * 3), ("direction", f32 * 3), ("time", f32), ("t_min", f32), ("t_max", f32)]


class RtbHit(C.Structure):
    _fields_ = [("object", i32), ("front_face", u32), ("t", f32), ("p", f32 * 3), ("normal", f32 * 3),
                ("u", f32), ("v", f32), ("n_box_tests", u32), ("n_object_tests", u32)]


class RtwSceneParams(C.Structure):
    _fields_ = [("kind", u32), ("flags", u32), ("scene_seed", u64), ("bvh_seed", u64), ("perlin_seed", u64),
                ("n_spheres", u32), ("image_width", u32), ("image_height", u32), ("reserved", u32),
                ("image_rgba", C.POINTER(u8))]


class RtwMaterialSpec(C.Structure):
    _fields_ = [("material", u32), ("texture", u32), ("color", f32 * 3), ("color2", f32 * 3), ("scale", f32),
                ("fuzz", f32), ("ir", f32), ("image_index", u32), ("perlin_seed", u64)]


class RtwCameraOptions(C.Structure):
    _fields_ = [("aspect_ratio", f32), ("image_width", u32), ("image_height", u32), ("samples_per_pixel", u32),
                ("max_depth", u32), ("background", f32 * 3), ("vfov", f32), ("lookfrom", f32 * 3),
                ("lookat", f32 * 3), ("vup", f32 * 3), ("defocus_angle", f32), ("focus_dist", f32),
                ("background_mode", u32)]


# numpy dtypes matching RtbRay / RtbHit (packed float/int records)
RAY_DTYPE = [("origin", "<f4", 3), ("direction", "<f4", 3), ("time", "<f4"), ("t_min", "<f4"), ("t_max", "<f4")]
HIT_DTYPE = [("object", "<i4"), ("front_face", "<u4"), ("t", "<f4"), ("p", "<f4", 3), ("normal", "<f4", 3),
             ("u", "<f4"), ("v", "<f4"), ("n_box_tests", "<u4"), ("n_object_tests", "<u4")]

# Every symbol include/rtb.h and include/rtw_host.h declare (checked by the CPU test-suite).
class RtbIpcHandle(C.Structure):
    _fields_ = [("bytes", C.c_uint8 * 64)]


RTB_SYMBOLS = [
    "rtb_abi_version", "rtb_last_error", "rtb_device_count", "rtb_scene_create", "rtb_scene_destroy",
    "rtb_trace_rays", "rtb_render", "rtb_render_device", "rtb_resolve_device", "rtb_resolve", "rtb_render_async",
    "rtb_job_progress", "rtb_job_cancel", "rtb_job_wait", "rtb_job_destroy", "rtb_philox_device_selftest",
    "rtb_measure_fp32_peak", "rtb_debug_build_layout", "rtb_debug_packed_layout", "rtb_buffer_alloc", "rtb_buffer_free", "rtb_ipc_export",
    "rtb_ipc_open", "rtb_ipc_close", "rtb_exchange_slice", "rtb_exchange_resolve",
    "rtb_group_create", "rtb_group_destroy", "rtb_group_size", "rtb_group_render",
]
RTB_PARTITION_SAMPLES, RTB_PARTITION_TILES = 0, 1
RTW_SYMBOLS = [
    "rtw_world_create", "rtw_world_new", "rtw_world_add_image", "rtw_world_add_sphere", "rtw_world_add_quad",
    "rtw_world_add_box", "rtw_world_add_medium", "rtw_obj_sphere", "rtw_obj_quad", "rtw_obj_box", "rtw_obj_list",
    "rtw_obj_translate", "rtw_obj_rotate_y", "rtw_obj_medium", "rtw_world_add_object",
    "rtw_world_build", "rtw_world_desc", "rtw_world_object_box", "rtw_world_destroy", "rtw_camera_defaults",
    "rtw_camera_init", "rtw_camera_render", "rtw_write_ppm", "rtw_write_png",
]

_rtb = None
_rtw = None


def _load(name: str) -> C.CDLL:
    path = os.path.join(LIB_DIR, name)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: the CUDA extension has not been built. Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or `make -C zig-raytracing-weekend_b200`). There is no CPU/Python fallback.")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


def rtb() -> C.CDLL:
    """librtb.so with argtypes set."""
    global _rtb
    if _rtb is not None:
        return _rtb
    lib = _load("librtb.so")
    vp = C.c_void_p
    lib.rtb_abi_version.restype = u32
    lib.rtb_last_error.restype = C.c_char_p
    lib.rtb_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.rtb_scene_create.argtypes = [C.POINTER(RtbSceneDesc), C.c_int, C.POINTER(vp)]
    lib.rtb_scene_destroy.argtypes = [vp]
    lib.rtb_trace_rays.argtypes = [vp, vp, u64, u32, vp]
    lib.rtb_render.argtypes = [vp, C.POINTER(RtbCamera), C.POINTER(RtbRenderOptions), vp, vp,
                               C.POINTER(RtbRenderStats)]
    lib.rtb_render_device.argtypes = [vp, C.POINTER(RtbCamera), C.POINTER(RtbRenderOptions), vp, vp,
                                      C.POINTER(RtbRenderStats)]
    lib.rtb_resolve_device.argtypes = [vp, vp, u64, f32, C.c_int, vp]
    lib.rtb_resolve.argtypes = [vp, vp, u64, f32, C.c_int]
    lib.rtb_render_async.argtypes = [vp, C.POINTER(RtbCamera), C.POINTER(RtbRenderOptions), vp, vp, C.POINTER(vp)]
    lib.rtb_job_progress.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(C.c_int)]
    lib.rtb_job_cancel.argtypes = [vp]
    lib.rtb_job_wait.argtypes = [vp, C.POINTER(RtbRenderStats)]
    lib.rtb_job_destroy.argtypes = [vp]
    lib.rtb_philox_device_selftest.argtypes = [vp, vp, u32, vp, C.c_int]
    lib.rtb_measure_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.rtb_debug_build_layout.argtypes = [C.POINTER(RtbSceneDesc), u32, u32, vp, C.POINTER(u32)]
    lib.rtb_debug_packed_layout.argtypes = [C.POINTER(RtbSceneDesc), u32, vp, C.POINTER(u32), C.POINTER(f32 * 3),
                                            C.POINTER(f32 * 3)]
    lib.rtb_buffer_alloc.argtypes = [C.c_int, u64, C.POINTER(vp)]
    lib.rtb_buffer_free.argtypes = [C.c_int, vp]
    lib.rtb_ipc_export.argtypes = [C.c_int, vp, C.POINTER(RtbIpcHandle)]
    lib.rtb_ipc_open.argtypes = [C.c_int, C.POINTER(RtbIpcHandle), C.POINTER(vp)]
    lib.rtb_ipc_close.argtypes = [C.c_int, vp]
    lib.rtb_exchange_slice.argtypes = [u64, u32, u32, u32, C.POINTER(u64), C.POINTER(u64)]
    lib.rtb_exchange_resolve.argtypes = [C.POINTER(vp), u32, u32, u32, vp, vp, u64, f32, C.c_int, vp]
    lib.rtb_group_create.argtypes = [C.POINTER(RtbSceneDesc), C.POINTER(C.c_int), u32, C.POINTER(vp)]
    lib.rtb_group_destroy.argtypes = [vp]
    lib.rtb_group_size.argtypes = [vp, C.POINTER(u32)]
    lib.rtb_group_render.argtypes = [vp, C.POINTER(RtbCamera), C.POINTER(RtbRenderOptions), u32, vp, vp,
                                     C.POINTER(RtbRenderStats)]
    for s in RTB_SYMBOLS:
        fn = getattr(lib, s)
        if s not in ("rtb_abi_version", "rtb_last_error"):
            fn.restype = C.c_int
    _rtb = lib
    return lib


def rtw() -> C.CDLL:
    """librtw_host.so with argtypes set (loads librtb.so first)."""
    global _rtw
    if _rtw is not None:
        return _rtw
    rtb()
    lib = _load("librtw_host.so")
    vp = C.c_void_p
    lib.rtw_world_create.argtypes = [C.POINTER(RtwSceneParams), C.POINTER(vp)]
    lib.rtw_world_new.argtypes = [C.POINTER(vp)]
    lib.rtw_world_add_image.argtypes = [vp, vp, u32, u32]
    lib.rtw_world_add_sphere.argtypes = [vp, C.POINTER(f32 * 3), C.POINTER(f32 * 3), f32, C.POINTER(RtwMaterialSpec)]
    lib.rtw_world_add_quad.argtypes = [vp, C.POINTER(f32 * 3), C.POINTER(f32 * 3), C.POINTER(f32 * 3),
                                       C.POINTER(RtwMaterialSpec)]
    lib.rtw_world_add_box.argtypes = [vp, C.POINTER(f32 * 3), C.POINTER(f32 * 3), C.c_int, f32, C.POINTER(f32 * 3),
                                      C.POINTER(RtwMaterialSpec)]
    lib.rtw_world_add_medium.argtypes = [vp, C.POINTER(f32 * 3), C.POINTER(f32 * 3), C.c_int, f32, C.POINTER(f32 * 3),
                                         f32, C.POINTER(f32 * 3)]
    hp = C.POINTER(u32)
    lib.rtw_obj_sphere.argtypes = [vp, C.POINTER(f32 * 3), C.POINTER(f32 * 3), f32, C.POINTER(RtwMaterialSpec), hp]
    lib.rtw_obj_quad.argtypes = [vp, C.POINTER(f32 * 3), C.POINTER(f32 * 3), C.POINTER(f32 * 3), C.POINTER(RtwMaterialSpec), hp]
    lib.rtw_obj_box.argtypes = [vp, C.POINTER(f32 * 3), C.POINTER(f32 * 3), C.POINTER(RtwMaterialSpec), hp]
    lib.rtw_obj_list.argtypes = [vp, C.POINTER(u32), u32, hp]
    lib.rtw_obj_translate.argtypes = [vp, u32, C.POINTER(f32 * 3), hp]
    lib.rtw_obj_rotate_y.argtypes = [vp, u32, f32, hp]
    lib.rtw_obj_medium.argtypes = [vp, u32, f32, C.POINTER(f32 * 3), hp]
    lib.rtw_world_add_object.argtypes = [vp, u32]
    lib.rtw_world_build.argtypes = [vp, u64]
    lib.rtw_world_desc.argtypes = [vp]
    lib.rtw_world_desc.restype = C.POINTER(RtbSceneDesc)
    lib.rtw_world_object_box.argtypes = [vp, u32, C.POINTER(f32 * 6)]
    lib.rtw_world_destroy.argtypes = [vp]
    lib.rtw_world_destroy.restype = None
    lib.rtw_camera_defaults.argtypes = [C.POINTER(RtwCameraOptions)]
    lib.rtw_camera_defaults.restype = None
    lib.rtw_camera_init.argtypes = [C.POINTER(RtwCameraOptions), C.POINTER(RtbCamera)]
    lib.rtw_camera_render.argtypes = [vp, C.POINTER(RtwCameraOptions), C.POINTER(RtbRenderOptions), C.c_int, vp, vp,
                                      C.POINTER(RtbRenderStats)]
    lib.rtw_write_ppm.argtypes = [C.c_char_p, vp, u32, u32]
    lib.rtw_write_png.argtypes = [C.c_char_p, vp, u32, u32]
    _rtw = lib
    return lib
