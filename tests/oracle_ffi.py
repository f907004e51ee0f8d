"""ctypes bindings of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (see oracle/oracle.h)."""
import ctypes as C
import importlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ffi = importlib.import_module("zig-raytracing-weekend_b200._ffi")  # PODs of include/rtb.h (data formats only)

f32, u32, u64 = C.c_float, C.c_uint32, C.c_uint64


class OrcCameraOptions(C.Structure):
    _fields_ = [("aspect_ratio", f32), ("image_width", u32), ("image_height", u32), ("samples_per_pixel", u32),
                ("max_depth", u32), ("background", f32 * 3), ("vfov", f32), ("lookfrom", f32 * 3),
                ("lookat", f32 * 3), ("vup", f32 * 3), ("defocus_angle", f32), ("focus_dist", f32),
                ("background_mode", u32)]


lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
vp = C.c_void_p
lib.orc_philox4x32_10.argtypes = [vp, vp, vp]
lib.orc_philox4x32_10.restype = None
lib.orc_u01.argtypes = [u32]
lib.orc_u01.restype = f32
lib.orc_aabb_hit.argtypes = [vp, vp, vp]
lib.orc_sphere_uv.argtypes = [vp, C.POINTER(f32), C.POINTER(f32)]
lib.orc_sphere_uv.restype = None
lib.orc_hittable_hit.argtypes = [vp, u32, vp, vp]
lib.orc_trace_rays.argtypes = [vp, vp, u64, vp]
lib.orc_trace_rays.restype = None
lib.orc_texture_value.argtypes = [vp, u32, f32, f32, vp, vp]
lib.orc_texture_value.restype = None
lib.orc_perlin_noise.argtypes = [vp, vp]
lib.orc_perlin_noise.restype = f32
lib.orc_perlin_turb.argtypes = [vp, vp, C.c_int]
lib.orc_perlin_turb.restype = f32
lib.orc_get_ray.argtypes = [vp, u64, u32, u32, vp]
lib.orc_get_ray.restype = None
lib.orc_scatter.argtypes = [vp, vp, vp, u64, u32, u32, u32, vp, vp]
lib.orc_path_radiance.argtypes = [vp, vp, u64, u32, u32, vp]
lib.orc_path_radiance.restype = None
lib.orc_render.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp]
lib.orc_resolve.argtypes = [vp, vp, u64, f32]
lib.orc_resolve.restype = None
lib.orc_host_random.argtypes = [C.POINTER(u64)]
lib.orc_host_random.restype = f32
lib.orc_host_random_int_range.argtypes = [C.POINTER(u64), u32, u32]
lib.orc_host_random_int_range.restype = u32
lib.orc_camera_init.argtypes = [C.POINTER(OrcCameraOptions), vp]
lib.orc_camera_init.restype = None
lib.orc_sphere_bbox.argtypes = [vp, vp, f32, vp, vp]
lib.orc_sphere_bbox.restype = None
lib.orc_quad_bbox.argtypes = [vp, vp, vp, vp, vp]
lib.orc_quad_bbox.restype = None
lib.orc_bvh_build.argtypes = [vp, vp, u32, C.POINTER(u64), vp]
lib.orc_bvh_build.restype = C.c_int32
lib.orc_perlin_init.argtypes = [C.POINTER(u64), vp]
lib.orc_perlin_init.restype = None


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib.orc_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def make_ray(origin, direction, time=0.0, t_min=0.001, t_max=np.inf):
    r = np.zeros(1, dtype=np.dtype(_ffi.RAY_DTYPE))
    r["origin"][0] = origin
    r["direction"][0] = direction
    r["time"][0] = time
    r["t_min"][0] = t_min
    r["t_max"][0] = t_max
    return r


def aabb_hit(bmin, bmax, ray):
    a = np.asarray(bmin, dtype=np.float32)
    b = np.asarray(bmax, dtype=np.float32)
    return bool(lib.orc_aabb_hit(a.ctypes.data, b.ctypes.data, ray.ctypes.data))


def sphere_uv(p):
    a = np.asarray(p, dtype=np.float32)
    u, v = f32(0), f32(0)
    lib.orc_sphere_uv(a.ctypes.data, C.byref(u), C.byref(v))
    return u.value, v.value


def trace_rays(desc, rays):
    rays = np.ascontiguousarray(rays, dtype=np.dtype(_ffi.RAY_DTYPE))
    hits = np.zeros(rays.shape[0], dtype=np.dtype(_ffi.HIT_DTYPE))
    lib.orc_trace_rays(C.cast(desc, vp), rays.ctypes.data, rays.shape[0], hits.ctypes.data)
    return hits


def get_rays(cam, seed, pixels, samples):
    pixels = np.asarray(pixels, dtype=np.uint32)
    samples = np.broadcast_to(np.asarray(samples, dtype=np.uint32), pixels.shape)
    rays = np.zeros(pixels.shape[0], dtype=np.dtype(_ffi.RAY_DTYPE))
    for i in range(pixels.shape[0]):
        lib.orc_get_ray(C.byref(cam), seed, int(pixels[i]), int(samples[i]), rays[i:i + 1].ctypes.data)
    return rays


def scatter(desc, ray, hit, seed, pixel, sample, segment):
    att = np.zeros(3, dtype=np.float32)
    out = np.zeros(1, dtype=np.dtype(_ffi.RAY_DTYPE))
    ok = lib.orc_scatter(C.cast(desc, vp), ray.ctypes.data, hit.ctypes.data, seed, pixel, sample, segment,
                         att.ctypes.data, out.ctypes.data)
    return bool(ok), att, out


def texture_value(desc, tex, u, v, p):
    pp = np.asarray(p, dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    lib.orc_texture_value(C.cast(desc, vp), tex, u, v, pp.ctypes.data, out.ctypes.data)
    return out


def render(desc, cam, options, n_threads=8, accum=None, want_rgba=True):
    n = cam.image_width * cam.image_height
    if accum is None:
        accum = np.zeros((n, 4), dtype=np.float32)
        accum[:, 3] = 1.0
    rgba = np.zeros((n, 4), dtype=np.uint8) if want_rgba else None
    st = _ffi.RtbRenderStats()
    rc = lib.orc_render(C.cast(desc, vp), C.byref(cam), C.byref(options), n_threads, accum.ctypes.data,
                        rgba.ctypes.data if want_rgba else None, C.byref(st))
    assert rc == 0, rc
    return accum, rgba, {k: getattr(st, k) for k, _ in _ffi.RtbRenderStats._fields_ if k != "reserved"}


def resolve(accum, override=0.0):
    accum = np.ascontiguousarray(accum, dtype=np.float32).reshape(-1, 4)
    rgba = np.zeros((accum.shape[0], 4), dtype=np.uint8)
    lib.orc_resolve(accum.ctypes.data, rgba.ctypes.data, accum.shape[0], override)
    return rgba


def camera_init(camera):
    """camera: the package's Camera dataclass."""
    o = OrcCameraOptions()
    src = camera.options()
    for name, _ in OrcCameraOptions._fields_:
        setattr(o, name, getattr(src, name))
    out = _ffi.RtbCamera()
    lib.orc_camera_init(C.byref(o), C.byref(out))
    return out
