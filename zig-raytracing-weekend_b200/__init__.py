"""zig-raytracing-weekend_b200 — B200-native per-pixel path-tracing hot path behind a C ABI.

Python here is only a driver (tests, bench): it calls the host-side mirror of the reference's
scene/camera API (``include/rtw_host.h``) and the drop-in boundary (``include/rtb.h``) through
ctypes.  All pixel work happens in the CUDA kernels of ``csrc/``; there is no fallback path.

The directory name contains hyphens; import it with
``importlib.import_module("zig-raytracing-weekend_b200")``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _ffi
from ._ffi import *  # noqa: F401,F403  (constants + ctypes structs)


class RtbError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = _ffi.rtb().rtb_last_error()
        super().__init__(f"{where} failed with status {code}: {msg.decode() if msg else ''}")


def _check(rc: int, where: str) -> None:
    if rc != _ffi.RTB_OK:
        raise RtbError(rc, where)


def device_count() -> int:
    n = C.c_int(0)
    _check(_ffi.rtb().rtb_device_count(C.byref(n)), "rtb_device_count")
    return n.value


def _vec3(v):
    return (_ffi.f32 * 3)(*[float(x) for x in v])


# ------------------------------------------------------------------------------------------------
# Camera (src/camera.zig:69-91, :118-154)
# ------------------------------------------------------------------------------------------------
@dataclass
class Camera:
    aspect_ratio: float = 16.0 / 9.0
    image_width: int = 800
    image_height: int = 0
    samples_per_pixel: int = 100
    max_depth: int = 16
    background: tuple = (0.0, 0.0, 0.0)
    vfov: float = 20.0
    lookfrom: tuple = (13.0, 2.0, 3.0)
    lookat: tuple = (0.0, 0.0, 0.0)
    vup: tuple = (0.0, 1.0, 0.0)
    defocus_angle: float = 0.6
    focus_dist: float = 10.0
    background_mode: int = _ffi.RTB_BACKGROUND_SOLID

    def options(self) -> _ffi.RtwCameraOptions:
        o = _ffi.RtwCameraOptions()
        o.aspect_ratio = self.aspect_ratio
        o.image_width = self.image_width
        o.image_height = self.image_height
        o.samples_per_pixel = self.samples_per_pixel
        o.max_depth = self.max_depth
        o.background = _vec3(self.background)
        o.vfov = self.vfov
        o.lookfrom = _vec3(self.lookfrom)
        o.lookat = _vec3(self.lookat)
        o.vup = _vec3(self.vup)
        o.defocus_angle = self.defocus_angle
        o.focus_dist = self.focus_dist
        o.background_mode = self.background_mode
        return o

    def init(self) -> _ffi.RtbCamera:
        """Camera.init: returns the derived camera the device consumes."""
        cam = _ffi.RtbCamera()
        o = self.options()
        _check(_ffi.rtw().rtw_camera_init(C.byref(o), C.byref(cam)), "rtw_camera_init")
        return cam


# ------------------------------------------------------------------------------------------------
# World (hittable list + BVH, src/main.zig scene builders)
# ------------------------------------------------------------------------------------------------
def material_spec(material=_ffi.RTB_MAT_LAMBERTIAN, texture=_ffi.RTB_TEX_SOLID, color=(0.5, 0.5, 0.5),
                  color2=(0.9, 0.9, 0.9), scale=1.0, fuzz=0.0, ir=1.5, image_index=0, perlin_seed=3):
    s = _ffi.RtwMaterialSpec()
    s.material, s.texture = material, texture
    s.color, s.color2 = _vec3(color), _vec3(color2)
    s.scale, s.fuzz, s.ir = scale, fuzz, ir
    s.image_index, s.perlin_seed = image_index, perlin_seed
    return s


class World:
    def __init__(self, handle, keepalive=None):
        self._h = handle
        self._keep = keepalive

    # -- canned scenes -------------------------------------------------------------------------
    @classmethod
    def create(cls, kind, flags=0, scene_seed=1, bvh_seed=2, perlin_seed=3, n_spheres=0, image=None):
        p = _ffi.RtwSceneParams()
        p.kind, p.flags = kind, flags
        p.scene_seed, p.bvh_seed, p.perlin_seed = scene_seed, bvh_seed, perlin_seed
        p.n_spheres = n_spheres
        keep = None
        if image is not None:
            keep = np.ascontiguousarray(image, dtype=np.uint8)
            assert keep.ndim == 3 and keep.shape[2] == 4, "image must be H x W x 4 uint8"
            p.image_height, p.image_width = keep.shape[0], keep.shape[1]
            p.image_rgba = keep.ctypes.data_as(C.POINTER(_ffi.u8))
        h = C.c_void_p()
        _check(_ffi.rtw().rtw_world_create(C.byref(p), C.byref(h)), "rtw_world_create")
        return cls(h, keep)

    @classmethod
    def book1(cls, checker_ground=False, earth=False, moving=True, image=None, **kw):
        flags = (_ffi.RTW_BOOK1_CHECKER_GROUND if checker_ground else 0) | \
                (_ffi.RTW_BOOK1_EARTH_SPHERE if earth else 0) | (0 if moving else _ffi.RTW_BOOK1_STATIC_SPHERES)
        return cls.create(_ffi.RTW_SCENE_BOOK1, flags=flags, image=image, **kw)

    # -- incremental builder -------------------------------------------------------------------
    @classmethod
    def new(cls):
        h = C.c_void_p()
        _check(_ffi.rtw().rtw_world_new(C.byref(h)), "rtw_world_new")
        return cls(h, [])

    def add_image(self, image):
        im = np.ascontiguousarray(image, dtype=np.uint8)
        _check(_ffi.rtw().rtw_world_add_image(self._h, im.ctypes.data, im.shape[1], im.shape[0]), "rtw_world_add_image")

    def add_sphere(self, center, radius, spec, center2=None):
        c1 = _vec3(center)
        c2 = C.byref(_vec3(center2)) if center2 is not None else None
        _check(_ffi.rtw().rtw_world_add_sphere(self._h, C.byref(c1), c2, float(radius), C.byref(spec)),
               "rtw_world_add_sphere")

    def add_quad(self, q, u, v, spec):
        _check(_ffi.rtw().rtw_world_add_quad(self._h, C.byref(_vec3(q)), C.byref(_vec3(u)), C.byref(_vec3(v)),
                                             C.byref(spec)), "rtw_world_add_quad")

    def add_box(self, a, b, spec, angle=None, offset=None):
        """Translate.init(RotateY.init(createBox(a, b, mat), angle), offset); angle/offset None = not applied."""
        off = C.byref(_vec3(offset)) if offset is not None else None
        _check(_ffi.rtw().rtw_world_add_box(self._h, C.byref(_vec3(a)), C.byref(_vec3(b)), int(angle is not None),
                                            float(angle or 0.0), off, C.byref(spec)), "rtw_world_add_box")

    def add_medium(self, a, b, density, color, angle=None, offset=None):
        """ConstantMedium.initFromColor(&Translate(RotateY(createBox(a, b))), density, color)."""
        off = C.byref(_vec3(offset)) if offset is not None else None
        _check(_ffi.rtw().rtw_world_add_medium(self._h, C.byref(_vec3(a)), C.byref(_vec3(b)), int(angle is not None),
                                               float(angle or 0.0), off, float(density), C.byref(_vec3(color))),
               "rtw_world_add_medium")

    # -- general instancing: detached objects, wrapped any number of times, then attached (src/objects.zig:264-443)
    def _handle(self, fn, *args):
        h = _ffi.u32(0)
        _check(fn(self._h, *args, C.byref(h)), fn.__name__)
        return h.value

    def obj_sphere(self, center, radius, spec, center2=None):
        c2 = C.byref(_vec3(center2)) if center2 is not None else None
        return self._handle(_ffi.rtw().rtw_obj_sphere, C.byref(_vec3(center)), c2, float(radius), C.byref(spec))

    def obj_quad(self, q, u, v, spec):
        return self._handle(_ffi.rtw().rtw_obj_quad, C.byref(_vec3(q)), C.byref(_vec3(u)), C.byref(_vec3(v)), C.byref(spec))

    def obj_box(self, a, b, spec):
        return self._handle(_ffi.rtw().rtw_obj_box, C.byref(_vec3(a)), C.byref(_vec3(b)), C.byref(spec))

    def obj_list(self, handles):
        arr = (_ffi.u32 * len(handles))(*handles)
        return self._handle(_ffi.rtw().rtw_obj_list, arr, len(handles))

    def obj_translate(self, handle, offset):
        return self._handle(_ffi.rtw().rtw_obj_translate, handle, C.byref(_vec3(offset)))

    def obj_rotate_y(self, handle, angle):
        return self._handle(_ffi.rtw().rtw_obj_rotate_y, handle, float(angle))

    def obj_medium(self, boundary, density, color):
        return self._handle(_ffi.rtw().rtw_obj_medium, boundary, float(density), C.byref(_vec3(color)))

    def add_object(self, handle):
        _check(_ffi.rtw().rtw_world_add_object(self._h, handle), "rtw_world_add_object")

    def build(self, bvh_seed=2):
        _check(_ffi.rtw().rtw_world_build(self._h, bvh_seed), "rtw_world_build")
        return self

    # -- accessors -----------------------------------------------------------------------------
    @property
    def desc(self):
        d = _ffi.rtw().rtw_world_desc(self._h)
        if not d:
            raise RuntimeError("world has not been built")
        return d

    @property
    def n_objects(self) -> int:
        return self.desc.contents.n_hittables

    @property
    def n_nodes(self) -> int:
        return self.desc.contents.n_nodes

    def object_box(self, i):
        b = (_ffi.f32 * 6)()
        _check(_ffi.rtw().rtw_world_object_box(self._h, i, C.byref(b)), "rtw_world_object_box")
        return np.array(list(b), dtype=np.float32)

    def close(self):
        if self._h:
            _ffi.rtw().rtw_world_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# Scene = device copy of a world; render / trace entry points of the C ABI
# ------------------------------------------------------------------------------------------------
def render_options(seed=1234, sample_begin=0, sample_count=0, pixel_begin=0, pixel_count=0, tile_rank=0,
                   tile_world=0, integrator=_ffi.RTB_INTEGRATOR_MEGAKERNEL, traversal=_ffi.RTB_TRAVERSAL_REFERENCE,
                   flags=0, samples_per_launch=0):
    o = _ffi.RtbRenderOptions()
    o.seed, o.sample_begin, o.sample_count = seed, sample_begin, sample_count
    o.pixel_begin, o.pixel_count, o.tile_rank, o.tile_world = pixel_begin, pixel_count, tile_rank, tile_world
    o.integrator, o.traversal, o.flags, o.samples_per_launch = integrator, traversal, flags, samples_per_launch
    return o


def stats_dict(st) -> dict:
    return {k: getattr(st, k) for k, _ in _ffi.RtbRenderStats._fields_ if k != "reserved"}


def new_writer(cam):
    """SharedStateImageWriter.init (src/camera.zig:29-39): buffer = (0,0,0,1), texture_buffer RGBA8."""
    n = cam.image_width * cam.image_height
    buf = np.zeros((n, 4), dtype=np.float32)
    buf[:, 3] = 1.0
    return buf, np.zeros((n, 4), dtype=np.uint8)


class Scene:
    def __init__(self, world: World, device: int = 0):
        self.world = world
        self.device = device
        h = C.c_void_p()
        _check(_ffi.rtb().rtb_scene_create(world.desc, device, C.byref(h)), "rtb_scene_create")
        self._h = h

    def trace_rays(self, rays: np.ndarray, traversal=_ffi.RTB_TRAVERSAL_REFERENCE) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.dtype(_ffi.RAY_DTYPE))
        hits = np.zeros(rays.shape[0], dtype=np.dtype(_ffi.HIT_DTYPE))
        _check(_ffi.rtb().rtb_trace_rays(self._h, rays.ctypes.data, rays.shape[0], traversal, hits.ctypes.data),
               "rtb_trace_rays")
        return hits

    def render(self, cam, options=None, accum=None, want_rgba=True):
        """rtb_render with host buffers. Returns (accum[N,4] f32, rgba[N,4] u8 | None, stats dict)."""
        options = options if options is not None else render_options()
        n = cam.image_width * cam.image_height
        if accum is None:
            accum, _ = new_writer(cam)
        assert accum.dtype == np.float32 and accum.size == 4 * n and accum.flags.c_contiguous
        rgba = np.zeros((n, 4), dtype=np.uint8) if want_rgba else None
        st = _ffi.RtbRenderStats()
        _check(_ffi.rtb().rtb_render(self._h, C.byref(cam), C.byref(options), accum.ctypes.data,
                                     rgba.ctypes.data if want_rgba else None, C.byref(st)), "rtb_render")
        return accum, rgba, stats_dict(st)

    def render_device(self, cam, options, d_accum_ptr: int, stream: int = 0, want_stats=True):
        st = _ffi.RtbRenderStats()
        _check(_ffi.rtb().rtb_render_device(self._h, C.byref(cam), C.byref(options), d_accum_ptr, stream,
                                            C.byref(st) if want_stats else None), "rtb_render_device")
        return stats_dict(st) if want_stats else None

    def render_async(self, cam, options, accum, rgba):
        job = C.c_void_p()
        _check(_ffi.rtb().rtb_render_async(self._h, C.byref(cam), C.byref(options), accum.ctypes.data,
                                           rgba.ctypes.data if rgba is not None else None, C.byref(job)),
               "rtb_render_async")
        return Job(job, (accum, rgba, cam, options, self))   # the job's worker uses the scene until it is waited for

    def close(self):
        if self._h:
            _ffi.rtb().rtb_scene_destroy(self._h)   # waits for renders in flight on this scene (takes its mutex)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SceneGroup:
    """rtb_group_*: one replica of the world per device, multi-GPU render behind one call (include/rtb.h)."""

    def __init__(self, world: World, devices):
        self.world = world
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        _check(_ffi.rtb().rtb_group_create(world.desc, arr, len(self.devices), C.byref(h)), "rtb_group_create")
        self._h = h

    def render(self, cam, options=None, partition=_ffi.RTB_PARTITION_SAMPLES, accum=None, want_rgba=True):
        options = options if options is not None else render_options()
        n = cam.image_width * cam.image_height
        if accum is None:
            accum, _ = new_writer(cam)
        assert accum.dtype == np.float32 and accum.size == 4 * n and accum.flags.c_contiguous
        rgba = np.zeros((n, 4), dtype=np.uint8) if want_rgba else None
        st = _ffi.RtbRenderStats()
        _check(_ffi.rtb().rtb_group_render(self._h, C.byref(cam), C.byref(options), partition, accum.ctypes.data,
                                           rgba.ctypes.data if want_rgba else None, C.byref(st)), "rtb_group_render")
        return accum, rgba, stats_dict(st)

    def close(self):
        if self._h:
            _ffi.rtb().rtb_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Job:
    def __init__(self, handle, keep):
        self._h, self._keep = handle, keep

    def progress(self):
        done, total, running = _ffi.u32(0), _ffi.u32(0), C.c_int(0)
        _check(_ffi.rtb().rtb_job_progress(self._h, C.byref(done), C.byref(total), C.byref(running)), "rtb_job_progress")
        return done.value, total.value, bool(running.value)

    def cancel(self):
        _check(_ffi.rtb().rtb_job_cancel(self._h), "rtb_job_cancel")

    def wait(self):
        st = _ffi.RtbRenderStats()
        rc = _ffi.rtb().rtb_job_wait(self._h, C.byref(st))
        return rc, stats_dict(st)

    def destroy(self):
        if self._h:
            _ffi.rtb().rtb_job_destroy(self._h)
            self._h = None


def resolve(accum: np.ndarray, n_samples_override: float = 0.0, device: int = 0) -> np.ndarray:
    accum = np.ascontiguousarray(accum, dtype=np.float32).reshape(-1, 4)
    rgba = np.zeros((accum.shape[0], 4), dtype=np.uint8)
    _check(_ffi.rtb().rtb_resolve(accum.ctypes.data, rgba.ctypes.data, accum.shape[0], n_samples_override, device),
           "rtb_resolve")
    return rgba


def philox_device(counters: np.ndarray, key, device: int = 0) -> np.ndarray:
    counters = np.ascontiguousarray(counters, dtype=np.uint32).reshape(-1, 4)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros_like(counters)
    _check(_ffi.rtb().rtb_philox_device_selftest(counters.ctypes.data, k.ctypes.data, counters.shape[0],
                                                 out.ctypes.data, device), "rtb_philox_device_selftest")
    return out


def measure_fp32_peak(device: int = 0) -> float:
    """FFMA microbenchmark, TFLOP/s (FMA = 2 flops)."""
    v = C.c_double(0.0)
    _check(_ffi.rtb().rtb_measure_fp32_peak(device, C.byref(v)), "rtb_measure_fp32_peak")
    return v.value


def write_ppm(path: str, rgba: np.ndarray, width: int, height: int) -> None:
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    _check(_ffi.rtw().rtw_write_ppm(path.encode(), rgba.ctypes.data, width, height), "rtw_write_ppm")


def write_png(path: str, rgba: np.ndarray, width: int, height: int) -> None:
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    assert rgba.size == 4 * width * height
    _check(_ffi.rtw().rtw_write_png(path.encode(), rgba.ctypes.data, width, height), "rtw_write_png")


# BASELINE.json configs: camera + world factory per named scene.
def book1_camera(width=1200, spp=500, max_depth=50):
    """Book-1 final scene camera = the reference's Camera defaults (src/camera.zig:70-91) with the legacy sky."""
    return Camera(image_width=width, samples_per_pixel=spp, max_depth=max_depth,
                  background_mode=_ffi.RTB_BACKGROUND_SKY)


def textured_camera(width=800, spp=256, max_depth=50):
    return Camera(image_width=width, samples_per_pixel=spp, max_depth=max_depth, lookfrom=(0.0, 4.0, 16.0),
                  lookat=(0.0, 2.0, 0.0), vfov=35.0, defocus_angle=0.0, background=(0.70, 0.80, 1.00))


def simple_light_camera(width=800, spp=100, max_depth=50):
    """simpleLightWorld's camera overrides (src/main.zig:156-161); HEAD's black background (src/camera.zig:80)."""
    return Camera(image_width=width, samples_per_pixel=spp, max_depth=max_depth, lookfrom=(26.0, 3.0, 6.0),
                  lookat=(0.0, 2.0, 0.0), vup=(0.0, 1.0, 0.0), defocus_angle=0.0)


def cornell_camera(width=600, spp=200, max_depth=200):
    """cornellBox's camera (src/main.zig:191-200); HEAD's black background."""
    return Camera(image_width=width, aspect_ratio=1.0, samples_per_pixel=spp, max_depth=max_depth, vfov=40.0,
                  lookfrom=(278.0, 278.0, -800.0), lookat=(278.0, 278.0, 0.0), vup=(0.0, 1.0, 0.0), defocus_angle=0.0)


def million_camera(width=3840, spp=64, max_depth=50):
    return Camera(image_width=width, samples_per_pixel=spp, max_depth=max_depth, lookfrom=(520.0, 80.0, 120.0),
                  lookat=(0.0, 10.0, 0.0), vfov=40.0, defocus_angle=0.0, background=(0.70, 0.80, 1.00))
