//! rtb.zig — Zig declarations of the C ABI in include/rtb.h (RTB_ABI_VERSION 1), for dariooddenino/zig-raytracing-weekend.
//!
//! Drop this file and lower.zig into `src/`, link `librtb.so` (build.zig: `exe.linkSystemLibrary("rtb");
//! exe.addLibraryPath(.{ .path = "zig-raytracing-weekend_b200/_lib" }); exe.linkLibC();`).
//!
//! NOT COMPILED in the build image (no Zig toolchain there; the reference's HEAD needs Zig 0.12-dev + Dawn prebuilts).
//! The compiled, tested twins of this file are include/rtb.h itself (tests/abi_smoke.c builds it as C99) and
//! zig-raytracing-weekend_b200/_ffi.py (ctypes), whose struct sizes are asserted in tests/test_host_and_abi.py.
//! `extern struct` has C layout; field order and types follow rtb.h one to one.

pub const ABI_VERSION: u32 = 1;

// RtbStatus
pub const OK: c_int = 0;
pub const ERR_INVALID_ARGUMENT: c_int = -1;
pub const ERR_CUDA: c_int = -2;
pub const ERR_NO_DEVICE: c_int = -3;
pub const ERR_OUT_OF_MEMORY: c_int = -4;
pub const ERR_CANCELLED: c_int = -5;
pub const ERR_UNSUPPORTED: c_int = -6;

pub const HITTABLE_SPHERE: u32 = 0;
pub const HITTABLE_QUAD: u32 = 1;
pub const HITTABLE_BOX: u32 = 2; // Translate(RotateY(createBox(a, b, mat), angle), offset)
pub const HITTABLE_CONSTANT_MEDIUM: u32 = 3;
pub const HITTABLE_TRANSLATE: u32 = 4; // general wrappers: `child` indexes another entry of the hittable array
pub const HITTABLE_ROTATE_Y: u32 = 5;
pub const HITTABLE_LIST: u32 = 6;
pub const HITTABLE_MEDIUM_OF: u32 = 7;

pub const MAT_LAMBERTIAN: u32 = 0;
pub const MAT_METAL: u32 = 1;
pub const MAT_DIELECTRIC: u32 = 2;
pub const MAT_DIFFUSE_LIGHT: u32 = 3;
pub const MAT_ISOTROPIC: u32 = 4;

pub const TEX_SOLID: u32 = 0;
pub const TEX_CHECKER: u32 = 1;
pub const TEX_IMAGE: u32 = 2;
pub const TEX_NOISE: u32 = 3;

pub const BACKGROUND_SOLID: u32 = 0; // HEAD: `return self.background` (src/camera.zig:207)
pub const BACKGROUND_SKY: u32 = 1; // the Book-1 gradient kept as a comment at src/camera.zig:204-206

pub const INTEGRATOR_MEGAKERNEL: u32 = 0;
pub const INTEGRATOR_WAVEFRONT: u32 = 1;

pub const TRAVERSAL_REFERENCE: u32 = 0; // src/bvh.zig:122-136 order over the host's tree: bit-exact hit index
pub const TRAVERSAL_ORDERED: u32 = 1;
pub const TRAVERSAL_SAH: u32 = 2;
pub const TRAVERSAL_SAH16: u32 = 3;

pub const FLAG_COUNT_WORK: u32 = 1;

pub const PARTITION_SAMPLES: u32 = 0;
pub const PARTITION_TILES: u32 = 1;

pub const RtbHittable = extern struct { // 64 bytes
    type: u32,
    material: u32,
    is_moving: u32 = 0,
    radius: f32 = 0,
    a: [3]f32 = .{ 0, 0, 0 },
    b: [3]f32 = .{ 0, 0, 0 },
    c: [3]f32 = .{ 0, 0, 0 },
    sin_theta: f32 = 0,
    cos_theta: f32 = 1,
    child: u32 = 0, // translate / rotate_y / medium_of: the wrapped hittable; list: its first member (count in `material`)
};

pub const RtbMaterial = extern struct { // 32 bytes
    type: u32,
    texture: u32 = 0,
    albedo: [3]f32 = .{ 0, 0, 0 },
    fuzz: f32 = 0,
    ir: f32 = 1,
    reserved: u32 = 0,
};

pub const RtbTexture = extern struct { // 48 bytes
    type: u32,
    index: u32 = 0,
    scale: f32 = 1,
    color: [3]f32 = .{ 0, 0, 0 },
    color2: [3]f32 = .{ 0, 0, 0 },
    reserved: [3]u32 = .{ 0, 0, 0 },
};

pub const RtbPerlin = extern struct {
    ranvec: [256][3]f32,
    perm_x: [256]u16,
    perm_y: [256]u16,
    perm_z: [256]u16,
};

pub const RtbImage = extern struct {
    width: u32,
    height: u32,
    bytes_per_row: u32,
    reserved: u32 = 0,
    data: ?[*]const u8,
};

pub const RtbBvhNode = extern struct { // 40 bytes
    bmin: [3]f32,
    bmax: [3]f32,
    left: i32 = -1,
    right: i32 = -1,
    leaf: i32 = -1,
    reserved: u32 = 0,
};

pub const RtbSceneDesc = extern struct {
    abi_version: u32 = ABI_VERSION,
    n_nodes: u32,
    n_hittables: u32,
    n_materials: u32,
    n_textures: u32,
    n_perlins: u32,
    n_images: u32,
    root: i32,
    nodes: ?[*]const RtbBvhNode,
    hittables: ?[*]const RtbHittable,
    materials: ?[*]const RtbMaterial,
    textures: ?[*]const RtbTexture,
    perlins: ?[*]const RtbPerlin,
    images: ?[*]const RtbImage,
};

pub const RtbCamera = extern struct {
    image_width: u32,
    image_height: u32,
    samples_per_pixel: u32,
    max_depth: u32,
    center: [3]f32,
    pixel00_loc: [3]f32,
    pixel_delta_u: [3]f32,
    pixel_delta_v: [3]f32,
    defocus_disk_u: [3]f32,
    defocus_disk_v: [3]f32,
    defocus_angle: f32,
    background: [3]f32,
    background_mode: u32 = BACKGROUND_SOLID,
    reserved: u32 = 0,
};

pub const RtbRenderOptions = extern struct {
    seed: u64 = 1234,
    sample_begin: u32 = 0,
    sample_count: u32 = 0, // 0 = camera.samples_per_pixel
    pixel_begin: u32 = 0, // Task{thread_idx, chunk_size}: pixel_begin = thread_idx * chunk_size, pixel_count = chunk_size
    pixel_count: u32 = 0,
    tile_rank: u32 = 0,
    tile_world: u32 = 0,
    integrator: u32 = INTEGRATOR_WAVEFRONT,
    traversal: u32 = TRAVERSAL_REFERENCE,
    flags: u32 = 0,
    samples_per_launch: u32 = 0,
};

pub const RtbRenderStats = extern struct {
    n_paths: u64 = 0,
    n_rays: u64 = 0,
    n_box_tests: u64 = 0,
    n_object_tests: u64 = 0,
    n_hits: u64 = 0,
    device_ms: f64 = 0,
    n_launches: u32 = 0,
    reserved: u32 = 0,
};

pub const RtbRay = extern struct { origin: [3]f32, direction: [3]f32, time: f32, t_min: f32 = 0.001, t_max: f32 = @import("std").math.inf(f32) };
pub const RtbHit = extern struct { object: i32, front_face: u32, t: f32, p: [3]f32, normal: [3]f32, u: f32, v: f32, n_box_tests: u32, n_object_tests: u32 };
pub const RtbIpcHandle = extern struct { bytes: [64]u8 };

pub const RtbScene = opaque {};
pub const RtbJob = opaque {};
pub const RtbSceneGroup = opaque {};

pub extern fn rtb_abi_version() u32;
pub extern fn rtb_last_error() [*:0]const u8;
pub extern fn rtb_device_count(count: *c_int) c_int;
pub extern fn rtb_scene_create(desc: *const RtbSceneDesc, device: c_int, scene_out: *?*RtbScene) c_int;
pub extern fn rtb_scene_destroy(scene: ?*RtbScene) c_int;
pub extern fn rtb_trace_rays(scene: *RtbScene, rays: [*]const RtbRay, n: u64, traversal: u32, hits_out: [*]RtbHit) c_int;
pub extern fn rtb_render(scene: *RtbScene, camera: *const RtbCamera, options: *const RtbRenderOptions, accum: [*]f32, rgba: ?[*]u8, stats: ?*RtbRenderStats) c_int;
pub extern fn rtb_render_device(scene: *RtbScene, camera: *const RtbCamera, options: *const RtbRenderOptions, d_accum: *anyopaque, cuda_stream: ?*anyopaque, stats: ?*RtbRenderStats) c_int;
pub extern fn rtb_resolve_device(d_accum: *const anyopaque, d_rgba: *anyopaque, n_pixels: u64, n_samples_override: f32, device: c_int, cuda_stream: ?*anyopaque) c_int;
pub extern fn rtb_resolve(accum: [*]const f32, rgba: [*]u8, n_pixels: u64, n_samples_override: f32, device: c_int) c_int;
pub extern fn rtb_render_async(scene: *RtbScene, camera: *const RtbCamera, options: *const RtbRenderOptions, accum: [*]f32, rgba: ?[*]u8, job_out: *?*RtbJob) c_int;
pub extern fn rtb_job_progress(job: *RtbJob, samples_done: ?*u32, samples_total: ?*u32, running: ?*c_int) c_int;
pub extern fn rtb_job_cancel(job: *RtbJob) c_int;
pub extern fn rtb_job_wait(job: *RtbJob, stats: ?*RtbRenderStats) c_int;
pub extern fn rtb_job_destroy(job: ?*RtbJob) c_int;
// multi-GPU behind one call (one process): one scene replica + one host thread per device, peer-memory exchange
pub extern fn rtb_group_create(desc: *const RtbSceneDesc, devices: [*]const c_int, n_devices: u32, group_out: *?*RtbSceneGroup) c_int;
pub extern fn rtb_group_destroy(group: ?*RtbSceneGroup) c_int;
pub extern fn rtb_group_size(group: *const RtbSceneGroup, n_devices_out: *u32) c_int;
pub extern fn rtb_group_render(group: *RtbSceneGroup, camera: *const RtbCamera, options: *const RtbRenderOptions, partition: u32, accum: [*]f32, rgba: ?[*]u8, stats: ?*RtbRenderStats) c_int;
// one process per GPU: peer-memory exchange building blocks
pub extern fn rtb_buffer_alloc(device: c_int, bytes: u64, device_ptr_out: *?*anyopaque) c_int;
pub extern fn rtb_buffer_free(device: c_int, device_ptr: ?*anyopaque) c_int;
pub extern fn rtb_ipc_export(device: c_int, device_ptr: *const anyopaque, handle_out: *RtbIpcHandle) c_int;
pub extern fn rtb_ipc_open(device: c_int, handle: *const RtbIpcHandle, device_ptr_out: *?*anyopaque) c_int;
pub extern fn rtb_ipc_close(device: c_int, device_ptr: ?*anyopaque) c_int;
pub extern fn rtb_exchange_slice(n_pixels: u64, world: u32, rank: u32, root: u32, begin_out: *u64, end_out: *u64) c_int;
pub extern fn rtb_exchange_resolve(peer_accum: [*]const ?*const f32, world: u32, rank: u32, root: u32, root_accum_out: *f32, root_rgba_out: *u8, n_pixels: u64, samples_per_pixel: f32, device: c_int, cuda_stream: ?*anyopaque) c_int;
