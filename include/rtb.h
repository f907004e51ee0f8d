/*
 * rtb.h — C ABI of the B200-native path-tracing hot path ("rtb" = ray-trace-B200).
 *
 * This is the drop-in boundary for dariooddenino/zig-raytracing-weekend.  The reference has no
 * FFI layer; the seam it replaces is the Zig method
 *     Camera.render(self:*Camera, raytrace:*RayTraceState, context:Task) !void   (src/camera.zig:93-116)
 * as invoked by RenderThread.renderFn (src/main.zig:66-68) on 8 threads spawned by startRender
 * (src/main.zig:314-326).  The Zig host keeps scene construction (Camera options + Camera.init,
 * the hittable list, BVHTree.init) and display; it lowers its pointer graph into the POD arrays
 * below ONCE (rtb_scene_create) and then calls rtb_render* instead of spawning render threads.
 *
 * Conventions: plain pointers + sizes, no C++/torch types.  Every function returns RTB_OK (0) or
 * a negative RtbStatus; rtb_last_error() gives a thread-local message.  Nothing throws or aborts
 * across the boundary.  Host buffers are owned by the caller and are not retained after a
 * synchronous call returns (async jobs: until rtb_job_wait / rtb_job_destroy).
 *
 * All floats are IEEE binary32.  Vectors are 3 packed floats (the reference uses @Vector(3,f32)).
 */
#ifndef RTB_H
#define RTB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB_ABI_VERSION 1u

typedef enum RtbStatus {
    RTB_OK = 0,
    RTB_ERR_INVALID_ARGUMENT = -1,
    RTB_ERR_CUDA = -2,
    RTB_ERR_NO_DEVICE = -3,
    RTB_ERR_OUT_OF_MEMORY = -4,
    RTB_ERR_CANCELLED = -5,
    RTB_ERR_UNSUPPORTED = -6
} RtbStatus;

/* ------------------------------------------------------------------------------------------
 * Scene description (host side, read once by rtb_scene_create).
 * ---------------------------------------------------------------------------------------- */

/* Hittable variants lowered from the tagged union at src/objects.zig:39-47.
 * LIST/TRANSLATE/ROTATE_Y/CONSTANT_MEDIUM are SURVEY §8(f) "next" rows; RoundBox is unfinished
 * in the reference (src/objects.zig:171-192) and has no tag. */
enum {
    RTB_HITTABLE_SPHERE = 0,
    RTB_HITTABLE_QUAD = 1,
    RTB_HITTABLE_BOX = 2,             /* Translate(RotateY(createBox)) as ONE record (the shape HEAD's scenes build) */
    RTB_HITTABLE_CONSTANT_MEDIUM = 3, /* ConstantMedium over such a box                                             */
    /* General instancing (src/objects.zig:264-443): wrappers refer to OTHER entries of RtbSceneDesc.hittables through
     * `child`; wrapped entries are not BVH leaves themselves.  child must be greater than the wrapper's own index
     * (no cycles), list members are contiguous. */
    RTB_HITTABLE_TRANSLATE = 4,       /* Translate{offset = a, object = hittables[child]}           (:308-346) */
    RTB_HITTABLE_ROTATE_Y = 5,        /* RotateY{sin_theta, cos_theta, object = hittables[child]}   (:348-443) */
    RTB_HITTABLE_LIST = 6,            /* HittableList{objects = hittables[child .. child + count)}, count in `material`
                                       * (:264-305): members tried in order with ray_t.max = closest so far          */
    RTB_HITTABLE_MEDIUM_OF = 7        /* ConstantMedium{boundary = hittables[child], neg_inv_density = radius,
                                       * phase_function = material} over ANY boundary (:445-508)                      */
};

/* Material variants, src/material.zig:11-16. */
enum {
    RTB_MAT_LAMBERTIAN = 0,
    RTB_MAT_METAL = 1,
    RTB_MAT_DIELECTRIC = 2,
    RTB_MAT_DIFFUSE_LIGHT = 3,
    RTB_MAT_ISOTROPIC = 4
};

/* Texture variants, src/textures.zig:10-14. */
enum { RTB_TEX_SOLID = 0, RTB_TEX_CHECKER = 1, RTB_TEX_IMAGE = 2, RTB_TEX_NOISE = 3 };

/* One `Hittable` of the world list (`world_objects.items[i]`, e.g. src/main.zig:309).  Its
 * position in RtbSceneDesc.hittables is the "object index" reported by rtb_trace_rays.
 *   sphere (src/objects.zig:68-75): a = center1, b = center_vec (zero unless is_moving), radius
 *   quad   (src/objects.zig:195-204): a = q, b = u, c = v
 *   box    = Translate(RotateY(createBox(a, b, material), angle), c) as one world object, the way
 *            cornellBox builds its two boxes (src/main.zig:182-190): a, b = the corners handed to
 *            createBox (src/objects.zig:510-532, a HittableList of 6 quads), sin_theta / cos_theta =
 *            RotateY's fields (src/objects.zig:350-358; 0 / 1 when there is no RotateY), c =
 *            Translate.offset (src/objects.zig:309; zero when there is no Translate).
 *   constant_medium (src/objects.zig:445-508) whose boundary is such a box, as in cornellBoxSmoke
 *            (src/main.zig:223-236): the box fields as above, radius = neg_inv_density (-1/density, :451),
 *            material = the Isotropic phase function (:451).  Its hit() draws one random number
 *            (:484); that draw is word 0 of Philox block (0x40000000 + object index) of the ray
 *            segment's stream, so it does not depend on the traversal order.  For rtb_trace_rays the
 *            stream is keyed (seed 0; pixel = ray index, sample 0, segment 1).
 *   translate / rotate_y / list / medium_of: the general wrappers, see the enum above.  The object index reported for
 *            a hit inside a wrapper is the index of the TOP-LEVEL object (the BVH leaf); its material is the hit
 *            primitive's.  A medium_of's random draw is keyed by the medium's own index (block 0x40000000 + index).
 * The bounding box is not carried here: the BVH nodes hold the boxes the host computed
 * (Sphere.init / initMoving, src/objects.zig:80-92; RotateY.init :360-397; Translate.init :314-319). */
typedef struct RtbHittable {
    uint32_t type;      /* RTB_HITTABLE_* */
    uint32_t material;  /* index into RtbSceneDesc.materials (the reference stores Material by value);
                         * list: the number of members; translate / rotate_y: unused */
    uint32_t is_moving; /* sphere only, src/objects.zig:72 */
    float radius;       /* sphere only */
    float a[3];
    float b[3];
    float c[3];
    float sin_theta; /* box, rotate_y */
    float cos_theta; /* box, rotate_y */
    uint32_t child;  /* translate / rotate_y / medium_of: index of the wrapped hittable; list: of its first member */
} RtbHittable; /* 64 bytes */

/* src/material.zig:32-144.
 *   lambertian:    texture = albedo texture          (:33)
 *   metal:         albedo, fuzz (already clamped <=1) (:58-63)
 *   dielectric:    ir                                 (:74)
 *   diffuse_light: texture = emit texture             (:109)
 *   isotropic:     texture = albedo texture           (:129) */
typedef struct RtbMaterial {
    uint32_t type;    /* RTB_MAT_* */
    uint32_t texture; /* index into RtbSceneDesc.textures, when the variant has one */
    float albedo[3];
    float fuzz;
    float ir;
    uint32_t reserved;
} RtbMaterial; /* 32 bytes */

/* src/textures.zig:29-124.
 *   solid:   color                                   (:30)
 *   checker: inv_scale (= 1/scale, :54), color = even, color2 = odd (SolidColor only, :50-51)
 *   image:   index into RtbSceneDesc.images          (:76)
 *   noise:   index into RtbSceneDesc.perlins, scale  (:108-109) */
typedef struct RtbTexture {
    uint32_t type;  /* RTB_TEX_* */
    uint32_t index; /* image / perlin table index */
    float scale;    /* checker: inv_scale; noise: scale */
    float color[3];
    float color2[3];
    uint32_t reserved[3];
} RtbTexture; /* 48 bytes */

/* Perlin tables of one NoiseTexture instance, generated on the host by Perlin.init
 * (src/perlin.zig:76-101) and uploaded, never regenerated on the device. */
typedef struct RtbPerlin {
    float ranvec[256][3];
    uint16_t perm_x[256];
    uint16_t perm_y[256];
    uint16_t perm_z[256];
} RtbPerlin;

/* A decoded RGBA8 image as zstbi.Image.loadFromFile(path, 4) yields it
 * (libs/zstbi/src/zstbi.zig:118-152; consumed by src/rtw_image.zig:51-62).
 * Only R,G,B of each 4-byte texel are read. */
typedef struct RtbImage {
    uint32_t width;
    uint32_t height;
    uint32_t bytes_per_row;
    uint32_t reserved;
    const uint8_t* data;
} RtbImage;

/* One BVHNode (src/bvh.zig:106-110) with its pointers turned into indices.  Any node order is
 * accepted; the library re-lays the tree out for the device.
 *   leaf  >= 0: index into hittables, left = right = -1
 *   leaf  <  0: interior, left/right index into nodes
 * bmin/bmax = bounding_box {x,y,z}.{min,max} (src/aabb.zig:13-16). */
typedef struct RtbBvhNode {
    float bmin[3];
    float bmax[3];
    int32_t left;
    int32_t right;
    int32_t leaf;
    uint32_t reserved;
} RtbBvhNode; /* 40 bytes */

typedef struct RtbSceneDesc {
    uint32_t abi_version; /* RTB_ABI_VERSION */
    uint32_t n_nodes;
    uint32_t n_hittables;
    uint32_t n_materials;
    uint32_t n_textures;
    uint32_t n_perlins;
    uint32_t n_images;
    int32_t root; /* index of BVHTree.root (src/bvh.zig:20) */
    const RtbBvhNode* nodes;
    const RtbHittable* hittables;
    const RtbMaterial* materials;
    const RtbTexture* textures;
    const RtbPerlin* perlins;
    const RtbImage* images;
} RtbSceneDesc;

/* ------------------------------------------------------------------------------------------
 * Camera: the fields Camera.init derives (src/camera.zig:118-154) plus the options the hot
 * loop reads (src/camera.zig:71-91).  The host computes them; the device only consumes them.
 * ---------------------------------------------------------------------------------------- */
enum {
    RTB_BACKGROUND_SOLID = 0, /* HEAD: return self.background           (src/camera.zig:207)     */
    RTB_BACKGROUND_SKY = 1    /* legacy Book-1 gradient kept as a comment (src/camera.zig:204-206) */
};

typedef struct RtbCamera {
    uint32_t image_width;
    uint32_t image_height;
    uint32_t samples_per_pixel;
    uint32_t max_depth;
    float center[3];
    float pixel00_loc[3];
    float pixel_delta_u[3];
    float pixel_delta_v[3];
    float defocus_disk_u[3];
    float defocus_disk_v[3];
    float defocus_angle;
    float background[3];
    uint32_t background_mode; /* RTB_BACKGROUND_* */
    uint32_t reserved;
} RtbCamera;

/* ------------------------------------------------------------------------------------------
 * Render options.
 * ---------------------------------------------------------------------------------------- */
enum {
    RTB_INTEGRATOR_MEGAKERNEL = 0, /* one persistent kernel: raygen + traversal + scatter per thread */
    RTB_INTEGRATOR_WAVEFRONT = 1   /* raygen / extend / shade kernels over compacted ray queues      */
};

enum {
    /* Reference order: left subtree, then right with t_max shrunk to the left hit; leaves are not
     * box-tested (src/bvh.zig:122-136).  Bit-exact nearest-hit index against the oracle. */
    RTB_TRAVERSAL_REFERENCE = 0,
    /* Same tree, same slab / sphere arithmetic, but at every interior node the child the ray meets
     * first (by the sign of its direction on the axis separating the children) is visited first, so
     * hits are found earlier and more subtrees are culled.  The nearest hit can differ from the
     * reference's only where float rounding puts a sphere's root on the other side of its own box
     * (measured: tests/test_gpu_parity.py::test_ordered_traversal_agrees). */
    RTB_TRAVERSAL_ORDERED = 1,
    /* The library re-partitions the SAME objects (same bounding boxes, one object per leaf) with a
     * binned surface-area-heuristic tree instead of walking the host's random-axis median-split tree
     * (src/bvh.zig:48-67), and visits the near child first.  Slab test, primitive test and every hit
     * value are computed by the same code; only the set of visited nodes changes, so the same caveat
     * as ORDERED applies.  Not used by the hit-query parity harness unless asked for.
     * Its slab test is one f32 FMA per plane on boxes padded by 2^-21 of the scene extent; that margin is derived for
     * ray origins within ~3.5x the scene extent of the origin.  For cameras much farther away use RTB_TRAVERSAL_SAH16,
     * whose margin is computed per ray and grows with the origin's distance. */
    RTB_TRAVERSAL_SAH = 2,
    /* The SAH tree above, packed for the shared-memory walk of the wavefront integrator: a box node is ONE 16-byte
     * slot — its six planes as binary16 pairs in the tree's own normalised frame, rounded outwards — and the slab test
     * is three packed half-precision FMAs on (entry, exit) pairs whose per-ray constants carry an error margin, so the
     * test is CONSERVATIVE: it may enter a box an exact test would cull, never the reverse.  Leaves keep the
     * reference's f32 arithmetic, so t / p / normal of every hit are the same bits as in the other modes, and the set
     * of leaves tested is a superset of RTB_TRAVERSAL_SAH's.  Half the shared-memory traffic and 12 instead of 18
     * issue slots per node visit.  The walk runs out of shared memory when one octant's packed
     * layout fits (<= 96 KB, ~3 000 objects) and out of global memory otherwise (one 16-byte load per visit instead
     * of two; binary16 planes over a large extent admit ~8 % more visits).  Environment: RTB_PACK_LARGE=0 renders
     * scenes that do not fit shared memory exactly as RTB_TRAVERSAL_SAH (the behaviour before round 2's r3h). */
    RTB_TRAVERSAL_SAH16 = 3
};

enum {
    RTB_FLAG_COUNT_WORK = 1u << 0 /* fill RtbRenderStats.n_box_tests / n_object_tests / n_rays (slower) */
};

typedef struct RtbRenderOptions {
    uint64_t seed;         /* Philox4x32-10 key; streams are keyed (pixel, sample, segment, block)   */
    uint32_t sample_begin; /* first sample index (0-based); Philox uses the global sample index    */
    uint32_t sample_count; /* number of samples to add; 0 means camera.samples_per_pixel            */
    /* Pixel partition.  Task{thread_idx, chunk_size} (src/camera.zig:19, :94-95) maps to
     * pixel_begin = thread_idx*chunk_size, pixel_count = chunk_size.  0/0 means the whole frame. */
    uint32_t pixel_begin;
    uint32_t pixel_count;
    /* Interleaved-tile partition for multi-GPU: of the 32-pixel-wide x 8-row tiles inside the pixel
     * range, this call renders those with tile_index % tile_world == tile_rank.  0/0 or 0/1 = all. */
    uint32_t tile_rank;
    uint32_t tile_world;
    uint32_t integrator; /* RTB_INTEGRATOR_* */
    uint32_t traversal;  /* RTB_TRAVERSAL_*  */
    uint32_t flags;      /* RTB_FLAG_*       */
    uint32_t samples_per_launch; /* progressive/cancel granularity; 0 = library default */
} RtbRenderOptions;

typedef struct RtbRenderStats {
    uint64_t n_paths;        /* camera samples traced                               */
    uint64_t n_rays;         /* ray segments (all bounces)       [RTB_FLAG_COUNT_WORK] */
    uint64_t n_box_tests;    /* interior-node slab tests          [RTB_FLAG_COUNT_WORK] */
    uint64_t n_object_tests; /* leaf primitive tests              [RTB_FLAG_COUNT_WORK] */
    uint64_t n_hits;         /* accepted nearest hits (= shaded)  [RTB_FLAG_COUNT_WORK] */
    double device_ms;        /* CUDA-event time of the render kernels on the launch stream */
    uint32_t n_launches;     /* kernels launched by this call */
    uint32_t reserved;
} RtbRenderStats;

/* ------------------------------------------------------------------------------------------
 * Ray queries for the parity harness (north_star: "given identical ray batches ...").
 * ---------------------------------------------------------------------------------------- */
typedef struct RtbRay {
    float origin[3];
    float direction[3]; /* not normalised, as in the reference (src/camera.zig:176) */
    float time;
    float t_min; /* Interval ray_t (src/camera.zig:187): 0.001 */
    float t_max; /*                                       +inf */
} RtbRay; /* 36 bytes */

typedef struct RtbHit {
    int32_t object;      /* index into hittables, -1 = miss */
    uint32_t front_face; /* HitRecord.front_face (src/objects.zig:28,34) */
    float t;
    float p[3];
    float normal[3];
    float u;
    float v;
    uint32_t n_box_tests;    /* interior slab tests this ray performed */
    uint32_t n_object_tests; /* leaf primitive tests this ray performed */
} RtbHit; /* 52 bytes */

/* ------------------------------------------------------------------------------------------
 * Entry points.
 * ---------------------------------------------------------------------------------------- */
typedef struct RtbScene RtbScene; /* opaque: device copy of one scene, bound to one CUDA device */
typedef struct RtbJob RtbJob;     /* opaque: one asynchronous render                             */

uint32_t rtb_abi_version(void);
const char* rtb_last_error(void);

/* Number of CUDA devices visible; RTB_ERR_NO_DEVICE when there is none (no CPU fallback). */
int rtb_device_count(int* count);

/* Copies the scene to `device` (cudaSetDevice ordinal).  Replaces: nothing is copied in the
 * reference — world is read in place by the worker threads (src/camera.zig:104).
 * The caller's arrays are not retained.  The device layouts of a traversal mode (eight per-octant copies of the tree)
 * are built and uploaded on the first render / ray query that uses the mode, from a host copy of the nodes and
 * hittables the scene keeps; that first call pays for it (tens of milliseconds on Book-1, seconds on a million
 * objects). */
int rtb_scene_create(const RtbSceneDesc* desc, int device, RtbScene** scene_out);
int rtb_scene_destroy(RtbScene* scene);

/* Nearest hit of each ray: `world.hit(r, ray_t)` (src/camera.zig:189 → src/bvh.zig:39-41).
 * rays/hits are HOST arrays of n elements.  traversal = RTB_TRAVERSAL_*. */
int rtb_trace_rays(RtbScene* scene, const RtbRay* rays, uint64_t n, uint32_t traversal, RtbHit* hits_out);

/* The render loop: replaces Camera.render on all threads (src/camera.zig:93-116) and
 * SharedStateImageWriter.writeColor (src/camera.zig:54-66).
 *   accum: HOST float[4*W*H], row-major i = y*W + x, (sum R, sum G, sum B, n) exactly like
 *          writer.buffer; samples are ADDED to what it holds and .w is set to
 *          sample_begin + sample_count for every pixel this call touched.
 *   rgba:  HOST uint8[4*W*H] like writer.texture_buffer, or NULL.  Quantised from accum with
 *          toGamma2 + truncation (src/color.zig:43-62, src/camera.zig:59-64), A = 255.
 *   stats: optional. */
int rtb_render(RtbScene* scene, const RtbCamera* camera, const RtbRenderOptions* options,
               float* accum, uint8_t* rgba, RtbRenderStats* stats);

/* Same, with DEVICE buffers on the scene's device and the caller's CUDA stream (cudaStream_t as
 * void*; NULL = default stream).  Asynchronous with respect to the host unless `stats` is
 * non-NULL (then it synchronises the stream to read the timers/counters).  Used by the
 * multi-GPU driver so that the per-rank accumulators can be reduced in place.
 * Two things a caller should know about the wavefront integrator:
 *   * the FIRST wavefront render of a (scene, depth, frame size) blocks the host once, for about one batch: it learns
 *     from the first batch at which bounce the paths have thinned out enough for the tail kernel (later renders of
 *     the same configuration reuse the value and do not block);
 *   * it keeps ray / hit / path-state queues on the device, allocated on first use and kept with the scene: 624 bytes
 *     per path in flight, 4 Mi paths per pipeline lane, 8 lanes = about 21 GB for frames that fill the batches (small
 *     frames or few samples allocate proportionally less).  If an allocation fails it pipelines over fewer lanes (slower,
 *     same result) and only reports RTB_ERR_OUT_OF_MEMORY when not even one lane fits. */
int rtb_render_device(RtbScene* scene, const RtbCamera* camera, const RtbRenderOptions* options,
                      float* d_accum, void* cuda_stream, RtbRenderStats* stats);

/* toGamma2 + truncation over n_pixels of a DEVICE accumulation buffer into a DEVICE RGBA8 buffer.
 * n_samples_override > 0 replaces accum.w (used after a sample-partitioned reduce). */
int rtb_resolve_device(const float* d_accum, uint8_t* d_rgba, uint64_t n_pixels,
                       float n_samples_override, int device, void* cuda_stream);

/* Host-buffer convenience form of the resolve (copies in, resolves on the GPU, copies out). */
int rtb_resolve(const float* accum, uint8_t* rgba, uint64_t n_pixels, float n_samples_override, int device);

/* ---- multi-GPU exchange over peer memory (one process per GPU) --------------------------------------
 * The reference's 8 render threads write disjoint strips of ONE buffer (src/main.zig:318-323, src/camera.zig:93-99);
 * across GPUs the strips become per-rank accumulation buffers that are combined once per render.  These calls let the
 * ranks map each other's buffers (CUDA IPC over NVLink/NVSwitch) and run the combine as ONE kernel per rank that also
 * does the resolve.  The rendezvous (exchanging the 64-byte handles, the barrier around the kernel) is the caller's:
 * torch.distributed / NCCL in this repo (zig-raytracing-weekend_b200/multigpu.py), any transport for a Zig host. */
typedef struct RtbIpcHandle {
    uint8_t bytes[64]; /* cudaIpcMemHandle_t */
} RtbIpcHandle;

/* Device memory that can be exported to other processes (a cudaMalloc allocation of its own). */
int rtb_buffer_alloc(int device, uint64_t bytes, void** device_ptr_out);
int rtb_buffer_free(int device, void* device_ptr);
int rtb_ipc_export(int device, const void* device_ptr, RtbIpcHandle* handle_out);
/* Maps a buffer exported by ANOTHER process into this one (peer access is enabled as needed). */
int rtb_ipc_open(int device, const RtbIpcHandle* handle, void** device_ptr_out);
int rtb_ipc_close(int device, void* device_ptr);

/* Slice of the frame that rank `rank` of `world` combines: [*begin_out, *end_out), 256-pixel aligned; the slices of
 * ranks 0..world-1 tile [0, n_pixels) in rank order.  The ROOT's NVLink ingress is the bottleneck of the step (it
 * receives every other rank's results on top of whatever it reads itself), so the split is not even: with world >= 3
 * the root combines nothing and only receives; with world == 2 the root combines the whole frame (it then reads one
 * remote buffer and receives nothing); world == 1 is a plain resolve. */
int rtb_exchange_slice(uint64_t n_pixels, uint32_t world, uint32_t rank, uint32_t root, uint64_t* begin_out,
                       uint64_t* end_out);

/* Fused exchange + resolve.  peer_accum[r] = rank r's float4 accumulation buffer as mapped into this process
 * (peer_accum[rank] is this rank's own buffer).  For the pixels of this rank's slice: sum over r in rank order
 * (x, y, z), .w = samples_per_pixel, toGamma2 + truncation; the sums go to root_accum_out and the pixels to
 * root_rgba_out — the ROOT rank's buffers as mapped into this process (root_accum_out may be peer_accum[root]: in
 * place).  Asynchronous on `cuda_stream`.  All ranks must have finished rendering before any rank's launch starts and
 * nobody may touch the buffers again before every rank's launch has finished: bracket it with a stream-ordered
 * barrier (e.g. a 1-element NCCL all-reduce on the same stream). */
int rtb_exchange_resolve(const float* const* peer_accum, uint32_t world, uint32_t rank, uint32_t root,
                         float* root_accum_out, uint8_t* root_rgba_out, uint64_t n_pixels, float samples_per_pixel,
                         int device, void* cuda_stream);

/* ---- multi-GPU behind ONE call (one process) -------------------------------------------------------------------
 * The literal counterpart of startRender (src/main.zig:314-326), which starts all 8 strips of one frame with one call
 * and lets them share one writer.buffer / texture_buffer (src/camera.zig:22-27): the caller creates a GROUP once — one
 * replica of the scene per device — and renders with its host buffers; the library partitions the frame, runs one
 * host thread per device, combines the per-device accumulators over peer memory (cudaDeviceEnablePeerAccess, i.e.
 * NVLink / NVSwitch; the same fused kernel as rtb_exchange_resolve, every device its slice) before gamma and
 * quantisation, and hands back ONE frame.  No launcher, rendezvous or IPC handles on the caller's side. */
typedef struct RtbSceneGroup RtbSceneGroup;
enum {
    RTB_PARTITION_SAMPLES = 0, /* device r renders a contiguous range of the global sample indices (the union of the
                                * paths is what one device would trace; sums differ by float association only) */
    RTB_PARTITION_TILES = 1    /* device r renders the 32x8-pixel tiles t with t % n == r (disjoint pixels: the frame
                                * is bit-identical to the single-device one) */
};
/* devices: CUDA ordinals, 1..16 entries; every pair must be peer-accessible (RTB_ERR_UNSUPPORTED otherwise).  An
 * ordinal may be listed more than once (several replicas on one device: useful for tests on a single-GPU box). */
int rtb_group_create(const RtbSceneDesc* desc, const int* devices, uint32_t n_devices, RtbSceneGroup** group_out);
int rtb_group_destroy(RtbSceneGroup* group);
int rtb_group_size(const RtbSceneGroup* group, uint32_t* n_devices_out);
/* Like rtb_render: accum / rgba are HOST buffers of the whole frame; the samples are ADDED to accum and .w is set to
 * sample_begin + sample_count.  options.pixel_* / tile_* must be 0 (the library partitions); integrator, traversal,
 * seed, sample range and flags apply to every device.  stats: sums over the devices, device_ms = the slowest.
 * Starting from a cleared buffer the TILES frame is bit-identical to rtb_render's; sums already in accum stay on the
 * first device and are added once at the exchange, i.e. in a different float association than sample-by-sample. */
int rtb_group_render(RtbSceneGroup* group, const RtbCamera* camera, const RtbRenderOptions* options, uint32_t partition,
                     float* accum, uint8_t* rgba, RtbRenderStats* stats);

/* Progressive / cancellable render, preserving the GUI behaviour of the reference (progress
 * polling: countSamples src/main.zig:470-477; STOP: stopRender :328-336; per-sample refresh of
 * texture_buffer: src/camera.zig:57-65).  A worker thread renders `samples_per_launch` samples at
 * a time and refreshes the caller's accum/rgba host buffers after each batch. */
int rtb_render_async(RtbScene* scene, const RtbCamera* camera, const RtbRenderOptions* options,
                     float* accum, uint8_t* rgba, RtbJob** job_out);
int rtb_job_progress(RtbJob* job, uint32_t* samples_done, uint32_t* samples_total, int* running);
int rtb_job_cancel(RtbJob* job);                      /* takes effect at the next batch boundary */
int rtb_job_wait(RtbJob* job, RtbRenderStats* stats); /* returns the job's final status */
int rtb_job_destroy(RtbJob* job);

/* Counter-based RNG exposed for tests: Philox4x32-10 (Salmon et al., SC'11) evaluated on the
 * device for n counters; replaces std.crypto.random.float (src/rtweekend.zig:14-16). */
int rtb_philox_device_selftest(const uint32_t* counters4, const uint32_t* key2, uint32_t n,
                               uint32_t* out4, int device);

/* Test hook (no device needed): the library's host-side re-layout of the scene's tree for traversal mode
 * `mode` (RTB_TRAVERSAL_*) and ray octant 0..7, as 8 floats per node — {entry.xyz | center1.xyz, bits(meta)},
 * {exit.xyz | center_vec.xyz, radius} — with the end sentinel at index *n_nodes_out.  out_nodes may be NULL
 * to query the node count; otherwise it must hold 8 * (n_nodes + 1) floats. */
int rtb_debug_build_layout(const RtbSceneDesc* desc, uint32_t mode, uint32_t octant, float* out_nodes,
                           uint32_t* n_nodes_out);

/* Test hook (no device needed): the RTB_TRAVERSAL_SAH16 layout of one octant as raw 16-byte slots (4 words each; see
 * DevScene::pk_nodes in csrc/rtb_device.cuh for the format), sentinel included, plus the normalised frame
 * n = (x - center) / scale.  out_slots may be NULL to query *n_slots_out; otherwise it must hold 4 * n_slots words.
 * Returns RTB_ERR_UNSUPPORTED when the scene is too large for the packed layout. */
int rtb_debug_packed_layout(const RtbSceneDesc* desc, uint32_t octant, uint32_t* out_slots, uint32_t* n_slots_out,
                            float center_out[3], float scale_out[3]);

/* Measurement aid for the roofline: runs a dependent-chain FFMA microbenchmark (8 independent chains
 * per thread, full grid) on `device` and returns the best-of-5 rate in TFLOP/s counting one FMA as two
 * flops.  The path tracer itself is compiled without multiply-add contraction, so its own ceiling is
 * half of this figure. */
int rtb_measure_fp32_peak(int device, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* RTB_H */
