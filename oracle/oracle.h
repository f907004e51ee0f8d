/*
 * oracle.h — CPU restatement ("oracle") of the reference's path-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (zig-raytracing-weekend_b200/, the
 * C ABI in include/rtb.h) may link, import or call this library.  It exists so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs have something
 * to check the CUDA path against and to time on the host cores.
 *
 * PARITY PARTLY PINNED: the reference (dariooddenino/zig-raytracing-weekend) cannot be compiled or run
 * here (no Zig toolchain, HEAD does not type-check, no headless mode — SURVEY.md §8c) and its
 * own tests pin none of the hot path's numbers.  The oracle is pinned against OUTPUTS OF THE REFERENCE
 * PROGRAM where they are deterministic — the sky, the big spheres' silhouettes, the metal sphere's sky
 * reflection, the refracted sky inside the glass sphere and the sky-lit top of the lambertian sphere in the two renders the reference commits (image.png, image2.png ->
 * tests/golden/ref_renders_top160.npz, tests/test_reference_renders.py: equal to 8-bit rounding) —
 * which anchors Camera.init / getRay, Sphere.hit, Metal.scatter, Dielectric.scatter and Lambertian.scatter (as
 * statistics) and the colour pipeline.  The BVH visiting order and tie rule, textures, quads and media remain UNPINNED.
 * Besides that it is pinned against the few known-answer vectors the reference source carries (the 3
 * AABB cases of the stale test at src/aabb.zig:117-136 and the UV table in the comment at
 * src/objects.zig:105-107, earthmap.jpg through the reference's own vendored stb_image) plus
 * published Philox4x32-10 known-answer vectors.  Every function cites the reference file:line it
 * follows.  Built with -O2 -ffp-contract=off (Zig's default float mode is strict; the reference
 * never calls @setFloatMode).
 *
 * Third-party arithmetic that is NOT under /root/reference: the Zig standard library (version
 * unpinned, build.zig.zon:10 commented out): std.math.acos/atan2/pow/sin/tan, std.sort.heap,
 * std.crypto.random.  The oracle uses glibc libm for the transcendentals (compared with a
 * tolerance), restates std.sort.heap's published sift-down heapsort, and — as BASELINE.json's
 * north_star prescribes — replaces the unseedable CSPRNG by counter-based Philox4x32-10 streams
 * keyed (seed; pixel, sample, segment, block), the SAME schedule the CUDA kernels use, so that a
 * GPU render and an oracle render trace the same paths and can be compared pixel by pixel.
 *
 * The scene/camera PODs are those of include/rtb.h (data formats only; no product code).
 */
#ifndef RTB_ORACLE_H
#define RTB_ORACLE_H

#include "../include/rtb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11; Random123 reference constants). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* u32 -> [0,1): (x >> 8) * 2^-24 */
float orc_u01(uint32_t x);

/* src/aabb.zig:82-114 */
int orc_aabb_hit(const float bmin[3], const float bmax[3], const RtbRay* ray);
/* src/objects.zig:101-114 (getSphereUV) */
void orc_sphere_uv(const float p[3], float* u, float* v);
/* Hittable.hit on a single object of the list (src/objects.zig:49-53 → Sphere.hit :116-148 /
 * Quad.hit :226-261).  Returns 1 on hit and fills `hit` (object = index). */
int orc_hittable_hit(const RtbSceneDesc* scene, uint32_t index, const RtbRay* ray, RtbHit* hit);

/* world.hit(r, ray_t) through BVHTree.hit / BVHNode.hit (src/bvh.zig:39-41, :122-136):
 * recursive, left then right with the shrunk interval, leaves not box-tested. */
void orc_trace_rays(const RtbSceneDesc* scene, const RtbRay* rays, uint64_t n, RtbHit* hits_out);

/* Texture.value (src/textures.zig:22-26) and Perlin.noise / turb (src/perlin.zig:103-152). */
void orc_texture_value(const RtbSceneDesc* scene, uint32_t texture, float u, float v, const float p[3], float out[3]);
float orc_perlin_noise(const RtbPerlin* perlin, const float p[3]);
float orc_perlin_turb(const RtbPerlin* perlin, const float p[3], int depth);

/* Camera.getRay for 0-based flat pixel index i and global sample index s
 * (src/camera.zig:100-103, :156-180). */
void orc_get_ray(const RtbCamera* cam, uint64_t seed, uint32_t pixel, uint32_t sample, RtbRay* ray_out);

/* Material.scatter for one hit (src/material.zig:18-22, :43-54, :65-70, :80-98).
 * segment = 1-based index of the hit along the path (RNG stream key).
 * Returns 1 if scattered; fills attenuation[3] and scattered ray. */
int orc_scatter(const RtbSceneDesc* scene, const RtbRay* r_in, const RtbHit* hit, uint64_t seed, uint32_t pixel,
                uint32_t sample, uint32_t segment, float attenuation[3], RtbRay* scattered);

/* Camera.rayColor (src/camera.zig:182-208), RECURSIVE as in the reference, for one camera sample. */
void orc_path_radiance(const RtbSceneDesc* scene, const RtbCamera* cam, uint64_t seed, uint32_t pixel,
                       uint32_t sample, float out_rgb[3]);

/* Camera.render on `n_threads` OS threads with static contiguous strips, sample-major
 * (src/camera.zig:93-116, src/main.zig:318-324; the reference hard-codes 8 threads, :41) +
 * SharedStateImageWriter.writeColor (src/camera.zig:54-66).  accum/rgba as in rtb_render.
 * Only options->seed, sample_begin, sample_count, pixel_begin, pixel_count are honoured.
 * stats: n_paths, n_rays, n_box_tests, n_object_tests, n_hits and device_ms (= host wall ms of the
 * render loop only, steady_clock) are always filled. */
int orc_render(const RtbSceneDesc* scene, const RtbCamera* cam, const RtbRenderOptions* options, int n_threads,
               float* accum, uint8_t* rgba, RtbRenderStats* stats);

/* color.toGamma2 + @intFromFloat truncation (src/color.zig:43-62, src/camera.zig:57-65). */
void orc_resolve(const float* accum, uint8_t* rgba, uint64_t n_pixels, float n_samples_override);

/* ---- host-side producers of hot-path inputs, restated for cross-checks ---- */

/* Host PRNG used for scene/BVH/perlin generation (the reference's is unseedable): SplitMix64,
 * randomDouble = (next >> 40) * 2^-24.  state is advanced in place. */
float orc_host_random(uint64_t* state);
/* rtweekend.randomIntRange (src/rtweekend.zig:23-27) incl. its max+1 quirk. */
uint32_t orc_host_random_int_range(uint64_t* state, uint32_t min, uint32_t max);

typedef struct OrcCameraOptions { /* src/camera.zig:70-91 */
    float aspect_ratio;
    uint32_t image_width;
    uint32_t image_height; /* 0 = derive */
    uint32_t samples_per_pixel;
    uint32_t max_depth;
    float background[3];
    float vfov;
    float lookfrom[3];
    float lookat[3];
    float vup[3];
    float defocus_angle;
    float focus_dist;
    uint32_t background_mode;
} OrcCameraOptions;
/* Camera.init (src/camera.zig:118-154) */
void orc_camera_init(const OrcCameraOptions* opt, RtbCamera* cam_out);

/* Bounding boxes: Sphere.init / initMoving (src/objects.zig:80-92; center2 = NULL for a static
 * sphere), Quad.init incl. pad (:206-211, src/aabb.zig:36-43). */
void orc_sphere_bbox(const float center1[3], const float* center2, float radius, float bmin[3], float bmax[3]);
void orc_quad_bbox(const float q[3], const float u[3], const float v[3], float bmin[3], float bmax[3]);
/* Translate(RotateY(createBox(a, b))) : HittableList.add (src/objects.zig:274-277, starts from Aabb{} = the
 * origin), RotateY.init (:360-397), Translate.init (:314-319). */
void orc_box_bbox(const float a[3], const float b[3], float sin_theta, float cos_theta, const float offset[3],
                  float bmin[3], float bmax[3]);

/* BVHTree.constructTree (src/bvh.zig:43-103): random axis, in-place heap sort of the object
 * slice (std.sort.heap restated), median split.  Permutes `hittables` and their boxes
 * (n x {min xyz, max xyz}) in place exactly as the reference permutes world_objects.items,
 * writes 2n-1 nodes in creation order (post-order) and returns the root index.  nodes_out must
 * hold 2*n-1 entries. */
int32_t orc_bvh_build(RtbHittable* hittables, float* boxes, uint32_t n, uint64_t* rng_state, RtbBvhNode* nodes_out);

/* Perlin.init (src/perlin.zig:83-101) with permute (:8-16); the reference's out-of-bounds
 * p[256] access (randomIntRange(0,255) may return 256) is clamped to 255. */
void orc_perlin_init(uint64_t* rng_state, RtbPerlin* out);

#ifdef __cplusplus
}
#endif
#endif
