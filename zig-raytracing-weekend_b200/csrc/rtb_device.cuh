// rtb_device.cuh — device-side building blocks of the path-tracing hot path (sm_100a).
//
// Everything the reference does per ray segment, restated for one GPU thread:
//   Camera.getRay          src/camera.zig:156-180      -> get_ray
//   BVHNode.hit / Aabb.hit src/bvh.zig:122-136, src/aabb.zig:82-114 -> traverse_reference
//   Sphere.hit             src/objects.zig:116-148     -> sphere_root / finish_sphere_hit
//   Material.scatter       src/material.zig:18-106     -> shade
//   Texture.value, Perlin  src/textures.zig:22-123, src/perlin.zig:30-152 -> texture_value
//   RNG                    src/rtweekend.zig:14-16     -> Philox4x32-10 keyed (pixel,sample,segment,block)
//
// This translation unit is compiled with -fmad=false: Zig's float mode is strict, so no
// multiply-add contraction may happen in decision arithmetic (slab test, discriminant, roots,
// front_face) — the nearest-hit index must match the CPU semantics bit for bit.  Division and
// square root are IEEE (nvcc defaults -prec-div=true -prec-sqrt=true, no --use_fast_math).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtb.h"

namespace rtb {

// ------------------------------------------------------------------ device scene layout
// BVH re-laid out in DFS pre-order with skip links ("threaded" tree): visiting order is exactly
// the reference's left-then-right recursion, with no stack.  One node = 2 x float4 (32 B):
//   interior: f0 = {bmin.xyz, bits(KIND_INTERIOR<<30 | skip)}   f1 = {bmax.xyz, 0}
//             hit  -> next = i + 1 (left child), miss -> next = skip (first node after the subtree)
//   sphere  : f0 = {center1.xyz, bits(kind<<30 | object)}       f1 = {center_vec.xyz, radius}
//             next = i + 1 (a leaf's successor in pre-order is the next node)
//   quad    : f0 = {0,0,0, bits(KIND_QUAD<<30 | object)}        f1 = {0,0,0, bits(quad slot)}
enum : uint32_t { KIND_INTERIOR = 0u, KIND_SPHERE = 1u, KIND_MOVING_SPHERE = 2u, KIND_QUAD = 3u };
// Shading classes for hit binning in the wavefront integrator: rays that missed, and hits by material kind.
enum : uint32_t { CLASS_MISS = 0u, CLASS_LAMBERT_SOLID = 1u, CLASS_METAL = 2u, CLASS_DIELECTRIC = 3u, CLASS_OTHER = 4u,
                  kShadeClasses = 5u };
#define RTB_META_INDEX_MASK 0x3fffffffu

// Material record, 2 x float4:
//   m0 = {bits(type | tex_type << 8), bits(texture index), fuzz, ir}
//   m1 = {albedo.rgb (metal) or inlined solid colour (lambertian/light/isotropic with a solid texture), 0}
// Texture record, 3 x float4:
//   t0 = {bits(type), bits(index), scale, 0}   t1 = {color.rgb, 0}   t2 = {color2.rgb, 0}
struct DevImage {
    const uchar4* texels;  // tightly packed RGBA8, row stride = width
    uint32_t width, height;
};
struct DevPerlin {
    float4 ranvec[256];
    uint8_t perm_x[256], perm_y[256], perm_z[256];
};
struct DevQuad {  // Quad fields incl. the ones Quad.init derives (src/objects.zig:195-211)
    float4 q_d;       // q.xyz, d
    float4 u;         // u.xyz
    float4 v;         // v.xyz
    float4 normal;    // normal.xyz
    float4 w;         // w.xyz
};

// One entry per hittable (top-level objects AND the children of wrappers), for the general instancing path (hit_any):
//   sphere / quad / box / constant_medium : slot = quad-table slot of a quad / box header (spheres read prims[])
//   translate : child, p = offset            rotate_y : child, p = (sin_theta, cos_theta)
//   list      : child = first member, count  medium_of : child = boundary, p.x = neg_inv_density
struct DevInst {
    uint32_t type, child, count, slot;
    float4 p;
};
// What a leaf test of a complex object may need; passed by value through the traversal functions.
struct ComplexTables {
    const DevQuad* quads;
    const DevInst* insts;
    const float4* prims;
};

struct DevScene {
    const float4* nodes;  // 2 per node: reference order, bounds as (min, max)
    // Per-octant layouts for the wavefront integrator, [mode][octant][2 * n_nodes], bounds pre-swapped
    // to (entry plane, exit plane) for the octant.  mode 0 = reference order, 1 = near-child-first,
    // 2 = the same objects re-partitioned by the library with a binned-SAH tree, near-child-first.
    // Each octant's array holds oct_n_nodes[mode] + 1 entries: the last one is the end sentinel (RTB_META_END).
    // Modes 0 and 1 have the host tree's n_nodes entries.  Mode 2 is laid out for speed, not for likeness: first the
    // "huge" objects (box >= half of the scene's, e.g. the ground sphere) as plain leaves, then the SAH tree of the
    // rest WITHOUT its root box (the walk starts with the root's two subtrees) and with a BOX node (the object's own
    // padded box, skip = past the leaf) in front of every leaf, so a primitive is only tested when the ray enters
    // its box: h + 3(n - h) - 2 entries for n objects of which h are huge.
    const float4* oct_nodes[3];
    uint32_t oct_n_nodes[3];
    // RTB_TRAVERSAL_SAH16: the mode-2 tree packed into 16-byte SLOTS, [octant][pk_slots] (see traverse_packed):
    //   box node (1 slot)   {half2(entry.x, exit.x), half2(entry.y, exit.y), half2(entry.z, exit.z), skip slot}
    //                       planes in the layout's own normalised frame n = (x - pk_center) * pk_inv_scale, |n| <= 1,
    //                       rounded OUTWARDS to binary16 on the host;
    //   leaf      (2 slots) {center1.xyz, kind|object} {center_vec.xyz, -radius}   (f32, world frame; complex objects:
    //                       {.., kind|object} {bits(subtype), 0, 0, 0x80000000 | quad slot}) — the last word of a leaf
    //                       always has bit 31 set, so "word 3 < 2^30" identifies exactly the box nodes' skip links;
    //   end sentinel (1 slot) word 3 = RTB_META_END.
    // Walked from shared memory when one octant fits, else from global memory; NULL only with RTB_PACK_LARGE=0 for a
    // scene that does not fit (then SAH16 renders as SAH).
    const uint4* pk_nodes;
    uint32_t pk_slots;  // slots per octant, sentinel included
    float pk_center[3], pk_inv_scale[3], pk_scale[3];
    // 4 per object: the leaf record {center1, kind|object}, {center_vec, radius} and the object's material
    // record {m0, m1} inlined (the reference stores Material by value in every Hittable anyway).
    const float4* prims;
    // Shading class of each object (ShadeClass), used to bin hits so that a warp shades one material kind.
    const uint8_t* object_class;
    uint32_t n_nodes;
    uint32_t n_objects;
    const uint32_t* object_material;  // object index -> material index
    const float4* materials;          // 2 per material
    const float4* textures;           // 3 per texture
    const DevPerlin* perlins;
    const DevImage* images;
    const DevQuad* quads;
    const DevInst* insts;  // n_objects entries
    uint32_t has_quads;
    uint32_t n_perlins;
};
__device__ __forceinline__ ComplexTables complex_tables(const DevScene& sc) { return ComplexTables{sc.quads, sc.insts, sc.prims}; }

struct DevCamera {
    float3 center, pixel00, du, dv, ddu, ddv, background;
    float defocus_angle;
    uint32_t width, height, max_depth, background_mode;
};

// ------------------------------------------------------------------ vec3.zig (element-wise, unfused)
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
__device__ __forceinline__ float3 splat3(float s) { return make_float3(s, s, s); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float length_squared(float3 u) { return u.x * u.x + u.y * u.y + u.z * u.z; }
__device__ __forceinline__ float dot3(float3 u, float3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }
__device__ __forceinline__ float3 cross3(float3 u, float3 v) {
    return f3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
__device__ __forceinline__ float3 unit_vector(float3 v) { return v / splat3(sqrtf(length_squared(v))); }
__device__ __forceinline__ float3 reflect3(float3 v, float3 n) { return v - n * splat3(dot3(v, n) * 2.0f); }
__device__ __forceinline__ float3 refract3(float3 uv, float3 n, float etai_over_etat) {
    const float cos_theta = fminf(dot3(-uv, n), 1.0f);
    const float3 r_out_perp = splat3(etai_over_etat) * (uv + n * splat3(cos_theta));
    const float3 r_out_parallel = n * splat3(-sqrtf(fabsf(1.0f - length_squared(r_out_perp))));
    return r_out_perp + r_out_parallel;
}

// ------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

struct RngKey {
    uint2 seed;
    uint32_t pixel, sample;
};
__device__ __forceinline__ float4 rng_block(const RngKey& k, uint32_t segment, uint32_t block) {
    const uint4 o = philox4x32_10(make_uint4(k.pixel, k.sample, segment, block), k.seed);
    return make_float4(u01(o.x), u01(o.y), u01(o.z), u01(o.w));
}
// randomDoubleRange(-1, 1) = min + (max - min) * r   (src/rtweekend.zig:18-20)
__device__ __forceinline__ float range_pm1(float r) { return -1.0f + 2.0f * r; }

// vec3.randomUnitVector (src/vec3.zig:59-68); try j = words 0..2 of block j of this segment.
__device__ __forceinline__ float3 random_unit_vector(const RngKey& k, uint32_t segment, float4 block0) {
    float4 b = block0;
    for (uint32_t j = 0;; ++j) {
        if (j) b = rng_block(k, segment, j);
        const float3 p = f3(range_pm1(b.x), range_pm1(b.y), range_pm1(b.z));
        if (length_squared(p) < 1.0f) return unit_vector(p);
    }
}

// vec3.randomUnitVector for a whole warp at once — the same value for every lane as random_unit_vector, in fewer
// rounds.  The rejection loop accepts with probability pi/6, so the slowest of 32 lanes needs ~5.5 tries while a lane
// needs 1.9 on average, and in the plain loop the lanes that are done idle through those rounds.  Try j of a path is a
// pure function of (seed; pixel, sample, segment, block j), so ANY lane can evaluate it: after the first round the
// lanes are redistributed over the paths still pending — c = 32 / pending (rounded down to a power of two) lanes per
// path evaluate its next c tries side by side, and the path takes the accepted one with the lowest index (the one the
// sequential loop would have stopped at).  Expected rounds per warp: ~3.2 instead of ~5.5.
// Must be called by all 32 lanes (`need` = this lane has a path); `scratch` = 32 uint4 of shared memory owned by the warp.
__device__ __forceinline__ float3 random_unit_vector_coop(bool need, const RngKey& key, uint32_t segment, uint4* scratch) {
    const uint32_t lane = threadIdx.x & 31u;
    float3 p = f3(1.0f, 0.0f, 0.0f);
    bool pending = need;
    if (need) {
        const float4 b = rng_block(key, segment, 0u);
        p = f3(range_pm1(b.x), range_pm1(b.y), range_pm1(b.z));
        pending = !(length_squared(p) < 1.0f);
    }
    uint32_t next_try = 1u;
    for (;;) {
        const uint32_t pend = __ballot_sync(0xffffffffu, pending);
        if (pend == 0u) break;
        const uint32_t n_p = __popc(pend);
        const uint32_t lg = (uint32_t)__clz((int)(n_p - 1u)) - 27u;  // log2 of the lanes per pending path: 5 .. 0
        const uint32_t rank = __popc(pend & ((1u << lane) - 1u));    // of this lane's own path among the pending ones
        if (pending) scratch[rank] = make_uint4(key.pixel, key.sample, next_try, 0u);
        __syncwarp();
        const uint32_t k = lane >> lg, off = lane & ((1u << lg) - 1u);  // this lane evaluates try `off` of pending path k
        float3 q = f3(0.0f, 0.0f, 0.0f);
        bool ok = false;
        if (k < n_p) {
            const uint4 w = scratch[k];
            RngKey kk;
            kk.seed = key.seed;
            kk.pixel = w.x;
            kk.sample = w.y;
            const float4 b = rng_block(kk, segment, w.z + off);
            q = f3(range_pm1(b.x), range_pm1(b.y), range_pm1(b.z));
            ok = length_squared(q) < 1.0f;
        }
        const uint32_t okm = __ballot_sync(0xffffffffu, ok);
        // the lanes that worked for this lane's path: [rank << lg, (rank + 1) << lg)
        const uint32_t group = (lg == 5u ? 0xffffffffu : ((1u << (1u << lg)) - 1u) << (rank << lg));
        const uint32_t won = pending ? (okm & group) : 0u;
        const uint32_t from = won ? (uint32_t)__ffs((int)won) - 1u : lane;
        const float qx = __shfl_sync(0xffffffffu, q.x, from);
        const float qy = __shfl_sync(0xffffffffu, q.y, from);
        const float qz = __shfl_sync(0xffffffffu, q.z, from);
        if (pending) {
            if (won) {
                p = f3(qx, qy, qz);
                pending = false;
            } else {
                next_try += 1u << lg;
            }
        }
        __syncwarp();  // scratch is rewritten in the next round
    }
    return unit_vector(p);
}

// ------------------------------------------------------------------ rays
struct DRay {
    float3 o, d;
    float time;
};

// Camera.getRay (src/camera.zig:169-180) for flat pixel index (x = i % W + 1, y = i / W + 1,
// 1-based: src/camera.zig:100-101).  Stream segment 0: block 0 = (jitter x, jitter y, time, -),
// defocus-disk try j = block 1 + (j >> 1), words 2(j&1), 2(j&1)+1.
__device__ __forceinline__ DRay get_ray(const DevCamera& cam, const RngKey& k, uint32_t x, uint32_t y);
__device__ __forceinline__ DRay get_ray(const DevCamera& cam, const RngKey& k) {
    return get_ray(cam, k, k.pixel % cam.width + 1u, k.pixel / cam.width + 1u);
}
// The same with the 1-based pixel coordinates supplied (a caller that already knows them saves two integer divisions).
__device__ __forceinline__ DRay get_ray(const DevCamera& cam, const RngKey& k, uint32_t x, uint32_t y) {
    const float4 b0 = rng_block(k, 0u, 0u);
    const float3 pixel_center = cam.pixel00 + cam.du * splat3((float)x) + cam.dv * splat3((float)y);
    const float px = -0.5f + b0.x;
    const float py = -0.5f + b0.y;
    const float3 pixel_sample = pixel_center + (splat3(px) * cam.du + splat3(py) * cam.dv);
    float3 origin = cam.center;
    if (!(cam.defocus_angle <= 0.0f)) {
        float px_d, py_d;
        for (uint32_t jb = 1u;; ++jb) {  // tries 2(jb - 1) and 2(jb - 1) + 1 share block jb: generated once
            const float4 b = rng_block(k, 0u, jb);
            px_d = range_pm1(b.x);
            py_d = range_pm1(b.y);
            if (px_d * px_d + py_d * py_d + 0.0f * 0.0f < 1.0f) break;
            px_d = range_pm1(b.z);
            py_d = range_pm1(b.w);
            if (px_d * px_d + py_d * py_d + 0.0f * 0.0f < 1.0f) break;
        }
        origin = cam.center + cam.ddu * splat3(px_d) + cam.ddv * splat3(py_d);
    }
    DRay r;
    r.o = origin;
    r.d = pixel_sample - origin;
    r.time = b0.z;
    return r;
}

// ------------------------------------------------------------------ traversal
struct Nearest {
    float t;        // closest accepted root so far (doubles as ray_t.max)
    uint32_t node;  // pre-order index of the leaf that produced it, 0xffffffff = none
};

// Sphere.hit up to the accepted root (src/objects.zig:121-137).  `center` already includes the
// motion term.  Returns true and the root iff it lies strictly inside (t_min, t_max).
__device__ __forceinline__ bool sphere_root(const DRay& r, float3 center, float radius, float t_min, float t_max,
                                            float& root_out) {
    const float3 oc = r.o - center;
    const float a = length_squared(r.d);
    const float half_b = dot3(oc, r.d);
    const float c = length_squared(oc) - radius * radius;
    const float discriminant = half_b * half_b - a * c;
    if (discriminant < 0.0f) return false;
    const float sqrtd = sqrtf(discriminant);
    float root = (-half_b - sqrtd) / a;
    if (!(t_min < root && root < t_max)) {  // Interval.surrounds, src/interval.zig:12-14
        root = (-half_b + sqrtd) / a;
        if (!(t_min < root && root < t_max)) return false;
    }
    root_out = root;
    return true;
}

// Quad.hit (src/objects.zig:226-261); returns t and the planar coordinates.
__device__ __forceinline__ bool quad_root(const DRay& r, const DevQuad& qd, float t_min, float t_max, float& t_out,
                                          float& alpha_out, float& beta_out) {
    const float3 normal = f3(qd.normal);
    const float denom = dot3(normal, r.d);
    if (fabsf(denom) < 1e-8f) return false;
    const float t = (qd.q_d.w - dot3(normal, r.o)) / denom;
    if (!(t_min <= t && t <= t_max)) return false;  // Interval.contains, src/interval.zig:8-10
    const float3 intersection = r.o + splat3(t) * r.d;
    const float3 planar = intersection - f3(qd.q_d);
    const float alpha = dot3(f3(qd.w), cross3(planar, f3(qd.v)));
    const float beta = dot3(f3(qd.w), cross3(f3(qd.u), planar));
    if ((alpha < 0.0f) || (1.0f < alpha) || (beta < 0.0f) || (1.0f < beta)) return false;
    t_out = t;
    alpha_out = alpha;
    beta_out = beta;
    return true;
}

// 3-input min/max (sm_100a FMNMX3).  PTX min/max return the non-NaN operand, which is exactly
// what the reference's `if (t0 > ray_t_min) ray_t_min = t0` does with a NaN t0: nothing.
__device__ __forceinline__ float max3f(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float min3f(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// Aabb.hit (src/aabb.zig:82-114) == false.  f0.xyz = box min, f1.xyz = box max.
//   * 1/direction is hoisted out of the node loop (the reference recomputes the same IEEE quotient
//     at every node);
//   * the swap `if (invD < 0)` is a select on the per-ray sign;
//   * `if (t0 > tmin) tmin = t0` / `if (t1 < tmax) tmax = t1` are max/min that ignore a NaN t (the
//     running tmin/tmax are never NaN), and the three per-axis early-outs fold into one final
//     `tmax <= tmin`, which is equivalent because tmin only grows and tmax only shrinks.
__device__ __forceinline__ bool slab_miss(float4 f0, float4 f1, float3 o, float inv_x, float inv_y, float inv_z,
                                          float t_min, float t_max) {
    const float ax = (f0.x - o.x) * inv_x, bx = (f1.x - o.x) * inv_x;
    const float ay = (f0.y - o.y) * inv_y, by = (f1.y - o.y) * inv_y;
    const float az = (f0.z - o.z) * inv_z, bz = (f1.z - o.z) * inv_z;
    const bool sx = inv_x < 0.0f, sy = inv_y < 0.0f, sz = inv_z < 0.0f;
    const float t0x = sx ? bx : ax, t1x = sx ? ax : bx;
    const float t0y = sy ? by : ay, t1y = sy ? ay : by;
    const float t0z = sz ? bz : az, t1z = sz ? az : bz;
    const float lo = fmaxf(max3f(t0x, t0y, t0z), t_min);
    const float hi = fminf(min3f(t1x, t1y, t1z), t_max);
    return hi <= lo;
}

// A box instance = Translate(RotateY(createBox(a, b, mat), angle), offset) (src/objects.zig:308-443, :510-532),
// stored in the quad table as a header entry {q_d = (offset.xyz, sin_theta), u.x = cos_theta} followed by
// the 6 quads createBox makes, in its order.  Leaf record of a complex object: f1.x = bits(subtype).
enum : uint32_t { COMPLEX_QUAD = 0u, COMPLEX_BOX = 1u, COMPLEX_MEDIUM = 2u, COMPLEX_GENERIC = 3u };

// Translate.hit (:331-335) then RotateY.hit (:410-421): the ray in the box's own frame.
__device__ __forceinline__ DRay box_local_ray(const DRay& r, const DevQuad& hdr) {
    const float3 offset = f3(hdr.q_d);
    const float sin_theta = hdr.q_d.w, cos_theta = hdr.u.x;
    const float3 mo = r.o - offset;
    DRay l;
    l.o = f3(cos_theta * mo.x - sin_theta * mo.z, mo.y, sin_theta * mo.x + cos_theta * mo.z);
    l.d = f3(cos_theta * r.d.x - sin_theta * r.d.z, r.d.y, sin_theta * r.d.x + cos_theta * r.d.z);
    l.time = r.time;
    return l;
}

// HittableList.hit over the 6 faces (src/objects.zig:286-304): every face is tried with
// ray_t.max = closest so far (inclusive for quads).  Returns the face index of the accepted hit.
__device__ __forceinline__ bool box_root(const DRay& r, const DevQuad* __restrict__ entry, float t_min, float t_max,
                                         float& t_out, uint32_t& face_out, float& alpha_out, float& beta_out) {
    const DRay l = box_local_ray(r, entry[0]);
    bool hit = false;
    float closest = t_max;
#pragma unroll 1
    for (uint32_t f = 0; f < 6u; ++f) {
        float t, alpha, beta;
        if (quad_root(l, entry[1u + f], t_min, closest, t, alpha, beta)) {
            closest = t;
            hit = true;
            t_out = t;
            face_out = f;
            alpha_out = alpha;
            beta_out = beta;
        }
    }
    return hit;
}

// ConstantMedium.hit (src/objects.zig:462-507) with a box instance as the boundary; the table entry is a box
// whose header carries neg_inv_density in u.y.  The one random number it draws (:484) is word 0 of block
// (0x40000000 + object) of the segment's stream — independent of the order in which the tree is walked.
__device__ __forceinline__ bool medium_root(const DRay& r, const DevQuad* __restrict__ entry, float t_min, float t_max,
                                            const RngKey& key, uint32_t segment, uint32_t object, float& t_out) {
    const float inf = __int_as_float(0x7f800000);
    float t1, t2, alpha, beta;
    uint32_t face;
    if (!box_root(r, entry, -inf, inf, t1, face, alpha, beta)) return false;          // intervals.universe
    if (!box_root(r, entry, t1 + 0.0001f, inf, t2, face, alpha, beta)) return false;
    if (t1 < t_min) t1 = t_min;
    if (t2 > t_max) t2 = t_max;
    if (t1 >= t2) return false;
    if (t1 < 0.0f) t1 = 0.0f;
    const float ray_length = sqrtf(length_squared(r.d));
    const float distance_inside_boundary = (t2 - t1) * ray_length;
    const float hit_distance = entry[0].u.y * logf(rng_block(key, segment, 0x40000000u + object).x);
    if (hit_distance > distance_inside_boundary) return false;
    t_out = t1 + hit_distance / ray_length;
    return true;
}

// The general instancing path (src/objects.zig:264-443): Translate / RotateY / HittableList / ConstantMedium wrapping
// ANY hittable, evaluated recursively like the reference's `inline else => |object| object.hit(r, ray_t)`.
struct AnyHit {
    float3 p, normal;
    float t, u, v;
    uint32_t prim;  // the hittable whose material shades this hit
    bool front_face;
};
static __device__ __noinline__ bool hit_any(ComplexTables ct, uint32_t idx, DRay r, float t_min, float t_max, AnyHit* h,
                                            RngKey key, uint32_t segment);

// Leaf test of a complex object (quad, box instance, constant medium, or a general wrapper);
// f1 = {bits(subtype), -, -, bits(slot)} (slot = the object's own index for COMPLEX_GENERIC).
__device__ __forceinline__ bool complex_root(const DRay& r, ComplexTables ct, float4 f1, float t_min,
                                             float t_max, float& t_out, const RngKey& key, uint32_t segment,
                                             uint32_t object) {
    const DevQuad* __restrict__ entry = ct.quads + __float_as_uint(f1.w);
    const uint32_t subtype = __float_as_uint(f1.x);
    float alpha, beta;
    if (subtype == COMPLEX_BOX) {
        uint32_t face;
        return box_root(r, entry, t_min, t_max, t_out, face, alpha, beta);
    }
    if (subtype == COMPLEX_MEDIUM) return medium_root(r, entry, t_min, t_max, key, segment, object, t_out);
    if (subtype == COMPLEX_GENERIC) {
        AnyHit h;
        if (!hit_any(ct, object, r, t_min, t_max, &h, key, segment)) return false;
        t_out = h.t;
        return true;
    }
    return quad_root(r, entry[0], t_min, t_max, t_out, alpha, beta);
}

// Octant of a ray = the three `invD < 0` predicates of Aabb.hit (src/aabb.zig:97), bit k = axis k.
__device__ __forceinline__ uint32_t ray_octant(float inv_x, float inv_y, float inv_z) {
    return (inv_x < 0.0f ? 1u : 0u) | (inv_y < 0.0f ? 2u : 0u) | (inv_z < 0.0f ? 4u : 0u);
}

// The same three predicates from the direction itself, without the divisions: 1/d < 0  <=>  d is negative, or -0
// (1/-0 = -inf), but neither -inf (1/-inf = -0, not < 0) nor NaN  <=>  0x80000000 <= bits(d) < 0xff800000.
__device__ __forceinline__ uint32_t ray_octant_of_direction(float3 d) {
    const uint32_t x = __float_as_uint(d.x) - 0x80000000u, y = __float_as_uint(d.y) - 0x80000000u,
                   z = __float_as_uint(d.z) - 0x80000000u;
    return (x < 0x7f800000u ? 1u : 0u) | (y < 0x7f800000u ? 2u : 0u) | (z < 0x7f800000u ? 4u : 0u);
}

// Slab test against a node of a PER-OCTANT layout: f0.xyz already holds the plane the ray enters
// through on each axis (min, or max where invD < 0) and f1.xyz the one it leaves through, so the
// reference's swap is done once per (node, octant) on the host instead of per visit.  The values
// t0/t1 are the same IEEE results as in slab_miss.
__device__ __forceinline__ bool slab_miss_preswapped(float4 f0, float4 f1, float3 o, float inv_x, float inv_y,
                                                     float inv_z, float t_min, float t_max) {
    const float t0x = (f0.x - o.x) * inv_x, t1x = (f1.x - o.x) * inv_x;
    const float t0y = (f0.y - o.y) * inv_y, t1y = (f1.y - o.y) * inv_y;
    const float t0z = (f0.z - o.z) * inv_z, t1z = (f1.z - o.z) * inv_z;
    const float lo = fmaxf(max3f(t0x, t0y, t0z), t_min);
    const float hi = fminf(min3f(t1x, t1y, t1z), t_max);
    return hi <= lo;
}

// RTB_TRAVERSAL_SAH only (the library's own tree, whose boxes are padded on the host by more than the
// rounding difference, see build_sah): the same slab test as one FMA per plane, t = plane * invD - origin * invD.
// It may only ever ADMIT more nodes than the exact test would; leaves are still tested with the reference's
// arithmetic, so hit values stay bit-identical.  A NaN (0 * inf) leaves its axis unconstrained.
__device__ __forceinline__ bool slab_miss_fma(float4 f0, float4 f1, float inv_x, float inv_y, float inv_z, float nox,
                                              float noy, float noz, float t_min, float t_max) {
    const float t0x = __fmaf_rn(f0.x, inv_x, nox), t1x = __fmaf_rn(f1.x, inv_x, nox);
    const float t0y = __fmaf_rn(f0.y, inv_y, noy), t1y = __fmaf_rn(f1.y, inv_y, noy);
    const float t0z = __fmaf_rn(f0.z, inv_z, noz), t1z = __fmaf_rn(f1.z, inv_z, noz);
    const float lo = fmaxf(max3f(t0x, t0y, t0z), t_min);
    const float hi = fminf(min3f(t1x, t1y, t1z), t_max);
    return hi <= lo;
}

// Sphere.hit up to the accepted root with a = |direction|^2 hoisted out of the node loop (the
// reference recomputes the same value at every leaf, src/objects.zig:124).
__device__ __forceinline__ bool sphere_root_a(float3 o, float3 d, float a, float3 center, float radius, float t_min,
                                              float t_max, float& root_out) {
    const float3 oc = o - center;
    const float half_b = dot3(oc, d);
    const float c = length_squared(oc) - radius * radius;
    const float discriminant = half_b * half_b - a * c;
    if (discriminant < 0.0f) return false;
    const float sqrtd = sqrtf(discriminant);
    float root = (-half_b - sqrtd) / a;
    if (!(t_min < root && root < t_max)) {
        root = (-half_b + sqrtd) / a;
        if (!(t_min < root && root < t_max)) return false;
    }
    root_out = root;
    return true;
}

// Packed FP32 pairs (sm_100: FADD2 / FMUL2 / FFMA2 on a 64-bit register pair; each half rounds like the scalar op).
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// Traversal over one octant's threaded layout (DevScene::oct_nodes).  Visiting order is whatever the
// host baked into the layout: the reference's left-then-right order (RTB_TRAVERSAL_REFERENCE) or
// near-child-first for this octant (RTB_TRAVERSAL_ORDERED).  Nearest.node is the OBJECT index.
// The per-octant layouts end with a SENTINEL node (meta = RTB_META_END) at index n_nodes, the target of
// every skip link that leaves the tree, so the hot slab path carries no loop-bound test.  With SMEM the
// node array is the kernel's dynamic shared memory (indexed directly, so the LDS address is base + i*32
// without a generic-pointer conversion per iteration).
#define RTB_META_END 0xffffffffu
extern __shared__ float4 rtb_smem_nodes[];

template <bool COUNT, bool QUADS, bool SMEM, bool FMA = false>
__device__ __forceinline__ Nearest traverse_octant(const float4* __restrict__ nodes, ComplexTables quads,
                                                   float3 o, float3 d, float time, float inv_x, float inv_y,
                                                   float inv_z, float t_min, float t_max, uint32_t& n_box,
                                                   uint32_t& n_obj, uint32_t smem_base = 0u, RngKey key = RngKey{},
                                                   uint32_t segment = 1u) {
    // smem_base (SMEM only): 32-bit shared-window address of the staged layout.  The caller adds a
    // run-time zero to it so that ptxas keeps it in a register instead of rematerialising the window
    // base (S2UR/UMOV/UIADD3/ULEA) in every iteration of the node loop.
    Nearest best;
    best.t = t_max;
    best.node = 0xffffffffu;
    const float a = length_squared(d);
    const float nox = -(o.x * inv_x), noy = -(o.y * inv_y), noz = -(o.z * inv_z);  // FMA only
    // SMEM: `i` is the node's 32-bit shared-window ADDRESS and the skip links of the staged copy are addresses too
    // (the staging loop rewrites them, see wf_extend), which saves the index -> address instruction of every visit.
    uint32_t i = SMEM ? smem_base : 0u;
    // SMEM: the staged copy holds a node as (entry, exit) PAIRS per axis — {e.x, x.x, e.y, x.y} {e.z, x.z, meta, f1.w}
    // (wf_extend shuffles while staging) — so that the six plane operations of a visit are three packed FP32
    // instructions (sm_100 FADD2 / FMUL2 / FFMA2 on 64-bit register pairs) instead of six scalar ones.  Each half is the
    // same IEEE round-to-nearest operation as the scalar form, and plane + (-o) is plane - o bit for bit, so the exact
    // variant stays the reference's arithmetic; the kernel is bound by issue slots, not by the FP32 pipe.
    const uint64_t IX = pack_f32x2(inv_x, inv_x), IY = pack_f32x2(inv_y, inv_y), IZ = pack_f32x2(inv_z, inv_z);
    const uint64_t AX = FMA ? pack_f32x2(nox, nox) : pack_f32x2(-o.x, -o.x);
    const uint64_t AY = FMA ? pack_f32x2(noy, noy) : pack_f32x2(-o.y, -o.y);
    const uint64_t AZ = FMA ? pack_f32x2(noz, noz) : pack_f32x2(-o.z, -o.z);
    for (;;) {
        float4 f0, f1;
        uint64_t px = 0, py = 0, pz = 0;
        if (SMEM) {
            uint64_t mw;
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(px), "=l"(py) : "r"(i));
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+16];" : "=l"(pz), "=l"(mw) : "r"(i));
            unpack_f32x2(px, f0.x, f1.x);
            unpack_f32x2(py, f0.y, f1.y);
            unpack_f32x2(pz, f0.z, f1.z);
            unpack_f32x2(mw, f0.w, f1.w);
        } else {
            f0 = nodes[2u * i];
            f1 = nodes[2u * i + 1u];
        }
        constexpr uint32_t kStep = SMEM ? 32u : 1u;
        const uint32_t meta = __float_as_uint(f0.w);
        if (meta < (1u << 30)) {  // KIND_INTERIOR: meta is the skip link
            if (COUNT) ++n_box;
            bool miss;
            if (SMEM) {
                float t0x, t1x, t0y, t1y, t0z, t1z;
                if (FMA) {
                    unpack_f32x2(fma_f32x2(px, IX, AX), t0x, t1x);
                    unpack_f32x2(fma_f32x2(py, IY, AY), t0y, t1y);
                    unpack_f32x2(fma_f32x2(pz, IZ, AZ), t0z, t1z);
                } else {
                    unpack_f32x2(mul_f32x2(add_f32x2(px, AX), IX), t0x, t1x);
                    unpack_f32x2(mul_f32x2(add_f32x2(py, AY), IY), t0y, t1y);
                    unpack_f32x2(mul_f32x2(add_f32x2(pz, AZ), IZ), t0z, t1z);
                }
                const float lo = fmaxf(max3f(t0x, t0y, t0z), t_min);
                const float hi = fminf(min3f(t1x, t1y, t1z), best.t);
                miss = hi <= lo;
            } else {
                miss = FMA ? slab_miss_fma(f0, f1, inv_x, inv_y, inv_z, nox, noy, noz, t_min, best.t)
                           : slab_miss_preswapped(f0, f1, o, inv_x, inv_y, inv_z, t_min, best.t);
            }
            i = miss ? meta : i + kStep;
        } else {
            if (meta == RTB_META_END) break;
            if (COUNT) ++n_obj;
            if (SMEM) {
                // the leaf's words are fetched again here: the slab path overwrites the pairs in place, and keeping a
                // copy for this (rare) branch would cost three moves in every iteration of the node loop
                uint64_t qx, qy, qz, qw;
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(qx), "=l"(qy) : "r"(i));
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+16];" : "=l"(qz), "=l"(qw) : "r"(i));
                unpack_f32x2(qx, f0.x, f1.x);
                unpack_f32x2(qy, f0.y, f1.y);
                unpack_f32x2(qz, f0.z, f1.z);
                unpack_f32x2(qw, f0.w, f1.w);
            }
            const uint32_t kind = meta >> 30;
            if (!QUADS || kind != KIND_QUAD) {
                const float3 c1 = f3(f0);
                const float3 center = (kind == KIND_MOVING_SPHERE) ? c1 + splat3(time) * f3(f1) : c1;
                float root;
                if (sphere_root_a(o, d, a, center, f1.w, t_min, best.t, root)) {
                    best.t = root;
                    best.node = meta & RTB_META_INDEX_MASK;
                }
            } else {
                DRay r;
                r.o = o;
                r.d = d;
                r.time = time;
                float t;
                if (complex_root(r, quads, f1, t_min, best.t, t, key, segment, meta & RTB_META_INDEX_MASK)) {
                    best.t = t;
                    best.node = meta & RTB_META_INDEX_MASK;
                }
            }
            i = i + kStep;
        }
    }
    return best;
}

// ------------------------------------------------------------------ RTB_TRAVERSAL_SAH16: packed half2 box tests
// The box test of the library's own SAH tree only has to be CONSERVATIVE (never reject a box the ray enters before
// its current nearest hit): leaves keep the reference's f32 arithmetic, so every hit value stays bit-identical.  That
// freedom is spent on the two things ncu shows wf_extend is bound by (profiles/r2a_*: l1tex 82-88 % busy on the node
// fetches, 18 issue slots per visit): a box node is ONE 16-byte slot (one LDS.128 instead of two) and the slab test is
// three HFMA2 on (entry, exit) plane pairs, two packed max and one packed compare — 12 instructions per visit.
//
// Per axis, with planes e (entry) and x (exit) in the normalised frame and the ray's own per-axis constants
//   (t_entry, -t_exit) = (e, x) * (I_lo, I_hi) + (N_lo, N_hi)            one HFMA2
//   I_lo =  inv * (1 - k)    N_lo =  nod * (1 - k) - E
//   I_hi = -inv * (1 + k)    N_hi = -nod * (1 + k) - E       inv = 1 / d_n,  nod = -o_n * inv  (f32, then rounded to f16)
// k = 1.01 * 2^-11 covers the final rounding of the HFMA2 (|t| only ever shrinks for the entry and grows for the exit),
// E = 1.01 * 2^-11 * (|inv| + |nod|) covers the roundings of I and N themselves (|e|, |x| <= 1).  A ray whose
// |inv| + |nod| would not fit binary16 is scaled as a whole by a power of two (t' = sigma * t: roundings are relative, so
// nothing else changes); an axis with d = 0 (1/d infinite) or a NaN term constrains nothing.
// So computed t_entry <= true t_entry and computed t_exit >= true t_exit, for every ray: the walk visits a SUPERSET
// of the nodes an exact test on the same boxes would.  In space the slack is ~2^-11 (1 + |o_n|) of the scene's
// half-extent plus 2^-11 of the distance travelled: on Book-1 ~0.02 units against leaf boxes of 0.4-1.0.
struct PackedRay {
    uint32_t ix, iy, iz, nx, ny, nz;  // half2 bit patterns (lo half: entry, hi half: exit)
    float sigma;                      // power of two all t values of this ray are scaled by to fit binary16
};
__device__ __forceinline__ __half2 as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h2_bits(__half lo, __half hi) {
    const __half2 h = __halves2half2(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// inv = 1 / d_n and nod = -o_n / d_n of one axis in the layout's normalised frame, and |inv| + |nod|
__device__ __forceinline__ void packed_axis_terms(float o, float d, float center, float inv_scale, float scale, float& inv,
                                                  float& nod, float& mag) {
    const float on = (o - center) * inv_scale;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));  // binary16 precision is all the test keeps
    inv = scale * r;                                       // 1 / d_n with d_n = d * inv_scale
    nod = -(on * inv);
    mag = fabsf(inv) + fabsf(nod);
}
__device__ __forceinline__ void packed_axis_pack(float inv, float nod, float mag, float sigma, uint32_t& I, uint32_t& N) {
    if (!(mag < 3.0e38f)) {  // d = 0 (1/d = inf) or NaN (0 * inf): this axis constrains nothing
        I = 0u;
        N = 0xfc00fc00u;     // (-inf, -inf): t_entry = -inf, t_exit = +inf
        return;
    }
    const float k = 1.01f / 2048.0f;
    const float E = mag * (1.01f / 2048.0f) + 1e-7f;
    I = h2_bits(__float2half_rn(sigma * (inv * (1.0f - k))), __float2half_rn(sigma * -(inv * (1.0f + k))));
    N = h2_bits(__float2half_rn(sigma * (nod * (1.0f - k) - E)), __float2half_rn(sigma * (-(nod * (1.0f + k)) - E)));
}
// binary16 holds |x| < 65504: a ray far from the scene (|o_n| large: a bounce off the ground sphere thousands of units
// away) or nearly parallel to an axis has |inv| + |nod| beyond that.  Rounding errors are relative, so the whole ray is
// simply scaled by a power of two: t' = sigma * t, with the interval (t_min, t_max) scaled alike.
__device__ __forceinline__ float packed_sigma(float mx, float my, float mz) {
    float m = 0.0f;
    if (mx < 3.0e38f) m = fmaxf(m, mx);
    if (my < 3.0e38f) m = fmaxf(m, my);
    if (mz < 3.0e38f) m = fmaxf(m, mz);
    const int e = (int)(__float_as_uint(m) >> 23) - 127;                 // floor(log2 m), m >= 0 finite
    const int kk = e > 12 ? (e - 12 > 120 ? 120 : e - 12) : 0;           // keep sigma * m < 2^13
    return __uint_as_float((uint32_t)(127 - kk) << 23);
}
__device__ __forceinline__ PackedRay packed_ray_setup(const DevScene& sc, float3 o, float3 d) {
    PackedRay pr;
    float ix, iy, iz, nx, ny, nz, mx, my, mz;
    packed_axis_terms(o.x, d.x, sc.pk_center[0], sc.pk_inv_scale[0], sc.pk_scale[0], ix, nx, mx);
    packed_axis_terms(o.y, d.y, sc.pk_center[1], sc.pk_inv_scale[1], sc.pk_scale[1], iy, ny, my);
    packed_axis_terms(o.z, d.z, sc.pk_center[2], sc.pk_inv_scale[2], sc.pk_scale[2], iz, nz, mz);
    pr.sigma = packed_sigma(mx, my, mz);
    packed_axis_pack(ix, nx, mx, pr.sigma, pr.ix, pr.nx);
    packed_axis_pack(iy, ny, my, pr.sigma, pr.iy, pr.ny);
    packed_axis_pack(iz, nz, mz, pr.sigma, pr.iz, pr.nz);
    return pr;
}
// (sigma * t_min rounded down, -(sigma * t_max rounded up)) as half2: the fourth operand of the packed max
__device__ __forceinline__ __half2 packed_interval(float t_min, float t_max, float sigma) {
    return __halves2half2(__float2half_rd(sigma * t_min), __hneg(__float2half_ru(sigma * t_max)));
}

// Walk over one octant's packed layout.  SMEM: `smem_base` is the shared-window address of the staged copy, whose skip
// links are byte distances from the node that holds them (rewritten while staging); otherwise `slots` is the octant's
// array and the links are slot indices.
template <bool COUNT, bool QUADS, bool SMEM>
__device__ __forceinline__ Nearest traverse_packed(const uint4* __restrict__ slots, ComplexTables quads,
                                                   float3 o, float3 d, float time, const PackedRay& pr, float t_min,
                                                   float t_max, uint32_t& n_box, uint32_t& n_obj,
                                                   uint32_t smem_base = 0u, RngKey key = RngKey{},
                                                   uint32_t segment = 1u) {
    Nearest best;
    best.t = t_max;
    best.node = 0xffffffffu;
    const float a = length_squared(d);
    const __half2 ix = as_h2(pr.ix), iy = as_h2(pr.iy), iz = as_h2(pr.iz);
    const __half2 nx = as_h2(pr.nx), ny = as_h2(pr.ny), nz = as_h2(pr.nz);
    __half2 K = packed_interval(t_min, t_max, pr.sigma);
    constexpr uint32_t kStep = SMEM ? 16u : 1u;
    uint32_t i = SMEM ? smem_base : 0u;
    // Written as "walk box nodes until a leaf or the end, then handle it" with the node fetch at the BOTTOM of the inner
    // loop, which is how the warp executes it anyway (the compiler reconverges at the leaf): one branch per visit
    // instead of two.
    auto fetch = [&](uint32_t at) {
        uint4 v;
        if (SMEM) {
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(at));
        } else {
            v = slots[at];
        }
        return v;
    };
    uint4 n = fetch(i);
    for (;;) {
        while (n.w < (1u << 30)) {  // box node: word 3 is the skip link
            if (COUNT) ++n_box;
            const __half2 tx = __hfma2(as_h2(n.x), ix, nx);
            const __half2 ty = __hfma2(as_h2(n.y), iy, ny);
            const __half2 tz = __hfma2(as_h2(n.z), iz, nz);
            const __half2 r = __hmax2(__hmax2(tx, ty), __hmax2(tz, K));  // (t_entry, -t_exit); NaN operands are ignored
            const bool miss = __hge(__high2half(r), __hneg(__low2half(r)));  // t_exit <= t_entry
            // SMEM: the staged links are byte distances from the node itself, so the walk adds either the link or one
            // slot to its address (SEL + IADD on a register that is not part of the fetched quad)
            if (SMEM) i += miss ? n.w : kStep;
            else      i = miss ? n.w : i + kStep;
            n = fetch(i);
        }
        if (n.w == RTB_META_END) break;
        if (COUNT) ++n_obj;
        const uint4 m = fetch(i + kStep);
        const uint32_t kind = n.w >> 30;
        bool hit;
        float root;
        if (!QUADS || kind != KIND_QUAD) {
            const float3 c1 = f3(__uint_as_float(n.x), __uint_as_float(n.y), __uint_as_float(n.z));
            const float3 cv = f3(__uint_as_float(m.x), __uint_as_float(m.y), __uint_as_float(m.z));
            const float3 center = (kind == KIND_MOVING_SPHERE) ? c1 + splat3(time) * cv : c1;
            // word 3 holds -radius (bit 31 marks a leaf's second slot); only radius^2 is used here
            hit = sphere_root_a(o, d, a, center, __uint_as_float(m.w), t_min, best.t, root);
        } else {
            DRay r;
            r.o = o;
            r.d = d;
            r.time = time;
            const float4 f1 = make_float4(__uint_as_float(m.x), 0.0f, 0.0f, __uint_as_float(m.w & 0x7fffffffu));
            hit = complex_root(r, quads, f1, t_min, best.t, root, key, segment, n.w & RTB_META_INDEX_MASK);
        }
        if (hit) {
            best.t = root;
            best.node = n.w & RTB_META_INDEX_MASK;
            K = packed_interval(t_min, root, pr.sigma);
        }
        i = i + 2u * kStep;
        n = fetch(i);
    }
    return best;
}

// BVHTree.hit in the reference's visiting order (src/bvh.zig:122-136) over the threaded layout.
//   * interior node: Aabb.hit with ray_t = (t_min, closest so far)  (src/aabb.zig:82-114);
//     1/direction is hoisted out of the loop (the reference recomputes the same IEEE quotient at
//     every node); the per-axis early-outs are folded into one final test, which is equivalent
//     because t_min only grows and t_max only shrinks and NaNs never enter them.
//   * leaf: primitive test directly, no box test.
//   * a later hit replaces an earlier one only if strictly nearer (Interval.surrounds is strict),
//     so ties go to the DFS-earlier object, as in `hit_record_right orelse hit_record_left`.
template <bool COUNT, bool QUADS>
__device__ __forceinline__ Nearest traverse_reference(const float4* __restrict__ nodes, uint32_t n_nodes,
                                                      ComplexTables quads, const DRay& r, float t_min,
                                                      float t_max, uint32_t& n_box, uint32_t& n_obj,
                                                      RngKey key = RngKey{}, uint32_t segment = 1u) {
    Nearest best;
    best.t = t_max;
    best.node = 0xffffffffu;
    const float inv_x = 1.0f / r.d.x, inv_y = 1.0f / r.d.y, inv_z = 1.0f / r.d.z;
    uint32_t i = 0;
    while (i < n_nodes) {
        const float4 f0 = nodes[2u * i];
        const float4 f1 = nodes[2u * i + 1u];
        const uint32_t meta = __float_as_uint(f0.w);
        const uint32_t kind = meta >> 30;
        if (kind == KIND_INTERIOR) {
            if (COUNT) ++n_box;
            i = slab_miss(f0, f1, r.o, inv_x, inv_y, inv_z, t_min, best.t) ? (meta & RTB_META_INDEX_MASK) : i + 1u;
        } else {
            if (COUNT) ++n_obj;
            if (!QUADS || kind != KIND_QUAD) {
                const float3 c1 = f3(f0);
                const float3 center = (kind == KIND_MOVING_SPHERE) ? c1 + splat3(r.time) * f3(f1) : c1;
                float root;
                if (sphere_root(r, center, f1.w, t_min, best.t, root)) {
                    best.t = root;
                    best.node = i;
                }
            } else {
                float t;
                if (complex_root(r, quads, f1, t_min, best.t, t, key, segment, meta & RTB_META_INDEX_MASK)) {
                    best.t = t;
                    best.node = i;
                }
            }
            i = i + 1u;
        }
    }
    return best;
}

// ------------------------------------------------------------------ hit record (src/objects.zig:21-37)
struct DHit {
    float3 p, normal;
    float t, u, v;
    uint32_t object;
    bool front_face;
};

// getSphereUV (src/objects.zig:101-114)
__device__ __forceinline__ void sphere_uv(float3 p, float& u, float& v) {
    const float pi = 3.1415926535897932385f;
    const float theta = acosf(-p.y);
    const float phi = atan2f(-p.z, p.x) + pi;
    u = phi / (2.0f * pi);
    v = theta / pi;
}

// The part of Sphere.hit / Quad.hit after the root is accepted (src/objects.zig:139-147, :250-260),
// evaluated once for the nearest hit instead of once per candidate.
template <bool QUADS, bool WANT_UV>
__device__ __forceinline__ DHit finish_hit_rec(float4 f0, float4 f1, ComplexTables ct, const DRay& r,
                                               float t, const RngKey& key = RngKey{}, uint32_t segment = 1u,
                                               uint32_t* prim_out = nullptr) {
    const DevQuad* __restrict__ quads = ct.quads;
    DHit h;
    const uint32_t meta = __float_as_uint(f0.w);
    const uint32_t kind = meta >> 30;
    h.object = meta & RTB_META_INDEX_MASK;
    if (prim_out) *prim_out = h.object;
    if (QUADS && kind == KIND_QUAD && __float_as_uint(f1.x) == COMPLEX_GENERIC) {
        // a wrapper: evaluate the object again for its nearest hit (the same arithmetic gives the same t) and take
        // the record the reference builds on the way out of the recursion
        AnyHit a;
        a.p = a.normal = f3(0.0f, 0.0f, 0.0f);
        a.t = t; a.u = a.v = 0.0f; a.prim = h.object; a.front_face = false;
        hit_any(ct, h.object, r, 0.001f, __int_as_float(0x7f800000), &a, key, segment);
        h.t = a.t; h.p = a.p; h.normal = a.normal; h.u = a.u; h.v = a.v; h.front_face = a.front_face;
        if (prim_out) *prim_out = a.prim;
        return h;
    }
    h.t = t;
    h.p = r.o + splat3(t) * r.d;  // Ray.at, src/ray.zig:9-11
    h.u = 0.0f;
    h.v = 0.0f;
    float3 outward;
    if (!QUADS || kind != KIND_QUAD) {
        const float3 c1 = f3(f0);
        const float3 center = (kind == KIND_MOVING_SPHERE) ? c1 + splat3(r.time) * f3(f1) : c1;
        outward = (h.p - center) / splat3(f1.w);
        if (WANT_UV) sphere_uv(outward, h.u, h.v);
    } else if (__float_as_uint(f1.x) == COMPLEX_MEDIUM) {
        h.normal = f3(1.0f, 0.0f, 0.0f);  // "arbitrary" (src/objects.zig:493-494)
        h.front_face = true;
        return h;
    } else if (__float_as_uint(f1.x) == COMPLEX_BOX) {
        // Re-run the list test to learn which face was hit (same arithmetic, same t), build the record in the
        // box's frame (Quad.hit :250-260), then rotate back (RotateY.hit :425-439) and translate (:338-342).
        const DevQuad* __restrict__ entry = quads + __float_as_uint(f1.w);
        const DRay l = box_local_ray(r, entry[0]);
        // The list keeps the LAST face in createBox order whose root equals the final t (quads accept
        // t == ray_t.max), so that is the face to rebuild; Interval [t, t] selects exactly those.
        float alpha = 0.0f, beta = 0.0f;
        uint32_t face = 0;
#pragma unroll 1
        for (uint32_t f = 0; f < 6u; ++f) {
            float tt, a2, b2;
            if (quad_root(l, entry[1u + f], t, t, tt, a2, b2)) {
                face = f;
                alpha = a2;
                beta = b2;
            }
        }
        const float3 lp = l.o + splat3(t) * l.d;
        const float3 qn = f3(entry[1u + face].normal);
        const bool front = dot3(l.d, qn) < 0.0f;
        const float3 ln = front ? qn : -qn;
        const float sin_theta = entry[0].q_d.w, cos_theta = entry[0].u.x;
        const float3 offset = f3(entry[0].q_d);
        h.p = f3(cos_theta * lp.x + sin_theta * lp.z, lp.y, -sin_theta * lp.x + cos_theta * lp.z) + offset;
        h.normal = f3(cos_theta * ln.x + sin_theta * ln.z, ln.y, -sin_theta * ln.x + cos_theta * ln.z);
        h.front_face = front;
        h.u = alpha;
        h.v = beta;
        return h;
    } else {
        const DevQuad& qd = quads[__float_as_uint(f1.w)];
        outward = f3(qd.normal);
        if (WANT_UV) {  // isInterior, src/objects.zig:217-224
            const float3 planar = h.p - f3(qd.q_d);
            h.u = dot3(f3(qd.w), cross3(planar, f3(qd.v)));
            h.v = dot3(f3(qd.w), cross3(f3(qd.u), planar));
        }
    }
    h.front_face = dot3(r.d, outward) < 0.0f;  // setFaceNormal, src/objects.zig:30-36
    h.normal = h.front_face ? outward : -outward;
    return h;
}

// Same, with the leaf record fetched from a node / prim array (2 float4 per entry).
template <bool QUADS, bool WANT_UV>
__device__ __forceinline__ DHit finish_hit(const float4* __restrict__ nodes, ComplexTables ct,
                                           const DRay& r, Nearest best, const RngKey& key = RngKey{},
                                           uint32_t segment = 1u) {
    return finish_hit_rec<QUADS, WANT_UV>(nodes[2u * best.node], nodes[2u * best.node + 1u], ct, r, best.t, key, segment);
}

// hit_any: Hittable.hit on any entry of the hittable table (src/objects.zig:49-53), recursive like the reference.
// Every level follows its reference function; t values are the primitives' own (the transforms do not scale t).
static __device__ __noinline__ bool hit_any(ComplexTables ct, uint32_t idx, DRay r, float t_min, float t_max, AnyHit* h,
                                            RngKey key, uint32_t segment) {
    const DevInst in = ct.insts[idx];
    const float inf = __int_as_float(0x7f800000);
    switch (in.type) {
        case RTB_HITTABLE_SPHERE: {  // Sphere.hit, :116-148
            const float4 f0 = ct.prims[4u * (size_t)idx], f1 = ct.prims[4u * (size_t)idx + 1u];
            const uint32_t kind = __float_as_uint(f0.w) >> 30;
            const float3 c1 = f3(f0);
            const float3 center = (kind == KIND_MOVING_SPHERE) ? c1 + splat3(r.time) * f3(f1) : c1;
            float root;
            if (!sphere_root(r, center, f1.w, t_min, t_max, root)) return false;
            const DHit d = finish_hit_rec<false, true>(f0, f1, ct, r, root);
            h->t = root; h->p = d.p; h->normal = d.normal; h->u = d.u; h->v = d.v; h->front_face = d.front_face;
            h->prim = idx;
            return true;
        }
        case RTB_HITTABLE_QUAD: {  // Quad.hit, :226-261
            const DevQuad qd = ct.quads[in.slot];
            float t, alpha, beta;
            if (!quad_root(r, qd, t_min, t_max, t, alpha, beta)) return false;
            h->t = t;
            h->p = r.o + splat3(t) * r.d;
            h->u = alpha;
            h->v = beta;
            const float3 n = f3(qd.normal);
            h->front_face = dot3(r.d, n) < 0.0f;
            h->normal = h->front_face ? n : -n;
            h->prim = idx;
            return true;
        }
        case RTB_HITTABLE_BOX: {  // the one-record box instance: same code as a top-level box
            float t, alpha, beta;
            uint32_t face;
            if (!box_root(r, ct.quads + in.slot, t_min, t_max, t, face, alpha, beta)) return false;
            const float4 f0 = ct.prims[4u * (size_t)idx], f1 = ct.prims[4u * (size_t)idx + 1u];
            const DHit d = finish_hit_rec<true, true>(f0, f1, ct, r, t);
            h->t = t; h->p = d.p; h->normal = d.normal; h->u = d.u; h->v = d.v; h->front_face = d.front_face;
            h->prim = idx;
            return true;
        }
        case RTB_HITTABLE_CONSTANT_MEDIUM: {
            float t;
            if (!medium_root(r, ct.quads + in.slot, t_min, t_max, key, segment, idx, t)) return false;
            h->t = t; h->p = r.o + splat3(t) * r.d; h->normal = f3(1.0f, 0.0f, 0.0f); h->u = h->v = 0.0f;
            h->front_face = true;
            h->prim = idx;
            return true;
        }
        case RTB_HITTABLE_TRANSLATE: {  // Translate.hit, :327-345
            const float3 offset = f3(in.p);
            DRay moved = r;
            moved.o = r.o - offset;
            if (!hit_any(ct, in.child, moved, t_min, t_max, h, key, segment)) return false;
            h->p = h->p + offset;
            return true;
        }
        case RTB_HITTABLE_ROTATE_Y: {  // RotateY.hit, :404-442
            const float sin_theta = in.p.x, cos_theta = in.p.y;
            DRay rot = r;
            rot.o = f3(cos_theta * r.o.x - sin_theta * r.o.z, r.o.y, sin_theta * r.o.x + cos_theta * r.o.z);
            rot.d = f3(cos_theta * r.d.x - sin_theta * r.d.z, r.d.y, sin_theta * r.d.x + cos_theta * r.d.z);
            if (!hit_any(ct, in.child, rot, t_min, t_max, h, key, segment)) return false;
            const float3 p = h->p, n = h->normal;
            h->p = f3(cos_theta * p.x + sin_theta * p.z, p.y, -sin_theta * p.x + cos_theta * p.z);
            h->normal = f3(cos_theta * n.x + sin_theta * n.z, n.y, -sin_theta * n.x + cos_theta * n.z);
            return true;
        }
        case RTB_HITTABLE_LIST: {  // HittableList.hit, :286-304
            bool hit = false;
            float closest = t_max;
            for (uint32_t k = 0; k < in.count; ++k) {
                AnyHit c;
                if (hit_any(ct, in.child + k, r, t_min, closest, &c, key, segment)) {
                    closest = c.t;
                    *h = c;
                    hit = true;
                }
            }
            return hit;
        }
        case RTB_HITTABLE_MEDIUM_OF: {  // ConstantMedium.hit, :462-507, over any boundary
            AnyHit r1, r2;
            if (!hit_any(ct, in.child, r, -inf, inf, &r1, key, segment)) return false;
            if (!hit_any(ct, in.child, r, r1.t + 0.0001f, inf, &r2, key, segment)) return false;
            float t1 = r1.t, t2 = r2.t;
            if (t1 < t_min) t1 = t_min;
            if (t2 > t_max) t2 = t_max;
            if (t1 >= t2) return false;
            if (t1 < 0.0f) t1 = 0.0f;
            const float ray_length = sqrtf(length_squared(r.d));
            const float distance_inside_boundary = (t2 - t1) * ray_length;
            const float hit_distance = in.p.x * logf(rng_block(key, segment, 0x40000000u + idx).x);
            if (hit_distance > distance_inside_boundary) return false;
            h->t = t1 + hit_distance / ray_length;
            h->p = r.o + splat3(h->t) * r.d;
            h->normal = f3(1.0f, 0.0f, 0.0f);
            h->u = h->v = 0.0f;
            h->front_face = true;
            h->prim = idx;
            return true;
        }
        default: return false;
    }
}

// ------------------------------------------------------------------ textures
// perlin_interp + Perlin.noise (src/perlin.zig:30-53, :117-152).  (i_f*uu + (1-i_f)*(1-uu)) is
// exactly uu for i=1 and (1-uu) for i=0, so the weights are selected instead of computed.
__device__ __forceinline__ float perlin_noise(const DevPerlin& pl, float3 p) {
    const float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    const float u = p.x - fx, v = p.y - fy, w = p.z - fz;
    const int i = (int)fx, j = (int)fy, k = (int)fz;
    const float uu = u * u * (3.0f - 2.0f * u);
    const float vv = v * v * (3.0f - 2.0f * v);
    const float ww = w * w * (3.0f - 2.0f * w);
    float accum = 0.0f;
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
        for (int dj = 0; dj < 2; ++dj)
#pragma unroll
            for (int dk = 0; dk < 2; ++dk) {
                const uint32_t idx =
                    pl.perm_x[(i + di) & 255] ^ pl.perm_y[(j + dj) & 255] ^ pl.perm_z[(k + dk) & 255];
                const float4 g = pl.ranvec[idx];
                const float3 weight_v = f3(u - (float)di, v - (float)dj, w - (float)dk);
                const float wu = di ? uu : (1.0f - uu);
                const float wv = dj ? vv : (1.0f - vv);
                const float ww_ = dk ? ww : (1.0f - ww);
                accum += wu * wv * ww_ * dot3(f3(g), weight_v);
            }
    return accum;
}
// Perlin.turb (src/perlin.zig:103-115)
__device__ __forceinline__ float perlin_turb(const DevPerlin& pl, float3 p, int depth) {
    float accum = 0.0f;
    float3 temp_p = p;
    float weight = 1.0f;
    for (int i = 0; i < depth; ++i) {
        accum += weight * perlin_noise(pl, temp_p);
        weight *= 0.5f;
        temp_p = temp_p * splat3(2.0f);
    }
    return fabsf(accum);
}

__device__ __forceinline__ bool texture_needs_uv(uint32_t tex_type) { return tex_type == RTB_TEX_IMAGE; }

// Texture.value (src/textures.zig:22-26) for the non-solid variants.
struct TexTables {
    const float4* textures;
    const DevPerlin* perlins;
    const DevImage* images;
};
static __device__ __noinline__ float3 texture_value_slow(TexTables sc, uint32_t tex, float u, float v, float3 p) {
    const float4 t0 = sc.textures[3u * tex];
    const uint32_t type = __float_as_uint(t0.x);
    const uint32_t index = __float_as_uint(t0.y);
    if (type == RTB_TEX_CHECKER) {  // src/textures.zig:60-72
        const int xi = (int)floorf(t0.z * p.x);
        const int yi = (int)floorf(t0.z * p.y);
        const int zi = (int)floorf(t0.z * p.z);
        const bool is_even = ((xi + yi + zi) % 2) == 0;
        return f3(sc.textures[3u * tex + (is_even ? 1u : 2u)]);
    }
    if (type == RTB_TEX_IMAGE) {  // src/textures.zig:85-104, src/rtw_image.zig:37-62
        const DevImage im = sc.images[index];
        if (im.height == 0u) return f3(0.0f, 1.0f, 1.0f);
        const float new_u = u < 0.0f ? 0.0f : (u > 1.0f ? 1.0f : u);
        const float cv = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
        const float new_v = 1.0f - cv;
        uint32_t i = (uint32_t)floorf(new_u * (float)im.width);
        uint32_t j = (uint32_t)floorf(new_v * (float)im.height);
        if (!(i < im.width)) i = im.width - 1u;
        if (!(j < im.height)) j = im.height - 1u;
        const uchar4 px = im.texels[(size_t)j * im.width + i];
        const float color_scale = 1.0f / 255.0f;
        return f3(color_scale * (float)px.x, color_scale * (float)px.y, color_scale * (float)px.z);
    }
    if (type == RTB_TEX_NOISE) {  // src/textures.zig:118-123
        const float3 s = splat3(t0.z) * p;
        return splat3(0.5f * (1.0f + sinf(s.z + 10.0f * perlin_turb(sc.perlins[index], s, 7))));
    }
    return f3(sc.textures[3u * tex + 1u]);  // solid
}

// ------------------------------------------------------------------ materials
// reflectance (src/material.zig:101-106); (1-cos)^5 by repeated multiplication.
__device__ __forceinline__ float schlick(float cosine, float ref_idx) {
    float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
    r0 = r0 * r0;
    const float x = 1.0f - cosine;
    const float x2 = x * x;
    return r0 + (1.0f - r0) * (x2 * x2 * x);
}

struct ShadeResult {
    float3 emitted;
    float3 attenuation;
    DRay scattered;
    bool scatters;
};

__device__ __forceinline__ ShadeResult shade_hit(const DevScene& sc, const DHit& h, float4 m0, float4 m1, const DRay& r,
                                                 const RngKey& key, uint32_t segment, const float3* unit = nullptr);

// emitted + scatter for the nearest hit (src/camera.zig:194-196 -> src/material.zig:18-30).
// `segment` (>= 1) keys this hit's RNG stream; block 0 word 3 is the dielectric reflectance draw.
// `f0,f1` = the hit object's leaf record, `m0,m1` = its material record.
template <bool QUADS>
__device__ __forceinline__ ShadeResult shade_rec(const DevScene& sc, float4 f0, float4 f1, float4 m0, float4 m1,
                                                 const DRay& r, float t, const RngKey& key, uint32_t segment,
                                                 const float3* unit = nullptr) {
    const uint32_t tex_type = (__float_as_uint(m0.x) >> 8) & 0xffu;
    DHit h;
    if (QUADS && (__float_as_uint(f0.w) >> 30) == KIND_QUAD && __float_as_uint(f1.x) == COMPLEX_GENERIC) {
        // a wrapper has no material of its own: shade with the material of the primitive that was hit inside it
        uint32_t prim = 0u;
        h = finish_hit_rec<QUADS, true>(f0, f1, complex_tables(sc), r, t, key, segment, &prim);
        m0 = sc.prims[4u * (size_t)prim + 2u];
        m1 = sc.prims[4u * (size_t)prim + 3u];
        return shade_hit(sc, h, m0, m1, r, key, segment);
    }
    if (texture_needs_uv(tex_type))
        h = finish_hit_rec<QUADS, true>(f0, f1, complex_tables(sc), r, t);
    else
        h = finish_hit_rec<QUADS, false>(f0, f1, complex_tables(sc), r, t);
    return shade_hit(sc, h, m0, m1, r, key, segment, unit);
}

// emitted + scatter once the hit record and the material record are known.  `unit` (optional): this hit's
// randomUnitVector, already drawn by random_unit_vector_coop (lambertian and metal only).
__device__ __forceinline__ ShadeResult shade_hit(const DevScene& sc, const DHit& h, float4 m0, float4 m1, const DRay& r,
                                                 const RngKey& key, uint32_t segment, const float3* unit) {
    ShadeResult out;
    const TexTables tt{sc.textures, sc.perlins, sc.images};
    out.emitted = f3(0.0f, 0.0f, 0.0f);
    const uint32_t tag = __float_as_uint(m0.x);
    const uint32_t type = tag & 0xffu;
    const uint32_t tex_type = (tag >> 8) & 0xffu;
    out.scattered.o = h.p;
    out.scattered.time = r.time;
    if (type == RTB_MAT_LAMBERTIAN) {  // src/material.zig:43-54
        float3 dir = h.normal + (unit ? *unit : random_unit_vector(key, segment, rng_block(key, segment, 0u)));
        const float s = 1e-8f;
        if (fabsf(dir.x) < s && fabsf(dir.y) < s && fabsf(dir.z) < s) dir = h.normal;  // nearZero
        out.scattered.d = dir;
        out.attenuation =
            (tex_type == RTB_TEX_SOLID) ? f3(m1) : texture_value_slow(tt, __float_as_uint(m0.y), h.u, h.v, h.p);
        out.scatters = true;
    } else if (type == RTB_MAT_METAL) {  // src/material.zig:65-70
        const float3 reflected = reflect3(unit_vector(r.d), h.normal);
        out.scattered.d =
            reflected + splat3(m0.z) * (unit ? *unit : random_unit_vector(key, segment, rng_block(key, segment, 0u)));
        out.attenuation = f3(m1);
        out.scatters = dot3(out.scattered.d, h.normal) > 0.0f;
    } else if (type == RTB_MAT_DIELECTRIC) {  // src/material.zig:80-98
        out.attenuation = f3(1.0f, 1.0f, 1.0f);
        const float refraction_ratio = h.front_face ? (1.0f / m0.w) : m0.w;
        const float3 unit_direction = unit_vector(r.d);
        const float cos_theta = fminf(dot3(-unit_direction, h.normal), 1.0f);
        const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
        const bool cannot_refract = refraction_ratio * sin_theta > 1.0f;
        bool do_reflect = cannot_refract;
        if (!do_reflect) do_reflect = schlick(cos_theta, refraction_ratio) > rng_block(key, segment, 0u).w;
        out.scattered.d =
            do_reflect ? reflect3(unit_direction, h.normal) : refract3(unit_direction, h.normal, refraction_ratio);
        out.scatters = true;
    } else if (type == RTB_MAT_DIFFUSE_LIGHT) {  // src/material.zig:119-125
        out.emitted =
            (tex_type == RTB_TEX_SOLID) ? f3(m1) : texture_value_slow(tt, __float_as_uint(m0.y), h.u, h.v, h.p);
        out.attenuation = f3(0.0f, 0.0f, 0.0f);
        out.scattered.d = f3(0.0f, 0.0f, 0.0f);
        out.scatters = false;
    } else {  // isotropic, src/material.zig:139-143
        const float4 b0 = rng_block(key, segment, 0u);
        out.scattered.d = random_unit_vector(key, segment, b0);
        out.attenuation =
            (tex_type == RTB_TEX_SOLID) ? f3(m1) : texture_value_slow(tt, __float_as_uint(m0.y), h.u, h.v, h.p);
        out.scatters = true;
    }
    return out;
}

// shade_rec with the records fetched through the node/prim array, object -> material table.
template <bool QUADS>
__device__ __forceinline__ ShadeResult shade(const DevScene& sc, const float4* __restrict__ nodes, const DRay& r,
                                             Nearest best, const RngKey& key, uint32_t segment) {
    const float4 f0 = nodes[2u * best.node];
    const float4 f1 = nodes[2u * best.node + 1u];
    const uint32_t mat = sc.object_material[__float_as_uint(f0.w) & RTB_META_INDEX_MASK];
    return shade_rec<QUADS>(sc, f0, f1, sc.materials[2u * mat], sc.materials[2u * mat + 1u], r, best.t, key, segment);
}

// Miss colour: solid background (src/camera.zig:207) or the legacy sky gradient (:204-206).
__device__ __forceinline__ float3 miss_color(const DevCamera& cam, const DRay& r) {
    if (cam.background_mode == RTB_BACKGROUND_SKY) {
        const float3 unit_direction = unit_vector(r.d);
        const float a = 0.5f * (unit_direction.y + 1.0f);
        return f3(1.0f, 1.0f, 1.0f) * splat3(1.0f - a) + f3(0.5f, 0.7f, 1.0f) * splat3(a);
    }
    return cam.background;
}

// color.toGamma2 + @intFromFloat (src/color.zig:43-62, src/camera.zig:57-65); NaN -> 0
// (the reference's @intFromFloat(NaN) is undefined behaviour).
__device__ __forceinline__ uchar4 quantise(float4 acc, float n) {
    const float scale = 1.0f / n;
    float c[3] = {acc.x, acc.y, acc.z};
    uint8_t o[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float x = c[k] * scale;
        x = sqrtf(x);
        if (x < 0.0f) x = 0.0f;
        if (x > 0.999f) x = 0.999f;
        float g = 256.0f * x;
        if (!(g == g)) g = 0.0f;
        o[k] = (uint8_t)g;
    }
    return make_uchar4(o[0], o[1], o[2], 255);
}

}  // namespace rtb
