// oracle.cpp — CPU restatement of the reference's path-tracing hot path.  See oracle.h: TEST
// INFRASTRUCTURE ONLY, pinned on the reference's committed renders for camera/spheres/materials/colour; visiting order, textures, quads, media unpinned.  Build: g++ -std=c++17 -O2 -ffp-contract=off (Makefile).
//
// Evaluation order of every float expression follows the Zig source (element-wise @Vector ops,
// left-associative + and *), because decision arithmetic (slab test, discriminant, roots,
// front_face) must be reproduced bit for bit by the CUDA kernels.
#include "oracle.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------- vec3.zig:4-38
struct V3 {
    float x, y, z;
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 v3(const float* p) { return V3{p[0], p[1], p[2]}; }
inline V3 splat3(float s) { return V3{s, s, s}; }  // vec3.zig:24-26
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
inline float lengthSquared(V3 u) { return u.x * u.x + u.y * u.y + u.z * u.z; }  // vec3.zig:15-17
inline float length(V3 u) { return std::sqrt(lengthSquared(u)); }               // vec3.zig:11-13
inline float dot(V3 u, V3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }      // vec3.zig:28-30
inline V3 cross(V3 u, V3 v) {                                                   // vec3.zig:32-34
    return V3{u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x};
}
inline V3 unitVector(V3 v) { return v / splat3(length(v)); }  // vec3.zig:36-38
inline bool nearZero(V3 u) {                                  // vec3.zig:19-22
    const float s = 1e-8f;
    return std::fabs(u.x) < s && std::fabs(u.y) < s && std::fabs(u.z) < s;
}
inline V3 reflect(V3 v, V3 n) { return v - n * splat3(dot(v, n) * 2); }  // vec3.zig:77-79
inline V3 refract(V3 uv, V3 n, float etai_over_etat) {                   // vec3.zig:81-86
    const float cos_theta = std::fmin(dot(-uv, n), 1.0f);
    const V3 r_out_perp = splat3(etai_over_etat) * (uv + n * splat3(cos_theta));
    const V3 r_out_parallel = n * splat3(-std::sqrt(std::fabs(1.0f - lengthSquared(r_out_perp))));
    return r_out_perp + r_out_parallel;
}
inline void store3(float* o, V3 v) {
    o[0] = v.x;
    o[1] = v.y;
    o[2] = v.z;
}

const float kInfinity = std::numeric_limits<float>::infinity();  // rtweekend.zig:3
const float kPi = 3.1415926535897932385f;                        // rtweekend.zig:4

// ---------------------------------------------------------------- ray.zig:4-12, interval.zig:4-20
struct Ray {
    V3 origin, direction;
    float time;
};
inline V3 at(const Ray& r, float t) { return r.origin + splat3(t) * r.direction; }  // ray.zig:9-11
struct Interval {
    float min, max;
};
inline bool contains(Interval i, float x) { return i.min <= x && x <= i.max; }  // interval.zig:8-10
inline bool surrounds(Interval i, float x) { return i.min < x && x < i.max; }   // interval.zig:12-14
inline float clampI(Interval i, float x) {                                      // interval.zig:16-20
    if (x < i.min) return i.min;
    if (x > i.max) return i.max;
    return x;
}

// ---------------------------------------------------------------- RNG (replaces rtweekend.zig:14-16)
inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k[0];
    const uint32_t n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
}
inline void philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0];
    out[1] = c[1];
    out[2] = c[2];
    out[3] = c[3];
}
inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// One Philox block of the stream (seed; pixel, sample, segment, block).
struct Block {
    float r[4];
};
inline Block rng_block(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t segment, uint32_t block) {
    const uint32_t ctr[4] = {pixel, sample, segment, block};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    philox(ctr, key, o);
    Block b;
    for (int i = 0; i < 4; ++i) b.r[i] = u01(o[i]);
    return b;
}
// rtweekend.randomDoubleRange(min,max) = min + (max-min)*r   (rtweekend.zig:18-20)
inline float rangeOf(float r, float min, float max) { return min + (max - min) * r; }

// vec3.randomUnitVector (vec3.zig:59-68): rejection try j of segment `segment` uses words 0..2
// of block j.
inline V3 randomUnitVector(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t segment) {
    for (uint32_t j = 0;; ++j) {
        const Block b = rng_block(seed, pixel, sample, segment, j);
        const V3 p = v3(rangeOf(b.r[0], -1, 1), rangeOf(b.r[1], -1, 1), rangeOf(b.r[2], -1, 1));
        if (lengthSquared(p) < 1) return unitVector(p);
    }
}

// ---------------------------------------------------------------- counters
struct Counters {
    uint64_t rays = 0, box = 0, obj = 0, hits = 0;
};

// ---------------------------------------------------------------- HitRecord (objects.zig:21-37)
struct HitRecord {
    V3 p{0, 0, 0}, normal{0, 0, 0};
    uint32_t mat = 0;  // index; the reference copies the Material by value
    float t = 0, u = 0, v = 0;
    bool front_face = false;
    int32_t object = -1;
};
inline void setFaceNormal(HitRecord& rec, const Ray& r, V3 outward_normal) {  // objects.zig:30-36
    rec.front_face = dot(r.direction, outward_normal) < 0;
    rec.normal = rec.front_face ? outward_normal : -outward_normal;
}

// objects.zig:101-114
inline void getSphereUV(V3 p, float* u, float* v) {
    const float theta = std::acos(-p.y);
    const float phi = std::atan2(-p.z, p.x) + kPi;
    *u = phi / (2 * kPi);
    *v = theta / kPi;
}

// Sphere.hit, objects.zig:116-148
bool sphereHit(const RtbHittable& s, const Ray& r, Interval ray_t, HitRecord& rec) {
    const V3 center1 = v3(s.a);
    const V3 center = s.is_moving ? center1 + splat3(r.time) * v3(s.b) : center1;  // :94-98, :121
    const V3 oc = r.origin - center;
    const float a = lengthSquared(r.direction);
    const float half_b = dot(oc, r.direction);
    const float c = lengthSquared(oc) - s.radius * s.radius;
    const float discriminant = half_b * half_b - a * c;
    if (discriminant < 0) return false;
    const float sqrtd = std::sqrt(discriminant);
    float root = (-half_b - sqrtd) / a;
    if (!surrounds(ray_t, root)) {
        root = (-half_b + sqrtd) / a;
        if (!surrounds(ray_t, root)) return false;
    }
    rec.t = root;
    rec.p = at(r, rec.t);
    const V3 outward_normal = (rec.p - center) / splat3(s.radius);
    setFaceNormal(rec, r, outward_normal);
    getSphereUV(outward_normal, &rec.u, &rec.v);
    rec.mat = s.material;
    return true;
}

// Quad.init derived fields (objects.zig:206-211) + Quad.hit (:226-261)
bool quadHitQUV(V3 q, V3 u, V3 v, uint32_t material, const Ray& r, Interval ray_t, HitRecord& rec);

bool quadHit(const RtbHittable& qd, const Ray& r, Interval ray_t, HitRecord& rec) {
    return quadHitQUV(v3(qd.a), v3(qd.b), v3(qd.c), qd.material, r, ray_t, rec);
}

// createBox (objects.zig:510-532): the six quads of the HittableList, in the order they are added.  (As in
// the reference, the z = min face appears twice — entries 0 and 2 — and there is no z = max face.)
void boxFaces(V3 a, V3 b, V3 q[6], V3 u[6], V3 v[6]) {
    const V3 mn = v3(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z));
    const V3 mx = v3(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z));
    const V3 dx = v3(mx.x - mn.x, 0, 0), dy = v3(0, mx.y - mn.y, 0), dz = v3(0, 0, mx.z - mn.z);
    q[0] = v3(mn.x, mn.y, mn.z); u[0] = dx;  v[0] = dy;
    q[1] = v3(mx.x, mn.y, mx.z); u[1] = -dz; v[1] = dy;
    q[2] = v3(mx.x, mn.y, mn.z); u[2] = -dx; v[2] = dy;
    q[3] = v3(mn.x, mn.y, mn.z); u[3] = dz;  v[3] = dy;
    q[4] = v3(mn.x, mx.y, mx.z); u[4] = dx;  v[4] = -dz;
    q[5] = v3(mn.x, mn.y, mn.z); u[5] = dx;  v[5] = dz;
}

// Translate.hit (objects.zig:326-345) around RotateY.hit (:404-442) around HittableList.hit (:286-304)
// over createBox's quads.
bool boxHit(const RtbHittable& bx, const Ray& r, Interval ray_t, HitRecord& rec) {
    const V3 offset = v3(bx.c);
    const float sin_theta = bx.sin_theta, cos_theta = bx.cos_theta;
    const Ray moved{r.origin - offset, r.direction, r.time};  // Translate.hit :331
    V3 origin = moved.origin, direction = moved.direction;    // RotateY.hit :410-419
    origin.x = cos_theta * moved.origin.x - sin_theta * moved.origin.z;
    origin.z = sin_theta * moved.origin.x + cos_theta * moved.origin.z;
    direction.x = cos_theta * moved.direction.x - sin_theta * moved.direction.z;
    direction.z = sin_theta * moved.direction.x + cos_theta * moved.direction.z;
    const Ray rotated{origin, direction, r.time};
    V3 q[6], u[6], v[6];
    boxFaces(v3(bx.a), v3(bx.b), q, u, v);
    bool hit = false;
    float closest_so_far = ray_t.max;  // HittableList.hit :291-301
    HitRecord h;
    for (int f = 0; f < 6; ++f) {
        HitRecord cand;
        if (quadHitQUV(q[f], u[f], v[f], bx.material, rotated, Interval{ray_t.min, closest_so_far}, cand)) {
            closest_so_far = cand.t;
            h = cand;
            hit = true;
        }
    }
    if (!hit) return false;
    rec = h;  // RotateY.hit :425-439
    rec.p.x = cos_theta * h.p.x + sin_theta * h.p.z;
    rec.p.z = -sin_theta * h.p.x + cos_theta * h.p.z;
    rec.normal.x = cos_theta * h.normal.x + sin_theta * h.normal.z;
    rec.normal.z = -sin_theta * h.normal.x + cos_theta * h.normal.z;
    rec.p = rec.p + offset;  // Translate.hit :340
    return true;
}

bool quadHitQUV(V3 q, V3 u, V3 v, uint32_t material, const Ray& r, Interval ray_t, HitRecord& rec) {
    const V3 n = cross(u, v);
    const V3 normal = unitVector(n);
    const float d = dot(normal, q);
    const V3 w = n / splat3(dot(n, n));
    const float denom = dot(normal, r.direction);
    if (std::fabs(denom) < 1e-8f) return false;
    const float t = (d - dot(normal, r.origin)) / denom;
    if (!contains(ray_t, t)) return false;
    const V3 intersection = at(r, t);
    const V3 planar = intersection - q;
    const float alpha = dot(w, cross(planar, v));
    const float beta = dot(w, cross(u, planar));
    if ((alpha < 0) || (1 < alpha) || (beta < 0) || (1 < beta)) return false;  // isInterior :217-224
    rec.u = alpha;
    rec.v = beta;
    rec.t = t;
    rec.p = intersection;
    rec.mat = material;
    setFaceNormal(rec, r, normal);
    return true;
}

// Hittable.hit dispatch, objects.zig:49-53
// The RNG stream of the ray segment being traced: ConstantMedium.hit draws a random number INSIDE hit
// (objects.zig:484).  Its draw is word 0 of block (0x40000000 + object index) of the segment's stream, so it
// does not depend on the order in which the tree is walked.  Set by the callers of worldHit.
struct HitCtx {
    uint64_t seed = 0;
    uint32_t pixel = 0, sample = 0, segment = 1;
};
thread_local HitCtx g_ctx;

// ConstantMedium.hit (objects.zig:462-507) with a box instance as the boundary (cornellBoxSmoke, main.zig:223-236).
bool mediumHit(const RtbHittable& m, uint32_t index, const Ray& r, Interval ray_t, HitRecord& rec) {
    HitRecord rec_1, rec_2;
    if (!boxHit(m, r, Interval{-kInfinity, kInfinity}, rec_1)) return false;               // intervals.universe
    if (!boxHit(m, r, Interval{rec_1.t + 0.0001f, kInfinity}, rec_2)) return false;
    if (rec_1.t < ray_t.min) rec_1.t = ray_t.min;
    if (rec_2.t > ray_t.max) rec_2.t = ray_t.max;
    if (rec_1.t >= rec_2.t) return false;
    if (rec_1.t < 0) rec_1.t = 0;
    const float ray_length = length(r.direction);
    const float distance_inside_boundary = (rec_2.t - rec_1.t) * ray_length;
    const float rnd = rng_block(g_ctx.seed, g_ctx.pixel, g_ctx.sample, g_ctx.segment, 0x40000000u + index).r[0];
    const float hit_distance = m.radius * std::log(rnd);  // radius carries neg_inv_density = -1/d (objects.zig:451)
    if (hit_distance > distance_inside_boundary) return false;
    rec = HitRecord{};
    rec.t = rec_1.t + hit_distance / ray_length;
    rec.p = at(r, rec.t);
    rec.normal = v3(1, 0, 0);  // arbitrary
    rec.front_face = true;     // also arbitrary
    rec.mat = m.material;      // phase_function (Isotropic)
    return true;
}

bool anyHit(const RtbSceneDesc* sc, uint32_t index, const Ray& r, Interval ray_t, HitRecord& rec);

// Translate.hit, objects.zig:327-345
bool translateHit(const RtbSceneDesc* sc, const RtbHittable& t, const Ray& r, Interval ray_t, HitRecord& rec) {
    const V3 offset = v3(t.a);
    const Ray moved{r.origin - offset, r.direction, r.time};
    if (!anyHit(sc, t.child, moved, ray_t, rec)) return false;
    rec.p = rec.p + offset;
    return true;
}

// RotateY.hit, objects.zig:404-442
bool rotateYHit(const RtbSceneDesc* sc, const RtbHittable& ry, const Ray& r, Interval ray_t, HitRecord& rec) {
    const float sin_theta = ry.sin_theta, cos_theta = ry.cos_theta;
    V3 origin = r.origin, direction = r.direction;
    origin.x = cos_theta * r.origin.x - sin_theta * r.origin.z;
    origin.z = sin_theta * r.origin.x + cos_theta * r.origin.z;
    direction.x = cos_theta * r.direction.x - sin_theta * r.direction.z;
    direction.z = sin_theta * r.direction.x + cos_theta * r.direction.z;
    const Ray rotated{origin, direction, r.time};
    HitRecord h;
    if (!anyHit(sc, ry.child, rotated, ray_t, h)) return false;
    rec = h;
    rec.p.x = cos_theta * h.p.x + sin_theta * h.p.z;
    rec.p.z = -sin_theta * h.p.x + cos_theta * h.p.z;
    rec.normal.x = cos_theta * h.normal.x + sin_theta * h.normal.z;
    rec.normal.z = -sin_theta * h.normal.x + cos_theta * h.normal.z;
    return true;
}

// HittableList.hit, objects.zig:286-304: members in order, ray_t.max = closest so far
bool listHit(const RtbSceneDesc* sc, const RtbHittable& l, const Ray& r, Interval ray_t, HitRecord& rec) {
    bool hit = false;
    float closest_so_far = ray_t.max;
    for (uint32_t k = 0; k < l.material; ++k) {  // `material` carries the member count of a list
        HitRecord cand;
        if (anyHit(sc, l.child + k, r, Interval{ray_t.min, closest_so_far}, cand)) {
            closest_so_far = cand.t;
            rec = cand;
            hit = true;
        }
    }
    return hit;
}

// ConstantMedium.hit (objects.zig:462-507) over ANY boundary hittable.
bool mediumOfHit(const RtbSceneDesc* sc, const RtbHittable& m, uint32_t index, const Ray& r, Interval ray_t, HitRecord& rec) {
    HitRecord rec_1, rec_2;
    if (!anyHit(sc, m.child, r, Interval{-kInfinity, kInfinity}, rec_1)) return false;
    if (!anyHit(sc, m.child, r, Interval{rec_1.t + 0.0001f, kInfinity}, rec_2)) return false;
    if (rec_1.t < ray_t.min) rec_1.t = ray_t.min;
    if (rec_2.t > ray_t.max) rec_2.t = ray_t.max;
    if (rec_1.t >= rec_2.t) return false;
    if (rec_1.t < 0) rec_1.t = 0;
    const float ray_length = length(r.direction);
    const float distance_inside_boundary = (rec_2.t - rec_1.t) * ray_length;
    const float rnd = rng_block(g_ctx.seed, g_ctx.pixel, g_ctx.sample, g_ctx.segment, 0x40000000u + index).r[0];
    const float hit_distance = m.radius * std::log(rnd);
    if (hit_distance > distance_inside_boundary) return false;
    rec = HitRecord{};
    rec.t = rec_1.t + hit_distance / ray_length;
    rec.p = at(r, rec.t);
    rec.normal = v3(1, 0, 0);
    rec.front_face = true;
    rec.mat = m.material;
    return true;
}

// Hittable.hit on any entry of the hittable array (top-level object or a wrapper's child).
bool anyHit(const RtbSceneDesc* sc, uint32_t index, const Ray& r, Interval ray_t, HitRecord& rec) {
    const RtbHittable& h = sc->hittables[index];
    switch (h.type) {
        case RTB_HITTABLE_SPHERE: return sphereHit(h, r, ray_t, rec);
        case RTB_HITTABLE_QUAD: return quadHit(h, r, ray_t, rec);
        case RTB_HITTABLE_BOX: return boxHit(h, r, ray_t, rec);
        case RTB_HITTABLE_CONSTANT_MEDIUM: return mediumHit(h, index, r, ray_t, rec);
        case RTB_HITTABLE_TRANSLATE: return translateHit(sc, h, r, ray_t, rec);
        case RTB_HITTABLE_ROTATE_Y: return rotateYHit(sc, h, r, ray_t, rec);
        case RTB_HITTABLE_LIST: return listHit(sc, h, r, ray_t, rec);
        case RTB_HITTABLE_MEDIUM_OF: return mediumOfHit(sc, h, index, r, ray_t, rec);
        default: return false;
    }
}

bool hittableHit(const RtbSceneDesc* sc, uint32_t index, const Ray& r, Interval ray_t, HitRecord& rec) {
    const bool ok = anyHit(sc, index, r, ray_t, rec);
    if (ok) rec.object = (int32_t)index;  // the top-level object (the BVH leaf), whatever is nested inside it
    return ok;
}

// Aabb.hit, aabb.zig:82-114
bool aabbHit(const float* bmin, const float* bmax, const Ray& r, Interval ray_t) {
    float ray_t_min = ray_t.min;
    float ray_t_max = ray_t.max;
    const float dir[3] = {r.direction.x, r.direction.y, r.direction.z};
    const float org[3] = {r.origin.x, r.origin.y, r.origin.z};
    for (int a = 0; a < 3; ++a) {
        const float invD = 1 / dir[a];
        const float orig = org[a];
        float t0 = (bmin[a] - orig) * invD;
        float t1 = (bmax[a] - orig) * invD;
        if (invD < 0) {
            const float temp = t1;
            t1 = t0;
            t0 = temp;
        }
        if (t0 > ray_t_min) ray_t_min = t0;
        if (t1 < ray_t_max) ray_t_max = t1;
        if (ray_t_max <= ray_t_min) return false;
    }
    return true;
}

// BVHNode.hit, bvh.zig:122-136 (recursive DFS)
bool nodeHit(const RtbSceneDesc* sc, int32_t node, const Ray& r, Interval ray_t, HitRecord& rec, Counters& cn) {
    const RtbBvhNode& n = sc->nodes[node];
    if (n.leaf >= 0) {
        ++cn.obj;
        return hittableHit(sc, (uint32_t)n.leaf, r, ray_t, rec);
    }
    ++cn.box;
    if (!aabbHit(n.bmin, n.bmax, r, ray_t)) return false;
    HitRecord left;
    const bool hit_left = nodeHit(sc, n.left, r, ray_t, left, cn);
    const Interval rInterval{ray_t.min, hit_left ? left.t : ray_t.max};
    HitRecord right;
    const bool hit_right = nodeHit(sc, n.right, r, rInterval, right, cn);
    if (hit_right) {  // hit_record_right orelse hit_record_left
        rec = right;
        return true;
    }
    if (hit_left) {
        rec = left;
        return true;
    }
    return false;
}

// BVHTree.hit, bvh.zig:39-41
inline bool worldHit(const RtbSceneDesc* sc, const Ray& r, Interval ray_t, HitRecord& rec, Counters& cn) {
    if (sc->n_nodes == 0 || sc->root < 0) return false;
    return nodeHit(sc, sc->root, r, ray_t, rec, cn);
}

// ---------------------------------------------------------------- perlin.zig
float perlinInterp(const V3 c[2][2][2], float u, float v, float w) {  // perlin.zig:30-53
    const float uu = u * u * (3 - 2 * u);
    const float vv = v * v * (3 - 2 * v);
    const float ww = w * w * (3 - 2 * w);
    float accum = 0;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {
                const float i_f = (float)i, j_f = (float)j, k_f = (float)k;
                const V3 weight_v = v3(u - i_f, v - j_f, w - k_f);
                accum += (i_f * uu + (1 - i_f) * (1 - uu)) * (j_f * vv + (1 - j_f) * (1 - vv)) *
                         (k_f * ww + (1 - k_f) * (1 - ww)) * dot(c[i][j][k], weight_v);
            }
    return accum;
}
float perlinNoise(const RtbPerlin* pl, V3 p) {  // perlin.zig:117-152
    const float u = p.x - std::floor(p.x);
    const float v = p.y - std::floor(p.y);
    const float w = p.z - std::floor(p.z);
    const int32_t i = (int32_t)std::floor(p.x);
    const int32_t j = (int32_t)std::floor(p.y);
    const int32_t k = (int32_t)std::floor(p.z);
    V3 c[2][2][2];
    for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj)
            for (int dk = 0; dk < 2; ++dk) {
                const uint32_t idi = (uint32_t)((i + di) & 255);
                const uint32_t idj = (uint32_t)((j + dj) & 255);
                const uint32_t idk = (uint32_t)((k + dk) & 255);
                const uint16_t px = pl->perm_x[idi], py = pl->perm_y[idj], pz = pl->perm_z[idk];
                c[di][dj][dk] = v3(pl->ranvec[(px ^ py ^ pz) & 255]);
            }
    return perlinInterp(c, u, v, w);
}
float perlinTurb(const RtbPerlin* pl, V3 p, int depth) {  // perlin.zig:103-115
    float accum = 0;
    V3 temp_p = p;
    float weight = 1.0f;
    for (int i = 0; i < depth; ++i) {
        accum += weight * perlinNoise(pl, temp_p);
        weight *= 0.5f;
        temp_p = temp_p * splat3(2);
    }
    return std::fabs(accum);
}

// ---------------------------------------------------------------- textures.zig, rtw_image.zig
inline uint32_t imgClamp(uint32_t x, uint32_t low, uint32_t high) {  // rtw_image.zig:37-45
    if (x < low) return low;
    if (x < high) return x;
    return high - 1;
}
V3 textureValue(const RtbSceneDesc* sc, uint32_t tex, float u, float v, V3 p) {  // textures.zig:22-26
    const RtbTexture& t = sc->textures[tex];
    switch (t.type) {
        case RTB_TEX_SOLID:  // textures.zig:43-45
            return v3(t.color);
        case RTB_TEX_CHECKER: {  // textures.zig:60-72
            const int32_t xi = (int32_t)std::floor(t.scale * p.x);
            const int32_t yi = (int32_t)std::floor(t.scale * p.y);
            const int32_t zi = (int32_t)std::floor(t.scale * p.z);
            const bool isEven = ((xi + yi + zi) % 2) == 0;  // @rem
            return isEven ? v3(t.color) : v3(t.color2);
        }
        case RTB_TEX_IMAGE: {  // textures.zig:85-104
            const RtbImage& im = sc->images[t.index];
            if (im.height <= 0) return v3(0, 1, 1);
            const Interval unit{0, 1};
            const float new_u = clampI(unit, u);
            const float new_v = 1.0f - clampI(unit, v);
            const float u_p = new_u * (float)im.width;
            const float v_p = new_v * (float)im.height;
            const uint32_t i = (uint32_t)std::floor(u_p);
            const uint32_t j = (uint32_t)std::floor(v_p);
            const uint32_t nx = imgClamp(i, 0, im.width);   // rtw_image.zig:54-55
            const uint32_t ny = imgClamp(j, 0, im.height);
            const size_t start = (size_t)ny * im.bytes_per_row + (size_t)nx * 4;  // :56-57
            const float color_scale = 1.0f / 255.0f;
            return v3(color_scale * (float)im.data[start], color_scale * (float)im.data[start + 1],
                      color_scale * (float)im.data[start + 2]);
        }
        case RTB_TEX_NOISE: {  // textures.zig:118-123
            const V3 s = splat3(t.scale) * p;
            return splat3(0.5f * (1 + std::sin(s.z + 10 * perlinTurb(&sc->perlins[t.index], s, 7))));
        }
    }
    return v3(0, 0, 0);
}

// ---------------------------------------------------------------- material.zig
float reflectance(float cosine, float ref_idx) {  // material.zig:101-106
    float r0 = (1 - ref_idx) / (1 + ref_idx);
    r0 = r0 * r0;
    return r0 + (1 - r0) * std::pow((1 - cosine), 5.0f);
}

V3 emitted(const RtbSceneDesc* sc, const HitRecord& rec) {  // material.zig:24-29, :123-125
    const RtbMaterial& m = sc->materials[rec.mat];
    if (m.type == RTB_MAT_DIFFUSE_LIGHT) return textureValue(sc, m.texture, rec.u, rec.v, rec.p);
    return v3(0, 0, 0);
}

bool scatter(const RtbSceneDesc* sc, const Ray& r_in, const HitRecord& rec, uint64_t seed, uint32_t pixel,
             uint32_t sample, uint32_t segment, V3& attenuation, Ray& scattered) {  // material.zig:18-22
    const RtbMaterial& m = sc->materials[rec.mat];
    switch (m.type) {
        case RTB_MAT_LAMBERTIAN: {  // material.zig:43-54
            V3 scatter_direction = rec.normal + randomUnitVector(seed, pixel, sample, segment);
            if (nearZero(scatter_direction)) scatter_direction = rec.normal;
            scattered = Ray{rec.p, scatter_direction, r_in.time};
            attenuation = textureValue(sc, m.texture, rec.u, rec.v, rec.p);
            return true;
        }
        case RTB_MAT_METAL: {  // material.zig:65-70
            const V3 reflected = reflect(unitVector(r_in.direction), rec.normal);
            scattered = Ray{rec.p, reflected + splat3(m.fuzz) * randomUnitVector(seed, pixel, sample, segment),
                            r_in.time};
            attenuation = v3(m.albedo);
            return dot(scattered.direction, rec.normal) > 0;
        }
        case RTB_MAT_DIELECTRIC: {  // material.zig:80-98
            attenuation = v3(1, 1, 1);
            const float refraction_ratio = rec.front_face ? (1.0f / m.ir) : m.ir;
            const V3 unit_direction = unitVector(r_in.direction);
            const float cos_theta = std::fmin(dot(-unit_direction, rec.normal), 1.0f);
            const float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
            const bool cannot_refract = refraction_ratio * sin_theta > 1.0f;
            V3 direction;
            // the reflectance random is word 3 of block 0 of this segment's stream
            if (cannot_refract ||
                reflectance(cos_theta, refraction_ratio) > rng_block(seed, pixel, sample, segment, 0).r[3])
                direction = reflect(unit_direction, rec.normal);
            else
                direction = refract(unit_direction, rec.normal, refraction_ratio);
            scattered = Ray{rec.p, direction, r_in.time};
            return true;
        }
        case RTB_MAT_DIFFUSE_LIGHT:  // material.zig:119-121
            return false;
        case RTB_MAT_ISOTROPIC: {  // material.zig:139-143
            scattered = Ray{rec.p, randomUnitVector(seed, pixel, sample, segment), r_in.time};
            attenuation = textureValue(sc, m.texture, rec.u, rec.v, rec.p);
            return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------- camera.zig
// Camera.getRay for 1-based (x, y), camera.zig:156-180.  Segment 0 of the path's stream:
// block 0 = (jitter x, jitter y, time, -); disk try j = block 1+(j>>1), words 2(j&1), 2(j&1)+1.
Ray getRay(const RtbCamera* cam, uint64_t seed, uint32_t pixel, uint32_t sample) {
    const uint32_t x = pixel % cam->image_width + 1;  // camera.zig:100
    const uint32_t y = pixel / cam->image_width + 1;  // camera.zig:101
    const Block b0 = rng_block(seed, pixel, sample, 0, 0);
    const V3 du = v3(cam->pixel_delta_u), dv = v3(cam->pixel_delta_v);
    const V3 pixel_center = v3(cam->pixel00_loc) + du * splat3((float)x) + dv * splat3((float)y);
    const float px = -0.5f + b0.r[0];  // pixelSampleSquare, camera.zig:162-167
    const float py = -0.5f + b0.r[1];
    const V3 pixel_sample = pixel_center + (splat3(px) * du + splat3(py) * dv);
    V3 ray_origin = v3(cam->center);
    if (!(cam->defocus_angle <= 0)) {  // defocusDiskSample, camera.zig:156-160; vec3.zig:40-45
        V3 p;
        for (uint32_t j = 0;; ++j) {
            const Block b = rng_block(seed, pixel, sample, 0, 1 + (j >> 1));
            const uint32_t w = 2 * (j & 1);
            p = v3(rangeOf(b.r[w], -1, 1), rangeOf(b.r[w + 1], -1, 1), 0);
            if (lengthSquared(p) < 1) break;
        }
        ray_origin = v3(cam->center) + v3(cam->defocus_disk_u) * splat3(p.x) + v3(cam->defocus_disk_v) * splat3(p.y);
    }
    const V3 ray_direction = pixel_sample - ray_origin;
    return Ray{ray_origin, ray_direction, b0.r[2]};
}

V3 background(const RtbCamera* cam, const Ray& r) {
    if (cam->background_mode == RTB_BACKGROUND_SKY) {  // camera.zig:204-206 (legacy, commented out at HEAD)
        const V3 unit_direction = unitVector(r.direction);
        const float a = 0.5f * (unit_direction.y + 1.0f);
        return v3(1, 1, 1) * splat3(1.0f - a) + v3(0.5f, 0.7f, 1.0f) * splat3(a);
    }
    return v3(cam->background);  // camera.zig:207
}

// Camera.rayColor, camera.zig:182-208 — recursive, like the reference.
V3 rayColor(const RtbSceneDesc* sc, const RtbCamera* cam, const Ray& r, uint32_t depth, uint64_t seed,
            uint32_t pixel, uint32_t sample, uint32_t segment, Counters& cn) {
    if (depth <= 0) return v3(0, 0, 0);
    const Interval ray_t{0.001f, kInfinity};
    ++cn.rays;
    HitRecord rec;
    g_ctx = HitCtx{seed, pixel, sample, segment};
    if (worldHit(sc, r, ray_t, rec, cn)) {
        ++cn.hits;
        Ray scattered{v3(0, 0, 0), v3(0, 0, 0), 0};
        V3 attenuation = v3(0, 0, 0);
        const V3 color_from_emission = emitted(sc, rec);
        if (scatter(sc, r, rec, seed, pixel, sample, segment, attenuation, scattered)) {
            const V3 color_from_scatter =
                attenuation * rayColor(sc, cam, scattered, depth - 1, seed, pixel, sample, segment + 1, cn);
            return color_from_emission + color_from_scatter;
        }
        return color_from_emission;
    }
    return background(cam, r);
}

// color.toGamma2 (color.zig:43-62) + writeColor's @intFromFloat (camera.zig:57-65).
// @intFromFloat(NaN) is undefined in the reference; the oracle maps NaN to 0.
inline void quantise(const float* acc, float n, uint8_t* out) {
    const float scale = 1.0f / n;
    const Interval intensity{0, 0.999f};
    for (int c = 0; c < 3; ++c) {
        float x = acc[c];
        x *= scale;
        x = std::sqrt(x);
        float g = 256 * clampI(intensity, x);
        if (!(g == g)) g = 0;
        out[c] = (uint8_t)g;
    }
    out[3] = 255;
}

inline Ray toRay(const RtbRay* r) { return Ray{v3(r->origin), v3(r->direction), r->time}; }
inline void fromRay(const Ray& r, RtbRay* o) {
    store3(o->origin, r.origin);
    store3(o->direction, r.direction);
    o->time = r.time;
    o->t_min = 0.001f;
    o->t_max = kInfinity;
}
inline void fillHit(const HitRecord& rec, bool ok, const Counters& cn, RtbHit* h) {
    std::memset(h, 0, sizeof(*h));
    h->object = ok ? rec.object : -1;
    if (ok) {
        h->front_face = rec.front_face ? 1u : 0u;
        h->t = rec.t;
        store3(h->p, rec.p);
        store3(h->normal, rec.normal);
        h->u = rec.u;
        h->v = rec.v;
    }
    h->n_box_tests = (uint32_t)cn.box;
    h->n_object_tests = (uint32_t)cn.obj;
}

// ---------------------------------------------------------------- host PRNG
inline uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// std.sort.heap (Zig std, not vendored): build a max-heap with sift-down from n/2-1 down to 0,
// then repeatedly swap the root with the last element and sift down.
struct Item {  // one Hittable of world_objects.items with its bounding_box {min xyz, max xyz}
    RtbHittable h;
    float box[6];
};

template <class Less>
void siftDown(Item* items, size_t start, size_t target, size_t end, Less less) {
    size_t cur = target;
    for (;;) {
        size_t child = (cur - start) * 2 + start + 1;
        if (!(child < end)) break;
        const size_t next_child = child + 1;
        if (next_child < end && less(items[child], items[next_child])) child = next_child;
        if (less(items[child], items[cur])) break;
        std::swap(items[child], items[cur]);
        cur = child;
    }
}
template <class Less>
void heapSort(Item* items, size_t a, size_t b, Less less) {
    size_t i = a + (b - a) / 2;
    while (i > a) {
        --i;
        siftDown(items, a, i, b, less);
    }
    i = b;
    while (i > a) {
        --i;
        std::swap(items[a], items[i]);
        siftDown(items, a, a, i, less);
    }
}

struct Builder {
    Item* objs;
    uint64_t* rng;
    RtbBvhNode* nodes;
    int32_t count = 0;

    // bvh.zig:91-103: axis 0 -> x, 1 -> y, anything else -> z
    static bool boxComparator(uint32_t axis, const Item& a, const Item& b) {
        const uint32_t ax = axis == 0 ? 0 : (axis == 1 ? 1 : 2);
        return a.box[ax] < b.box[ax];
    }
    int32_t makeLeaf(uint32_t index) {  // bvh.zig:82-89
        RtbBvhNode& n = nodes[count];
        std::memset(&n, 0, sizeof(n));
        std::memcpy(n.bmin, objs[index].box, sizeof(n.bmin));
        std::memcpy(n.bmax, objs[index].box + 3, sizeof(n.bmax));
        n.left = n.right = -1;
        n.leaf = (int32_t)index;
        return count++;
    }
    int32_t makeNode(int32_t l, int32_t r) {  // bvh.zig:73-80 + Aabb.fromBoxes aabb.zig:28-34
        RtbBvhNode& n = nodes[count];
        std::memset(&n, 0, sizeof(n));
        for (int a = 0; a < 3; ++a) {
            n.bmin[a] = std::fmin(nodes[l].bmin[a], nodes[r].bmin[a]);  // interval.fromIntervals :42-44
            n.bmax[a] = std::fmax(nodes[l].bmax[a], nodes[r].bmax[a]);
        }
        n.left = l;
        n.right = r;
        n.leaf = -1;
        return count++;
    }
    int32_t construct(size_t start, size_t end) {  // bvh.zig:43-71
        int32_t left, right;
        const size_t obj_span = end - start;
        const uint32_t axis = orc_host_random_int_range(rng, 0, 2);  // drawn on every call, :48
        if (obj_span == 1) return makeLeaf((uint32_t)start);
        if (obj_span == 2) {
            if (boxComparator(axis, objs[start], objs[start + 1])) {
                left = makeLeaf((uint32_t)start);
                right = makeLeaf((uint32_t)start + 1);
            } else {
                left = makeLeaf((uint32_t)start + 1);
                right = makeLeaf((uint32_t)start);
            }
        } else {
            heapSort(objs, start, end,
                     [axis](const Item& a, const Item& b) { return boxComparator(axis, a, b); });
            const size_t mid = start + obj_span / 2;
            left = construct(start, mid);
            right = construct(mid, end);
        }
        return makeNode(left, right);
    }
};

}  // namespace

// ============================================================================ C API
extern "C" {

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox(ctr, key, out); }
float orc_u01(uint32_t x) { return u01(x); }

int orc_aabb_hit(const float bmin[3], const float bmax[3], const RtbRay* ray) {
    return aabbHit(bmin, bmax, toRay(ray), Interval{ray->t_min, ray->t_max}) ? 1 : 0;
}

void orc_sphere_uv(const float p[3], float* u, float* v) { getSphereUV(v3(p), u, v); }

int orc_hittable_hit(const RtbSceneDesc* scene, uint32_t index, const RtbRay* ray, RtbHit* hit) {
    HitRecord rec;
    Counters cn;
    cn.obj = 1;
    g_ctx = HitCtx{};
    const bool ok = hittableHit(scene, index, toRay(ray), Interval{ray->t_min, ray->t_max}, rec);
    fillHit(rec, ok, cn, hit);
    return ok ? 1 : 0;
}

void orc_trace_rays(const RtbSceneDesc* scene, const RtbRay* rays, uint64_t n, RtbHit* hits_out) {
    for (uint64_t i = 0; i < n; ++i) {
        HitRecord rec;
        Counters cn;
        g_ctx = HitCtx{0, (uint32_t)i, 0, 1};  // ray queries: medium draws are keyed (seed 0; pixel = ray index)
        const bool ok = worldHit(scene, toRay(&rays[i]), Interval{rays[i].t_min, rays[i].t_max}, rec, cn);
        fillHit(rec, ok, cn, &hits_out[i]);
    }
}

void orc_texture_value(const RtbSceneDesc* scene, uint32_t texture, float u, float v, const float p[3], float out[3]) {
    store3(out, textureValue(scene, texture, u, v, v3(p)));
}
float orc_perlin_noise(const RtbPerlin* perlin, const float p[3]) { return perlinNoise(perlin, v3(p)); }
float orc_perlin_turb(const RtbPerlin* perlin, const float p[3], int depth) { return perlinTurb(perlin, v3(p), depth); }

void orc_get_ray(const RtbCamera* cam, uint64_t seed, uint32_t pixel, uint32_t sample, RtbRay* ray_out) {
    fromRay(getRay(cam, seed, pixel, sample), ray_out);
}

int orc_scatter(const RtbSceneDesc* scene, const RtbRay* r_in, const RtbHit* hit, uint64_t seed, uint32_t pixel,
                uint32_t sample, uint32_t segment, float attenuation[3], RtbRay* scattered) {
    HitRecord rec;
    rec.p = v3(hit->p);
    rec.normal = v3(hit->normal);
    rec.t = hit->t;
    rec.u = hit->u;
    rec.v = hit->v;
    rec.front_face = hit->front_face != 0;
    rec.object = hit->object;
    rec.mat = scene->hittables[hit->object].material;
    V3 att = v3(0, 0, 0);
    Ray sc{v3(0, 0, 0), v3(0, 0, 0), 0};
    const bool ok = scatter(scene, toRay(r_in), rec, seed, pixel, sample, segment, att, sc);
    store3(attenuation, att);
    fromRay(sc, scattered);
    return ok ? 1 : 0;
}

void orc_path_radiance(const RtbSceneDesc* scene, const RtbCamera* cam, uint64_t seed, uint32_t pixel,
                       uint32_t sample, float out_rgb[3]) {
    Counters cn;
    const Ray r = getRay(cam, seed, pixel, sample);
    store3(out_rgb, rayColor(scene, cam, r, cam->max_depth, seed, pixel, sample, 1, cn));
}

int orc_render(const RtbSceneDesc* scene, const RtbCamera* cam, const RtbRenderOptions* options, int n_threads,
               float* accum, uint8_t* rgba, RtbRenderStats* stats) {
    if (!scene || !cam || !options || !accum || n_threads < 1) return RTB_ERR_INVALID_ARGUMENT;
    const uint64_t size = (uint64_t)cam->image_width * cam->image_height;
    uint64_t begin = options->pixel_begin, count = options->pixel_count;
    if (begin == 0 && count == 0) count = size;
    if (begin + count > size) return RTB_ERR_INVALID_ARGUMENT;
    const uint32_t s_begin = options->sample_begin;
    const uint32_t s_count = options->sample_count ? options->sample_count : cam->samples_per_pixel;
    const uint64_t seed = options->seed;
    // startRender: chunk_size = size / number_of_threads (main.zig:319); the last strip also takes
    // the remainder here (the reference leaves size % 8 trailing pixels unrendered).
    const uint64_t chunk = count / (uint64_t)n_threads;
    std::vector<Counters> cns((size_t)n_threads);
    std::vector<std::thread> threads;
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < n_threads; ++t) {
        const uint64_t start_at = begin + (uint64_t)t * chunk;                                 // camera.zig:94
        const uint64_t end_before = (t == n_threads - 1) ? begin + count : start_at + chunk;  // camera.zig:95
        threads.emplace_back([=, &cns]() {
            Counters cn;
            for (uint32_t n = s_begin + 1; n < s_begin + s_count + 1; ++n) {  // camera.zig:98 (1-based count)
                for (uint64_t i = start_at; i < end_before; ++i) {            // camera.zig:99
                    const uint32_t sample = n - 1;
                    const Ray r = getRay(cam, seed, (uint32_t)i, sample);
                    const V3 c = rayColor(scene, cam, r, cam->max_depth, seed, (uint32_t)i, sample, 1, cn);
                    float* px = accum + 4 * i;  // writeColor, camera.zig:54-66
                    px[0] += c.x;
                    px[1] += c.y;
                    px[2] += c.z;
                    px[3] = (float)n;
                    if (rgba) quantise(px, px[3], rgba + 4 * i);
                }
            }
            cns[(size_t)t] = cn;
        });
    }
    for (auto& th : threads) th.join();
    const auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->n_paths = count * s_count;
        for (const auto& c : cns) {
            stats->n_rays += c.rays;
            stats->n_box_tests += c.box;
            stats->n_object_tests += c.obj;
            stats->n_hits += c.hits;
        }
        stats->device_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    }
    return RTB_OK;
}

void orc_resolve(const float* accum, uint8_t* rgba, uint64_t n_pixels, float n_samples_override) {
    for (uint64_t i = 0; i < n_pixels; ++i) {
        const float n = n_samples_override > 0 ? n_samples_override : accum[4 * i + 3];
        quantise(accum + 4 * i, n, rgba + 4 * i);
    }
}

float orc_host_random(uint64_t* state) { return (float)(splitmix64(state) >> 40) * (1.0f / 16777216.0f); }

uint32_t orc_host_random_int_range(uint64_t* state, uint32_t min, uint32_t max) {  // rtweekend.zig:23-27
    const float min_f = (float)min;
    const float max_f = (float)(max + 1);
    return (uint32_t)std::round(rangeOf(orc_host_random(state), min_f, max_f));  // can return max+1
}

void orc_camera_init(const OrcCameraOptions* o, RtbCamera* cam) {  // camera.zig:118-154
    std::memset(cam, 0, sizeof(*cam));
    uint32_t image_height = o->image_height;
    if (image_height == 0) image_height = (uint32_t)std::round((float)o->image_width / o->aspect_ratio);
    if (image_height < 1) image_height = 1;
    cam->image_width = o->image_width;
    cam->image_height = image_height;
    cam->samples_per_pixel = o->samples_per_pixel;
    cam->max_depth = o->max_depth;
    const V3 lookfrom = v3(o->lookfrom), lookat = v3(o->lookat), vup = v3(o->vup);
    const V3 center = lookfrom;
    const float theta = o->vfov * kPi / 180.0f;  // degreesToRadians, rtweekend.zig:10-12
    const float h = std::tan(theta / 2.0f);
    const float viewport_height = 2 * h * o->focus_dist;
    const float viewport_width = viewport_height * ((float)o->image_width / (float)image_height);
    const V3 w = unitVector(lookfrom - lookat);
    const V3 u = unitVector(cross(vup, w));
    const V3 v = cross(w, u);
    const V3 viewport_u = splat3(viewport_width) * u;
    const V3 viewport_v = splat3(viewport_height) * -v;
    const V3 du = viewport_u / splat3((float)o->image_width);
    const V3 dv = viewport_v / splat3((float)image_height);
    const V3 upper_left = center - splat3(o->focus_dist) * w - viewport_u / splat3(2.0f) - viewport_v / splat3(2.0f);
    const V3 pixel00 = upper_left + splat3(0.5f) * (du + dv);
    const float defocus_radius = o->focus_dist * std::tan((o->defocus_angle / 2.0f) * kPi / 180.0f);
    store3(cam->center, center);
    store3(cam->pixel00_loc, pixel00);
    store3(cam->pixel_delta_u, du);
    store3(cam->pixel_delta_v, dv);
    store3(cam->defocus_disk_u, u * splat3(defocus_radius));
    store3(cam->defocus_disk_v, v * splat3(defocus_radius));
    cam->defocus_angle = o->defocus_angle;
    std::memcpy(cam->background, o->background, sizeof(cam->background));
    cam->background_mode = o->background_mode;
}

void orc_sphere_bbox(const float center1[3], const float* center2, float radius, float bmin[3], float bmax[3]) {
    // Sphere.init / initMoving, objects.zig:80-92; Aabb.fromPoints aabb.zig:18-26; fromBoxes :28-34
    const V3 rvec = v3(radius, radius, radius);
    const V3 c1 = v3(center1);
    const V3 lo = c1 - rvec, hi = c1 + rvec;
    float mn[3] = {std::fmin(lo.x, hi.x), std::fmin(lo.y, hi.y), std::fmin(lo.z, hi.z)};
    float mx[3] = {std::fmax(lo.x, hi.x), std::fmax(lo.y, hi.y), std::fmax(lo.z, hi.z)};
    if (center2) {
        const V3 c2 = v3(center2);
        const V3 lo2 = c2 - rvec, hi2 = c2 + rvec;
        const float mn2[3] = {std::fmin(lo2.x, hi2.x), std::fmin(lo2.y, hi2.y), std::fmin(lo2.z, hi2.z)};
        const float mx2[3] = {std::fmax(lo2.x, hi2.x), std::fmax(lo2.y, hi2.y), std::fmax(lo2.z, hi2.z)};
        for (int a = 0; a < 3; ++a) {
            mn[a] = std::fmin(mn[a], mn2[a]);
            mx[a] = std::fmax(mx[a], mx2[a]);
        }
    }
    for (int a = 0; a < 3; ++a) {
        bmin[a] = mn[a];
        bmax[a] = mx[a];
    }
}

void orc_quad_bbox(const float q_[3], const float u_[3], const float v_[3], float bmin[3], float bmax[3]) {
    // Quad.init objects.zig:206-211: fromPoints(q, q+u+v).pad()  (aabb.zig:36-43)
    const V3 q = v3(q_);
    const V3 far = q + v3(u_) + v3(v_);
    const float lo[3] = {std::fmin(q.x, far.x), std::fmin(q.y, far.y), std::fmin(q.z, far.z)};
    const float hi[3] = {std::fmax(q.x, far.x), std::fmax(q.y, far.y), std::fmax(q.z, far.z)};
    const float delta = 0.0001f;
    for (int a = 0; a < 3; ++a) {
        if (hi[a] - lo[a] >= delta) {
            bmin[a] = lo[a];
            bmax[a] = hi[a];
        } else {  // Interval.expand, interval.zig:26-29
            const float padding = delta / 2.0f;
            bmin[a] = lo[a] - padding;
            bmax[a] = hi[a] + padding;
        }
    }
}

void orc_box_bbox(const float a_[3], const float b_[3], float sin_theta, float cos_theta, const float offset[3],
                  float bmin[3], float bmax[3]) {
    // createBox -> HittableList.add (objects.zig:274-277): the list's box starts as Aabb{} = [0,0]^3 and is
    // united with every quad's (padded) box, so it always contains the origin.
    V3 q[6], u[6], v[6];
    boxFaces(v3(a_), v3(b_), q, u, v);
    float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    for (int f = 0; f < 6; ++f) {
        float qmn[3], qmx[3];
        const float qa[3] = {q[f].x, q[f].y, q[f].z}, ua[3] = {u[f].x, u[f].y, u[f].z}, va[3] = {v[f].x, v[f].y, v[f].z};
        orc_quad_bbox(qa, ua, va, qmn, qmx);
        for (int k = 0; k < 3; ++k) {
            mn[k] = std::fmin(mn[k], qmn[k]);
            mx[k] = std::fmax(mx[k], qmx[k]);
        }
    }
    // RotateY.init (objects.zig:360-397): the 8 corners, rotated
    const float inf = kInfinity;
    float rmin[3] = {inf, inf, inf}, rmax[3] = {-inf, -inf, -inf};
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {
                const float i_f = (float)i, j_f = (float)j, k_f = (float)k;
                const float x = i_f * mx[0] + (1 - i_f) * mn[0];
                const float y = j_f * mx[1] + (1 - j_f) * mn[1];
                const float z = k_f * mx[2] + (1 - k_f) * mn[2];
                const float newx = cos_theta * x + sin_theta * z;
                const float newz = -sin_theta * x + cos_theta * z;
                const float tester[3] = {newx, y, newz};
                for (int c = 0; c < 3; ++c) {
                    rmin[c] = std::fmin(rmin[c], tester[c]);
                    rmax[c] = std::fmax(rmax[c], tester[c]);
                }
            }
    // Aabb.fromPoints(min, max) then Translate.init: bbox.add(offset) (objects.zig:315, aabb.zig:51-57)
    for (int c = 0; c < 3; ++c) {
        bmin[c] = std::fmin(rmin[c], rmax[c]) + offset[c];
        bmax[c] = std::fmax(rmin[c], rmax[c]) + offset[c];
    }
}

int32_t orc_bvh_build(RtbHittable* hittables, float* boxes, uint32_t n, uint64_t* rng_state, RtbBvhNode* nodes_out) {
    if (n == 0) return -1;
    std::vector<Item> items(n);
    for (uint32_t i = 0; i < n; ++i) {
        items[i].h = hittables[i];
        std::memcpy(items[i].box, boxes + 6 * (size_t)i, sizeof(items[i].box));
    }
    Builder b{items.data(), rng_state, nodes_out};
    const int32_t root = b.construct(0, n);
    for (uint32_t i = 0; i < n; ++i) {
        hittables[i] = items[i].h;
        std::memcpy(boxes + 6 * (size_t)i, items[i].box, sizeof(items[i].box));
    }
    return root;
}

void orc_perlin_init(uint64_t* rng, RtbPerlin* out) {  // perlin.zig:83-101
    for (int i = 0; i < 256; ++i) {
        const float x = rangeOf(orc_host_random(rng), -1, 1);  // vec3.randomRange, vec3.zig:51-57
        const float y = rangeOf(orc_host_random(rng), -1, 1);
        const float z = rangeOf(orc_host_random(rng), -1, 1);
        store3(out->ranvec[i], unitVector(v3(x, y, z)));
    }
    uint16_t* perms[3] = {out->perm_x, out->perm_y, out->perm_z};
    for (int t = 0; t < 3; ++t) {  // perlin_generate_perm :18-28, permute :8-16
        uint16_t* p = perms[t];
        for (int i = 0; i < 256; ++i) p[i] = (uint16_t)i;
        for (int i = 255; i > 0; --i) {
            uint32_t target = orc_host_random_int_range(rng, 0, (uint32_t)i);
            if (target > 255) target = 255;  // reference reads p[256] here (out of bounds)
            const uint16_t tmp = p[i];
            p[i] = p[target];
            p[target] = tmp;
        }
    }
}

}  // extern "C"
