// rtw_host.hpp — host-side mirror of the reference's scene-construction and camera API.
//
// In a real deployment this layer stays in Zig (BASELINE.json north_star: "the scene-construction
// and camera API stays in Zig"); there is no Zig toolchain in this image, so the same interface is
// written in C++ above the C ABI of include/rtb.h, with the reference's names and argument
// meaning, so that tests and the bench read like the reference's own call sites:
//
//   Camera{...options...}.init()                    src/camera.zig:69-91, :118-154
//   Sphere::init / initMoving, Quad::init           src/objects.zig:80-92, :206-211
//   Lambertian::init / fromColor, Metal::fromColor, Dielectric::init, DiffuseLight::...
//                                                   src/material.zig:32-126
//   SolidColor / CheckerTexture / ImageTexture / NoiseTexture ::init   src/textures.zig:29-124
//   Perlin::init                                    src/perlin.zig:83-101
//   BVHTree::init(objects, start, end)              src/bvh.zig:22-29, :43-103
//   SharedStateImageWriter{buffer, texture_buffer}  src/camera.zig:22-67
//   Camera::render(world, writer)                   src/camera.zig:93-116  ← calls rtb_render
//   scene builders generateWorld / earthWorld / twoSpheresWorld / twoPerlinWorld
//                                                   src/main.zig:88-125, :253-312
//
// None of this runs per pixel: it produces the inputs of the hot path and lowers
// the pointer graph into the POD arrays of RtbSceneDesc (`World::lower`).
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/rtb.h"

namespace rtw {

struct Vec3 {
    float x = 0, y = 0, z = 0;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
Vec3 operator+(Vec3 a, Vec3 b);
Vec3 operator-(Vec3 a, Vec3 b);
Vec3 operator*(Vec3 a, Vec3 b);
Vec3 operator/(Vec3 a, Vec3 b);
Vec3 operator-(Vec3 a);
Vec3 splat3(float s);
float dot(Vec3 a, Vec3 b);
Vec3 cross(Vec3 a, Vec3 b);
float length(Vec3 a);
Vec3 unitVector(Vec3 a);

// Seedable stand-in for std.crypto.random on the HOST side (scene, BVH axes, perlin tables).
// SplitMix64; randomDouble = (next >> 40) * 2^-24.
struct HostRng {
    uint64_t state;
    explicit HostRng(uint64_t seed) : state(seed) {}
    float randomDouble();                              // rtweekend.zig:14-16
    float randomDoubleRange(float min, float max);     // rtweekend.zig:18-20
    uint32_t randomIntRange(uint32_t min, uint32_t max);  // rtweekend.zig:23-27 (may return max+1)
    Vec3 random();                                     // vec3.zig:47-49
    Vec3 randomRange(float min, float max);            // vec3.zig:51-57
};

struct Interval {
    float min, max;
};
struct Aabb {  // src/aabb.zig:13-57
    Interval x{0, 0}, y{0, 0}, z{0, 0};
    static Aabb fromPoints(Vec3 a, Vec3 b);
    static Aabb fromBoxes(const Aabb& a, const Aabb& b);
    Aabb pad() const;
    Interval axis(uint32_t n) const;
};

struct Perlin {  // src/perlin.zig:76-101
    RtbPerlin tables;
    static std::shared_ptr<Perlin> init(HostRng& rng);
};

struct Image {  // zstbi.Image after loadFromFile(path, 4)
    uint32_t width = 0, height = 0, bytes_per_row = 0;
    std::vector<uint8_t> data;
};

struct Texture {  // src/textures.zig:10-27
    uint32_t type = RTB_TEX_SOLID;
    Vec3 color_value;           // solid; checker even
    Vec3 odd;                   // checker odd
    float inv_scale = 1;        // checker
    float scale = 1;            // noise
    uint32_t image_index = 0;   // image
    std::shared_ptr<Perlin> noise;
};
struct SolidColor {
    static Texture init(Vec3 color);
};
struct CheckerTexture {
    static Texture init(float scale, const Texture& even, const Texture& odd);
};
struct ImageTexture {
    static Texture init(uint32_t image_index);
};
struct NoiseTexture {
    static Texture init(float scale, HostRng& rng);
};

struct Material {  // src/material.zig:11-16
    uint32_t type = RTB_MAT_LAMBERTIAN;
    Texture texture;  // lambertian albedo / diffuse_light emit / isotropic albedo
    Vec3 albedo;      // metal
    float fuzz = 1;
    float ir = 1;
};
struct Lambertian {
    static Material init(const Texture& t);
    static Material fromColor(Vec3 c);
};
struct Metal {
    static Material fromColor(Vec3 c, float f);
};
struct Dielectric {
    static Material init(float ir);
};
struct DiffuseLight {
    static Material init(const Texture& t);
    static Material fromColor(Vec3 c);
};

struct Hittable {  // src/objects.zig:39-47 (sphere, quad, and list/rotate_y/translate as used for boxes)
    uint32_t type = RTB_HITTABLE_SPHERE;
    Vec3 a, b, c;  // sphere: center1, center_vec, -; quad: q, u, v; box: corner a, corner b, Translate.offset
    float radius = 0;
    bool is_moving = false;
    float sin_theta = 0, cos_theta = 1;  // box: RotateY (identity until RotateY::init is applied)
    bool rotated = false, translated = false;
    Material mat;
    Aabb bounding_box;
    // General instancing (objects.zig:264-443): a Translate / RotateY / HittableList / ConstantMedium that wraps
    // ANYTHING holds its wrapped objects by value here (the reference copies `obj` into allocator.create(Hittable)).
    std::vector<Hittable> children;
    const Aabb& boundingBox() const { return bounding_box; }
};
struct Sphere {
    static Hittable init(Vec3 center1, float radius, const Material& mat);
    static Hittable initMoving(Vec3 center1, Vec3 center2, float radius, const Material& mat);
};
struct Quad {
    static Hittable init(Vec3 q, Vec3 u, Vec3 v, const Material& mat);
};
// createBox (src/objects.zig:510-532): a HittableList of six quads.  The reference wraps it as
// Translate.init(RotateY.init(box, angle), offset) (src/main.zig:182-190); the three calls below compose the
// same way on one flat object (RotateY must come before Translate, each at most once) and maintain the
// bounding box exactly like HittableList.add / RotateY.init / Translate.init do.
Hittable createBox(Vec3 a, Vec3 b, const Material& mat);
struct RotateY {
    static Hittable init(const Hittable& box, float angle_degrees);  // src/objects.zig:354-397
};
struct Translate {
    static Hittable init(const Hittable& box, Vec3 offset);  // src/objects.zig:314-319
};
// The general forms.  Translate::init / RotateY::init above keep a createBox-derived box as ONE flat record (the fast
// path HEAD's scenes use); these wrap any hittable, any number of times, in any order.
struct HittableList {  // src/objects.zig:264-305; the list's box starts as Aabb{} = the origin (:266, :274-277)
    static Hittable init(const std::vector<Hittable>& objects);
};
struct TranslateAny {
    static Hittable init(const Hittable& obj, Vec3 offset);  // src/objects.zig:314-319
};
struct RotateYAny {
    static Hittable init(const Hittable& obj, float angle_degrees);  // src/objects.zig:354-397
};
struct ConstantMediumOf {
    static Hittable initFromColor(const Hittable& boundary, float d, Vec3 c);  // src/objects.zig:450-460
};
struct Isotropic {
    static Material init(const Texture& t);  // src/material.zig:131-133
    static Material fromColor(Vec3 c);       // src/material.zig:135-137
};
struct ConstantMedium {
    // src/objects.zig:450-452: boundary (a box instance here, as in cornellBoxSmoke), neg_inv_density = -1/d,
    // phase_function = Isotropic(SolidColor(c)); boundingBox() = the boundary's (:458-460).
    static Hittable initFromColor(const Hittable& boundary, float d, Vec3 c);
};

using ObjectList = std::vector<Hittable>;

struct BVHNode {  // src/bvh.zig:106-110
    int32_t leaf = -1;  // index into the object list (the reference holds a *const Hittable)
    std::unique_ptr<BVHNode> left, right;
    Aabb bounding_box;
};
struct BVHTree {  // src/bvh.zig:17-29
    std::unique_ptr<BVHNode> root;
    Aabb bounding_box;
    // Sorts objects[start, end) in place like the reference (random axis, box-min order, median
    // split).  Keys are sorted through an index permutation, then the objects are permuted once.
    static BVHTree init(ObjectList& objects, size_t start, size_t end, HostRng& rng);
};

// The lowered, upload-ready form of a world: owns the POD arrays an RtbSceneDesc points into.
struct LoweredScene {
    std::vector<RtbBvhNode> nodes;
    std::vector<RtbHittable> hittables;
    std::vector<RtbMaterial> materials;
    std::vector<RtbTexture> textures;
    std::vector<RtbPerlin> perlins;
    std::vector<RtbImage> image_descs;
    std::vector<Image> images;
    RtbSceneDesc desc{};
    void finalize();  // (re)points desc at the vectors
};

struct World {  // Hittable{ .tree = BVHTree } + the list it indexes (src/main.zig:309-311)
    ObjectList objects;
    BVHTree tree;
    std::vector<Image> images;
    // DFS over BVHNode pointers -> linear RtbBvhNode array (pre-order), unions -> tagged PODs.
    std::unique_ptr<LoweredScene> lower() const;
};

struct SharedStateImageWriter {  // src/camera.zig:22-45
    std::vector<float> buffer;           // 4 floats per pixel: (sum R, sum G, sum B, n)
    std::vector<uint8_t> texture_buffer; // RGBA8
    uint32_t width = 0, height = 0;
    static SharedStateImageWriter init(uint32_t image_width, uint32_t image_height);
    void scrub();
};

struct Camera {  // src/camera.zig:69-91
    float aspect_ratio = 16.0f / 9.0f;
    uint16_t image_width = 800;
    uint16_t image_height = 0;
    uint32_t size = 0;
    Vec3 center, pixel00_loc, pixel_delta_u, pixel_delta_v;
    uint16_t samples_per_pixel = 100;
    uint8_t max_depth = 16;
    Vec3 background{0, 0, 0};
    float vfov = 20;
    Vec3 lookfrom{13, 2, 3};
    Vec3 lookat{0, 0, 0};
    Vec3 vup{0, 1, 0};
    Vec3 u, v, w;
    float defocus_angle = 0.6f;
    float focus_dist = 10;
    Vec3 defocus_disk_u, defocus_disk_v;
    // Not a reference field: selects the legacy sky gradient (src/camera.zig:204-206).
    uint32_t background_mode = RTB_BACKGROUND_SOLID;

    void init();                 // src/camera.zig:118-154
    RtbCamera lowered() const;   // the derived fields the device reads
    // Drop-in for the 8 x Camera.render threads (src/camera.zig:93-116, src/main.zig:318-324):
    // one call renders every strip on the GPU and fills writer.buffer / texture_buffer.
    int render(RtbScene* scene, SharedStateImageWriter& writer, const RtbRenderOptions* options = nullptr,
               RtbRenderStats* stats = nullptr) const;
};

// Scene builders (src/main.zig).  `images` plays the role of the `images: ArrayList(zstbi.Image)`
// argument; decoding stays outside (src/main.zig:1120-1125).
struct Book1Options {
    bool checker_ground = false;  // HEAD: checker 0.32 ground (main.zig:257-260); Book-1: solid (0.5,0.5,0.5)
    bool earth_sphere = false;    // HEAD: earthmap sphere at (-4,1,0) (main.zig:299-303); Book-1: lambertian (0.4,0.2,0.1)
    bool moving_spheres = true;   // HEAD: diffuse spheres are initMoving (main.zig:279-281)
};
World generateWorld(HostRng& scene_rng, HostRng& bvh_rng, const Book1Options& opt, std::vector<Image> images);
World earthWorld(HostRng& bvh_rng, std::vector<Image> images);           // main.zig:88-99
World twoSpheresWorld(HostRng& bvh_rng);                                 // main.zig:101-113
World twoPerlinWorld(HostRng& perlin_rng, HostRng& bvh_rng);             // main.zig:115-125
World cornellBox(HostRng& bvh_rng);                                      // main.zig:168-205 (objects only)
World cornellBoxSmoke(HostRng& bvh_rng);                                 // main.zig:207-251 (objects only)
World quadsWorld(HostRng& bvh_rng);                                      // main.zig:127-143
World simpleLightWorld(HostRng& perlin_rng, HostRng& bvh_rng);           // main.zig:145-166
// BASELINE config 3: the three textured worlds side by side in one BVH.
World texturedWorld(HostRng& perlin_rng, HostRng& bvh_rng, std::vector<Image> images);
// BASELINE config 4: n random spheres (80/15/5 % lambertian/metal/dielectric) + ground sphere.
World randomSpheresWorld(HostRng& scene_rng, HostRng& bvh_rng, uint32_t n);

// P3 PPM like src/stdout.zig:5-18 but from the RGBA8 texture buffer (values already <= 255).
bool writePpm(const std::string& path, const uint8_t* rgba, uint32_t width, uint32_t height);
// 8-bit RGB PNG of the same buffer (stored deflate blocks; no third-party encoder).
bool writePng(const std::string& path, const uint8_t* rgba, uint32_t width, uint32_t height);

}  // namespace rtw
