// rtw_host.cpp — host-side mirror of the reference's scene/camera API (see rtw_host.hpp).
// Product code: must not depend on oracle/.
#include "rtw_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>

namespace rtw {

// ------------------------------------------------------------------ vec3.zig
Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
Vec3 operator*(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
Vec3 operator/(Vec3 a, Vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
Vec3 splat3(float s) { return {s, s, s}; }
float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
float length(Vec3 a) { return std::sqrt(dot(a, a)); }
Vec3 unitVector(Vec3 a) { return a / splat3(length(a)); }

static const float kPi = 3.1415926535897932385f;  // rtweekend.zig:4

// ------------------------------------------------------------------ host RNG
float HostRng::randomDouble() {
    uint64_t z = (state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}
float HostRng::randomDoubleRange(float min, float max) { return min + (max - min) * randomDouble(); }
uint32_t HostRng::randomIntRange(uint32_t min, uint32_t max) {
    return (uint32_t)std::round(randomDoubleRange((float)min, (float)(max + 1)));
}
Vec3 HostRng::random() {
    Vec3 r;
    r.x = randomDouble();
    r.y = randomDouble();
    r.z = randomDouble();
    return r;
}
Vec3 HostRng::randomRange(float min, float max) {
    Vec3 r;
    r.x = randomDoubleRange(min, max);
    r.y = randomDoubleRange(min, max);
    r.z = randomDoubleRange(min, max);
    return r;
}

// ------------------------------------------------------------------ aabb.zig
Aabb Aabb::fromPoints(Vec3 a, Vec3 b) {
    Aabb r;
    r.x = {std::fmin(a.x, b.x), std::fmax(a.x, b.x)};
    r.y = {std::fmin(a.y, b.y), std::fmax(a.y, b.y)};
    r.z = {std::fmin(a.z, b.z), std::fmax(a.z, b.z)};
    return r;
}
Aabb Aabb::fromBoxes(const Aabb& a, const Aabb& b) {
    Aabb r;
    r.x = {std::fmin(a.x.min, b.x.min), std::fmax(a.x.max, b.x.max)};
    r.y = {std::fmin(a.y.min, b.y.min), std::fmax(a.y.max, b.y.max)};
    r.z = {std::fmin(a.z.min, b.z.min), std::fmax(a.z.max, b.z.max)};
    return r;
}
static Interval padInterval(Interval i) {  // aabb.zig:36-43, interval.zig:26-29
    const float delta = 0.0001f;
    if (i.max - i.min >= delta) return i;
    const float padding = delta / 2.0f;
    return {i.min - padding, i.max + padding};
}
Aabb Aabb::pad() const {
    Aabb r;
    r.x = padInterval(x);
    r.y = padInterval(y);
    r.z = padInterval(z);
    return r;
}
Interval Aabb::axis(uint32_t n) const {
    if (n == 1) return y;
    if (n == 2) return z;
    return x;
}

// ------------------------------------------------------------------ perlin.zig:8-28, :83-101
std::shared_ptr<Perlin> Perlin::init(HostRng& rng) {
    auto p = std::make_shared<Perlin>();
    for (int i = 0; i < 256; ++i) {
        const Vec3 v = unitVector(rng.randomRange(-1, 1));
        p->tables.ranvec[i][0] = v.x;
        p->tables.ranvec[i][1] = v.y;
        p->tables.ranvec[i][2] = v.z;
    }
    uint16_t* perms[3] = {p->tables.perm_x, p->tables.perm_y, p->tables.perm_z};
    for (uint16_t* perm : perms) {
        for (int i = 0; i < 256; ++i) perm[i] = (uint16_t)i;
        for (int i = 255; i > 0; --i) {
            // randomIntRange(0, i) can return i+1; for i == 255 the reference indexes p[256]
            // (out of bounds) — clamped here.
            const uint32_t target = std::min<uint32_t>(rng.randomIntRange(0, (uint32_t)i), 255u);
            std::swap(perm[i], perm[target]);
        }
    }
    return p;
}

// ------------------------------------------------------------------ textures.zig
Texture SolidColor::init(Vec3 color) {
    Texture t;
    t.type = RTB_TEX_SOLID;
    t.color_value = color;
    return t;
}
Texture CheckerTexture::init(float scale, const Texture& even, const Texture& odd) {
    Texture t;
    t.type = RTB_TEX_CHECKER;
    t.inv_scale = 1.0f / scale;
    t.color_value = even.color_value;
    t.odd = odd.color_value;
    return t;
}
Texture ImageTexture::init(uint32_t image_index) {
    Texture t;
    t.type = RTB_TEX_IMAGE;
    t.image_index = image_index;
    return t;
}
Texture NoiseTexture::init(float scale, HostRng& rng) {
    Texture t;
    t.type = RTB_TEX_NOISE;
    t.scale = scale;
    t.noise = Perlin::init(rng);
    return t;
}

// ------------------------------------------------------------------ material.zig
Material Lambertian::init(const Texture& t) {
    Material m;
    m.type = RTB_MAT_LAMBERTIAN;
    m.texture = t;
    return m;
}
Material Lambertian::fromColor(Vec3 c) { return init(SolidColor::init(c)); }
Material Metal::fromColor(Vec3 c, float f) {
    Material m;
    m.type = RTB_MAT_METAL;
    m.albedo = c;
    m.fuzz = f < 1 ? f : 1;
    return m;
}
Material Dielectric::init(float ir) {
    Material m;
    m.type = RTB_MAT_DIELECTRIC;
    m.ir = ir;
    return m;
}
Material DiffuseLight::init(const Texture& t) {
    Material m;
    m.type = RTB_MAT_DIFFUSE_LIGHT;
    m.texture = t;
    return m;
}
Material DiffuseLight::fromColor(Vec3 c) { return init(SolidColor::init(c)); }

// ------------------------------------------------------------------ objects.zig
Hittable Sphere::init(Vec3 center1, float radius, const Material& mat) {
    Hittable h;
    h.type = RTB_HITTABLE_SPHERE;
    h.a = center1;
    h.radius = radius;
    h.mat = mat;
    const Vec3 rvec{radius, radius, radius};
    h.bounding_box = Aabb::fromPoints(center1 - rvec, center1 + rvec);
    return h;
}
Hittable Sphere::initMoving(Vec3 center1, Vec3 center2, float radius, const Material& mat) {
    Hittable h;
    h.type = RTB_HITTABLE_SPHERE;
    h.a = center1;
    h.b = center2 - center1;
    h.is_moving = true;
    h.radius = radius;
    h.mat = mat;
    const Vec3 rvec{radius, radius, radius};
    const Aabb box1 = Aabb::fromPoints(center1 - rvec, center1 + rvec);
    const Aabb box2 = Aabb::fromPoints(center2 - rvec, center2 + rvec);
    h.bounding_box = Aabb::fromBoxes(box1, box2);
    return h;
}
Hittable Quad::init(Vec3 q, Vec3 u, Vec3 v, const Material& mat) {
    Hittable h;
    h.type = RTB_HITTABLE_QUAD;
    h.a = q;
    h.b = u;
    h.c = v;
    h.mat = mat;
    h.bounding_box = Aabb::fromPoints(q, q + u + v).pad();
    return h;
}

// createBox (objects.zig:510-532) -> HittableList.add (:274-277): the list's box starts as Aabb{} (the origin)
// and is united with each quad's padded box.  (The reference adds the z = min face twice and no z = max face;
// the box of the list is the same either way.)
Hittable createBox(Vec3 a, Vec3 b, const Material& mat) {
    Hittable h;
    h.type = RTB_HITTABLE_BOX;
    h.a = a;
    h.b = b;
    h.mat = mat;
    const Vec3 mn{std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)};
    const Vec3 mx{std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)};
    const Vec3 dx{mx.x - mn.x, 0, 0}, dy{0, mx.y - mn.y, 0}, dz{0, 0, mx.z - mn.z};
    const Vec3 q[6] = {{mn.x, mn.y, mn.z}, {mx.x, mn.y, mx.z}, {mx.x, mn.y, mn.z},
                       {mn.x, mn.y, mn.z}, {mn.x, mx.y, mx.z}, {mn.x, mn.y, mn.z}};
    const Vec3 u[6] = {dx, -dz, -dx, dz, dx, dx};
    const Vec3 v[6] = {dy, dy, dy, dy, -dz, dz};
    Aabb box;  // Aabb{}: x = y = z = [0, 0]
    for (int f = 0; f < 6; ++f) box = Aabb::fromBoxes(box, Aabb::fromPoints(q[f], q[f] + u[f] + v[f]).pad());
    h.bounding_box = box;
    return h;
}

Hittable RotateY::init(const Hittable& box, float angle_degrees) {  // objects.zig:354-397
    Hittable h = box;
    const float radians = angle_degrees * kPi / 180.0f;
    h.sin_theta = std::sin(radians);
    h.cos_theta = std::cos(radians);
    h.rotated = true;
    const Aabb& bbox = box.bounding_box;
    const float inf = INFINITY;
    float mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {
                const float i_f = (float)i, j_f = (float)j, k_f = (float)k;
                const float x = i_f * bbox.x.max + (1 - i_f) * bbox.x.min;
                const float y = j_f * bbox.y.max + (1 - j_f) * bbox.y.min;
                const float z = k_f * bbox.z.max + (1 - k_f) * bbox.z.min;
                const float newx = h.cos_theta * x + h.sin_theta * z;
                const float newz = -h.sin_theta * x + h.cos_theta * z;
                const float tester[3] = {newx, y, newz};
                for (int c = 0; c < 3; ++c) {
                    mn[c] = std::fmin(mn[c], tester[c]);
                    mx[c] = std::fmax(mx[c], tester[c]);
                }
            }
    h.bounding_box = Aabb::fromPoints({mn[0], mn[1], mn[2]}, {mx[0], mx[1], mx[2]});
    return h;
}

Hittable Translate::init(const Hittable& box, Vec3 offset) {  // objects.zig:314-319, aabb.zig:51-57
    Hittable h = box;
    h.c = offset;
    h.translated = true;
    h.bounding_box.x = {box.bounding_box.x.min + offset.x, box.bounding_box.x.max + offset.x};
    h.bounding_box.y = {box.bounding_box.y.min + offset.y, box.bounding_box.y.max + offset.y};
    h.bounding_box.z = {box.bounding_box.z.min + offset.z, box.bounding_box.z.max + offset.z};
    return h;
}

// ---- general instancing: wrappers around any hittable --------------------------------------------------------
Hittable HittableList::init(const std::vector<Hittable>& objects) {  // objects.zig:269-277
    Hittable h;
    h.type = RTB_HITTABLE_LIST;
    Aabb box;  // Aabb{}: [0, 0]^3 — the reference's list box always contains the origin
    for (const Hittable& o : objects) {
        h.children.push_back(o);
        box = Aabb::fromBoxes(box, o.bounding_box);
    }
    h.bounding_box = box;
    return h;
}
Hittable TranslateAny::init(const Hittable& obj, Vec3 offset) {  // objects.zig:314-319, aabb.zig:51-57
    Hittable h;
    h.type = RTB_HITTABLE_TRANSLATE;
    h.a = offset;
    h.children.push_back(obj);
    h.bounding_box.x = {obj.bounding_box.x.min + offset.x, obj.bounding_box.x.max + offset.x};
    h.bounding_box.y = {obj.bounding_box.y.min + offset.y, obj.bounding_box.y.max + offset.y};
    h.bounding_box.z = {obj.bounding_box.z.min + offset.z, obj.bounding_box.z.max + offset.z};
    return h;
}
Hittable RotateYAny::init(const Hittable& obj, float angle_degrees) {  // objects.zig:354-397
    Hittable flat = obj;  // reuse RotateY::init's box arithmetic (it only reads the bounding box and the angle)
    flat.children.clear();
    const Hittable r = RotateY::init(flat, angle_degrees);
    Hittable h;
    h.type = RTB_HITTABLE_ROTATE_Y;
    h.sin_theta = r.sin_theta;
    h.cos_theta = r.cos_theta;
    h.bounding_box = r.bounding_box;
    h.children.push_back(obj);
    return h;
}
Hittable ConstantMediumOf::initFromColor(const Hittable& boundary, float d, Vec3 c) {  // objects.zig:450-460
    Hittable h;
    h.type = RTB_HITTABLE_MEDIUM_OF;
    h.radius = -1.0f / d;
    h.mat = Isotropic::fromColor(c);
    h.bounding_box = boundary.bounding_box;
    h.children.push_back(boundary);
    return h;
}

Material Isotropic::init(const Texture& t) {
    Material m;
    m.type = RTB_MAT_ISOTROPIC;
    m.texture = t;
    return m;
}
Material Isotropic::fromColor(Vec3 c) { return init(SolidColor::init(c)); }

Hittable ConstantMedium::initFromColor(const Hittable& boundary, float d, Vec3 c) {  // objects.zig:450-452
    Hittable h = boundary;  // keeps the box fields and the boundary's bounding box (:458-460)
    h.type = RTB_HITTABLE_CONSTANT_MEDIUM;
    h.radius = -1.0f / d;   // neg_inv_density
    h.mat = Isotropic::fromColor(c);
    return h;
}

// ------------------------------------------------------------------ bvh.zig:43-103
namespace {
struct TreeBuilder {
    const ObjectList& objects;
    std::vector<uint32_t>& perm;  // perm[i] = original index of the object that ends up at slot i
    HostRng& rng;

    float key(uint32_t slot, uint32_t axis) const { return objects[perm[slot]].bounding_box.axis(axis).min; }
    static uint32_t comparatorAxis(uint32_t axis) { return axis == 0 ? 0u : (axis == 1 ? 1u : 2u); }  // :95-103

    std::unique_ptr<BVHNode> leaf(uint32_t slot) const {
        auto n = std::make_unique<BVHNode>();
        n->leaf = (int32_t)slot;
        n->bounding_box = objects[perm[slot]].bounding_box;
        return n;
    }
    std::unique_ptr<BVHNode> construct(size_t start, size_t end) {
        const size_t span = end - start;
        const uint32_t axis = comparatorAxis(rng.randomIntRange(0, 2));  // drawn on every call (:48)
        if (span == 1) return leaf((uint32_t)start);
        std::unique_ptr<BVHNode> left, right;
        if (span == 2) {
            if (key((uint32_t)start, axis) < key((uint32_t)start + 1, axis)) {
                left = leaf((uint32_t)start);
                right = leaf((uint32_t)start + 1);
            } else {
                left = leaf((uint32_t)start + 1);
                right = leaf((uint32_t)start);
            }
        } else {
            const ObjectList& objs = objects;
            std::sort(perm.begin() + (long)start, perm.begin() + (long)end, [&objs, axis](uint32_t a, uint32_t b) {
                return objs[a].bounding_box.axis(axis).min < objs[b].bounding_box.axis(axis).min;
            });
            const size_t mid = start + span / 2;
            left = construct(start, mid);
            right = construct(mid, end);
        }
        auto n = std::make_unique<BVHNode>();
        n->bounding_box = Aabb::fromBoxes(left->bounding_box, right->bounding_box);
        n->left = std::move(left);
        n->right = std::move(right);
        return n;
    }
};
}  // namespace

BVHTree BVHTree::init(ObjectList& objects, size_t start, size_t end, HostRng& rng) {
    BVHTree t;
    if (end <= start) return t;
    std::vector<uint32_t> perm(objects.size());
    std::iota(perm.begin(), perm.end(), 0u);
    TreeBuilder b{objects, perm, rng};
    t.root = b.construct(start, end);
    t.bounding_box = t.root->bounding_box;
    ObjectList sorted;
    sorted.reserve(objects.size());
    for (uint32_t src : perm) sorted.push_back(objects[src]);
    objects.swap(sorted);
    return t;
}

// ------------------------------------------------------------------ lowering
void LoweredScene::finalize() {
    image_descs.clear();
    for (const Image& im : images) {
        RtbImage d{};
        d.width = im.width;
        d.height = im.height;
        d.bytes_per_row = im.bytes_per_row;
        d.data = im.data.data();
        image_descs.push_back(d);
    }
    desc = RtbSceneDesc{};
    desc.abi_version = RTB_ABI_VERSION;
    desc.n_nodes = (uint32_t)nodes.size();
    desc.n_hittables = (uint32_t)hittables.size();
    desc.n_materials = (uint32_t)materials.size();
    desc.n_textures = (uint32_t)textures.size();
    desc.n_perlins = (uint32_t)perlins.size();
    desc.n_images = (uint32_t)image_descs.size();
    desc.root = nodes.empty() ? -1 : 0;
    desc.nodes = nodes.data();
    desc.hittables = hittables.data();
    desc.materials = materials.data();
    desc.textures = textures.data();
    desc.perlins = perlins.data();
    desc.images = image_descs.data();
}

static void put3(float* o, Vec3 v) {
    o[0] = v.x;
    o[1] = v.y;
    o[2] = v.z;
}

std::unique_ptr<LoweredScene> World::lower() const {
    auto ls = std::make_unique<LoweredScene>();
    ls->images = images;
    std::map<const Perlin*, uint32_t> perlin_ids;
    auto lowerTexture = [&](const Texture& t) -> uint32_t {
        RtbTexture r{};
        r.type = t.type;
        put3(r.color, t.color_value);
        put3(r.color2, t.odd);
        if (t.type == RTB_TEX_CHECKER) r.scale = t.inv_scale;
        if (t.type == RTB_TEX_IMAGE) r.index = t.image_index;
        if (t.type == RTB_TEX_NOISE) {
            r.scale = t.scale;
            auto it = perlin_ids.find(t.noise.get());
            if (it == perlin_ids.end()) {
                it = perlin_ids.emplace(t.noise.get(), (uint32_t)ls->perlins.size()).first;
                ls->perlins.push_back(t.noise->tables);
            }
            r.index = it->second;
        }
        ls->textures.push_back(r);
        return (uint32_t)ls->textures.size() - 1;
    };
    // Top-level objects keep their position in the list (= the object index BVH leaves and rtb_trace_rays use);
    // the objects wrapped by Translate / RotateY / HittableList / ConstantMedium follow, each wrapper's children
    // contiguous and after the wrapper itself (include/rtb.h).
    std::vector<const Hittable*> order;
    for (const Hittable& h : objects) order.push_back(&h);
    ls->hittables.resize(order.size());
    for (size_t i = 0; i < order.size(); ++i) {   // `order` grows while children are appended
        const Hittable& h = *order[i];
        RtbHittable r{};
        r.type = h.type;
        r.is_moving = h.is_moving ? 1u : 0u;
        r.radius = h.radius;
        r.sin_theta = h.sin_theta;
        r.cos_theta = h.cos_theta;
        put3(r.a, h.a);
        put3(r.b, h.b);
        put3(r.c, h.c);
        const bool has_material = h.type != RTB_HITTABLE_TRANSLATE && h.type != RTB_HITTABLE_ROTATE_Y && h.type != RTB_HITTABLE_LIST;
        if (has_material) {
            RtbMaterial m{};
            m.type = h.mat.type;
            put3(m.albedo, h.mat.albedo);
            m.fuzz = h.mat.fuzz;
            m.ir = h.mat.ir;
            if (m.type == RTB_MAT_LAMBERTIAN || m.type == RTB_MAT_DIFFUSE_LIGHT || m.type == RTB_MAT_ISOTROPIC)
                m.texture = lowerTexture(h.mat.texture);
            ls->materials.push_back(m);
            r.material = (uint32_t)ls->materials.size() - 1;
        }
        if (!h.children.empty()) {
            r.child = (uint32_t)order.size();
            if (h.type == RTB_HITTABLE_LIST) r.material = (uint32_t)h.children.size();
            for (const Hittable& c : h.children) order.push_back(&c);
            ls->hittables.resize(order.size());
        }
        ls->hittables[i] = r;
    }
    // Pre-order walk with an explicit stack (the million-sphere tree is only ~21 deep, but the
    // Zig shim walks arbitrary pointer graphs the same way).
    struct Pending {
        const BVHNode* node;
        int32_t parent;
        bool is_right;
    };
    std::vector<Pending> stack;
    if (tree.root) stack.push_back({tree.root.get(), -1, false});
    while (!stack.empty()) {
        const Pending p = stack.back();
        stack.pop_back();
        const int32_t me = (int32_t)ls->nodes.size();
        RtbBvhNode n{};
        n.bmin[0] = p.node->bounding_box.x.min;
        n.bmin[1] = p.node->bounding_box.y.min;
        n.bmin[2] = p.node->bounding_box.z.min;
        n.bmax[0] = p.node->bounding_box.x.max;
        n.bmax[1] = p.node->bounding_box.y.max;
        n.bmax[2] = p.node->bounding_box.z.max;
        n.left = n.right = -1;
        n.leaf = p.node->leaf;
        ls->nodes.push_back(n);
        if (p.parent >= 0) {
            if (p.is_right)
                ls->nodes[(size_t)p.parent].right = me;
            else
                ls->nodes[(size_t)p.parent].left = me;
        }
        if (p.node->leaf < 0) {
            stack.push_back({p.node->right.get(), me, true});
            stack.push_back({p.node->left.get(), me, false});
        }
    }
    ls->finalize();
    return ls;
}

// ------------------------------------------------------------------ camera.zig
SharedStateImageWriter SharedStateImageWriter::init(uint32_t image_width, uint32_t image_height) {
    SharedStateImageWriter w;
    w.width = image_width;
    w.height = image_height;
    w.buffer.resize((size_t)image_width * image_height * 4);
    w.texture_buffer.resize((size_t)image_width * image_height * 4);
    w.scrub();
    return w;
}
void SharedStateImageWriter::scrub() {
    for (size_t i = 0; i < (size_t)width * height; ++i) {
        buffer[4 * i + 0] = 0;
        buffer[4 * i + 1] = 0;
        buffer[4 * i + 2] = 0;
        buffer[4 * i + 3] = 1;
    }
}

void Camera::init() {
    if (image_height == 0) image_height = (uint16_t)std::round((float)image_width / aspect_ratio);
    if (image_height < 1) image_height = 1;
    size = (uint32_t)image_height * (uint32_t)image_width;
    center = lookfrom;
    const float theta = vfov * kPi / 180.0f;
    const float h = std::tan(theta / 2.0f);
    const float viewport_height = 2 * h * focus_dist;
    const float viewport_width = viewport_height * ((float)image_width / (float)image_height);
    w = unitVector(lookfrom - lookat);
    u = unitVector(cross(vup, w));
    v = cross(w, u);
    const Vec3 viewport_u = splat3(viewport_width) * u;
    const Vec3 viewport_v = splat3(viewport_height) * -v;
    pixel_delta_u = viewport_u / splat3((float)image_width);
    pixel_delta_v = viewport_v / splat3((float)image_height);
    const Vec3 viewport_upper_left =
        center - splat3(focus_dist) * w - viewport_u / splat3(2.0f) - viewport_v / splat3(2.0f);
    pixel00_loc = viewport_upper_left + splat3(0.5f) * (pixel_delta_u + pixel_delta_v);
    const float defocus_radius = focus_dist * std::tan((defocus_angle / 2.0f) * kPi / 180.0f);
    defocus_disk_u = u * splat3(defocus_radius);
    defocus_disk_v = v * splat3(defocus_radius);
}

RtbCamera Camera::lowered() const {
    RtbCamera c{};
    c.image_width = image_width;
    c.image_height = image_height;
    c.samples_per_pixel = samples_per_pixel;
    c.max_depth = max_depth;
    put3(c.center, center);
    put3(c.pixel00_loc, pixel00_loc);
    put3(c.pixel_delta_u, pixel_delta_u);
    put3(c.pixel_delta_v, pixel_delta_v);
    put3(c.defocus_disk_u, defocus_disk_u);
    put3(c.defocus_disk_v, defocus_disk_v);
    c.defocus_angle = defocus_angle;
    put3(c.background, background);
    c.background_mode = background_mode;
    return c;
}

int Camera::render(RtbScene* scene, SharedStateImageWriter& writer, const RtbRenderOptions* options,
                   RtbRenderStats* stats) const {
    RtbRenderOptions opt{};
    if (options) opt = *options;
    const RtbCamera cam = lowered();
    return rtb_render(scene, &cam, &opt, writer.buffer.data(), writer.texture_buffer.data(), stats);
}

// ------------------------------------------------------------------ scenes (main.zig)
static World finishWorld(ObjectList objects, HostRng& bvh_rng, std::vector<Image> images) {
    World w;
    w.objects = std::move(objects);
    w.images = std::move(images);
    w.tree = BVHTree::init(w.objects, 0, w.objects.size(), bvh_rng);
    return w;
}

World generateWorld(HostRng& rng, HostRng& bvh_rng, const Book1Options& opt, std::vector<Image> images) {
    ObjectList objs;
    Material ground_material;
    if (opt.checker_ground) {  // main.zig:257-260
        const Texture checker = CheckerTexture::init(0.32f, SolidColor::init({0.2f, 0.3f, 0.1f}),
                                                     SolidColor::init({0.9f, 0.9f, 0.9f}));
        ground_material = Lambertian::init(checker);
    } else {
        ground_material = Lambertian::fromColor({0.5f, 0.5f, 0.5f});
    }
    objs.push_back(Sphere::init({0, -1000, 0}, 1000, ground_material));  // main.zig:262-263
    for (float a = -11; a < 11; a += 1) {
        for (float b = -11; b < 11; b += 1) {
            const float choose_mat = rng.randomDouble();
            Vec3 center;
            center.x = a + 0.9f * rng.randomDouble();
            center.y = 0.4f * choose_mat;
            center.z = b + 0.9f * rng.randomDouble();
            if (length(center - Vec3{4, 0.2f, 0}) > 0.9f) {
                if (choose_mat < 0.8f) {  // diffuse, main.zig:276-282
                    const Vec3 albedo = rng.random() * rng.random();
                    const Material m = Lambertian::fromColor(albedo);
                    Vec3 off;
                    off.x = rng.randomDoubleRange(0, 0.5f);
                    off.y = rng.randomDoubleRange(0, 0.5f);
                    off.z = rng.randomDoubleRange(0, 0.5f);
                    const Vec3 center2 = center + off;
                    if (opt.moving_spheres)
                        objs.push_back(Sphere::initMoving(center, center2, 0.4f * choose_mat, m));
                    else
                        objs.push_back(Sphere::init(center, 0.4f * choose_mat, m));
                } else if (choose_mat < 0.95f) {  // metal, main.zig:283-288
                    const Vec3 albedo = rng.randomRange(0.5f, 1);
                    const float fuzz = rng.randomDoubleRange(0, 0.5f);
                    objs.push_back(Sphere::init(center, 0.5f * choose_mat, Metal::fromColor(albedo, fuzz)));
                } else {  // glass, main.zig:289-293
                    objs.push_back(Sphere::init(center, 0.3f * choose_mat, Dielectric::init(rng.randomDoubleRange(1, 2))));
                }
            }
        }
    }
    objs.push_back(Sphere::init({0, 1, 0}, 1.0f, Dielectric::init(1.5f)));  // main.zig:297-298
    if (opt.earth_sphere && !images.empty())                                // main.zig:299-303
        objs.push_back(Sphere::init({-4, 1, 0}, 1.0f, Lambertian::init(ImageTexture::init(0))));
    else
        objs.push_back(Sphere::init({-4, 1, 0}, 1.0f, Lambertian::fromColor({0.4f, 0.2f, 0.1f})));
    objs.push_back(Sphere::init({4, 1, 0}, 1.0f, Metal::fromColor({0.7f, 0.6f, 0.5f}, 0.1f)));  // main.zig:305-306
    return finishWorld(std::move(objs), bvh_rng, std::move(images));
}

World earthWorld(HostRng& bvh_rng, std::vector<Image> images) {
    ObjectList objs;
    objs.push_back(Sphere::init({0, 0, 0}, 2, Lambertian::init(ImageTexture::init(0))));
    return finishWorld(std::move(objs), bvh_rng, std::move(images));
}

World twoSpheresWorld(HostRng& bvh_rng) {
    const Texture checker =
        CheckerTexture::init(0.8f, SolidColor::init({0.2f, 0.3f, 0.1f}), SolidColor::init({0.9f, 0.9f, 0.9f}));
    const Material m = Lambertian::init(checker);
    ObjectList objs;
    objs.push_back(Sphere::init({0, -10, 0}, 10, m));
    objs.push_back(Sphere::init({0, 10, 0}, 10, m));
    return finishWorld(std::move(objs), bvh_rng, {});
}

World twoPerlinWorld(HostRng& perlin_rng, HostRng& bvh_rng) {
    const Material m = Lambertian::init(NoiseTexture::init(4, perlin_rng));
    ObjectList objs;
    objs.push_back(Sphere::init({0, -1000, 0}, 1000, m));
    objs.push_back(Sphere::init({0, 2, 0}, 2, m));
    return finishWorld(std::move(objs), bvh_rng, {});
}

World cornellBox(HostRng& bvh_rng) {  // main.zig:168-205; the camera half is cornell_camera() on the caller's side
    const Material red = Lambertian::fromColor({0.65f, 0.05f, 0.05f});
    const Material white = Lambertian::fromColor({0.73f, 0.73f, 0.73f});
    const Material green = Lambertian::fromColor({0.12f, 0.45f, 0.15f});
    const Material light = DiffuseLight::fromColor({15, 15, 15});
    ObjectList objs;
    objs.push_back(Quad::init({555, 0, 0}, {0, 555, 0}, {0, 0, 555}, green));
    objs.push_back(Quad::init({0, 0, 0}, {0, 555, 0}, {0, 0, 555}, red));
    objs.push_back(Quad::init({343, 554, 332}, {-130, 0, 0}, {0, 0, -105}, light));
    objs.push_back(Quad::init({0, 0, 0}, {555, 0, 0}, {0, 0, 555}, white));
    objs.push_back(Quad::init({555, 555, 555}, {-555, 0, 0}, {0, 0, -555}, white));
    objs.push_back(Quad::init({0, 0, 555}, {555, 0, 0}, {0, 555, 0}, white));
    Hittable box1 = createBox({0, 0, 0}, {165, 330, 165}, white);
    box1 = RotateY::init(box1, 15);
    box1 = Translate::init(box1, {265, 0, 295});
    objs.push_back(box1);
    Hittable box2 = createBox({0, 0, 0}, {165, 165, 165}, white);
    box2 = RotateY::init(box2, -18);
    box2 = Translate::init(box2, {130, 0, 65});
    objs.push_back(box2);
    return finishWorld(std::move(objs), bvh_rng, {});
}

World cornellBoxSmoke(HostRng& bvh_rng) {  // main.zig:207-251
    const Material red = Lambertian::fromColor({0.65f, 0.05f, 0.05f});
    const Material white = Lambertian::fromColor({0.73f, 0.73f, 0.73f});
    const Material green = Lambertian::fromColor({0.12f, 0.45f, 0.15f});
    const Material light = DiffuseLight::fromColor({7, 7, 7});
    ObjectList objs;
    objs.push_back(Quad::init({555, 0, 0}, {0, 555, 0}, {0, 0, 555}, green));
    objs.push_back(Quad::init({0, 0, 0}, {0, 555, 0}, {0, 0, 555}, red));
    objs.push_back(Quad::init({113, 554, 127}, {330, 0, 0}, {0, 0, 305}, light));
    objs.push_back(Quad::init({0, 0, 0}, {555, 0, 0}, {0, 0, 555}, white));
    objs.push_back(Quad::init({555, 555, 555}, {-555, 0, 0}, {0, 0, -555}, white));
    objs.push_back(Quad::init({0, 0, 555}, {555, 0, 0}, {0, 555, 0}, white));
    Hittable box1 = createBox({0, 0, 0}, {165, 330, 165}, white);
    box1 = RotateY::init(box1, 15);
    box1 = Translate::init(box1, {265, 0, 295});
    objs.push_back(ConstantMedium::initFromColor(box1, 0.01f, {0, 0, 0}));
    Hittable box2 = createBox({0, 0, 0}, {165, 165, 165}, white);
    box2 = RotateY::init(box2, -18);
    box2 = Translate::init(box2, {130, 0, 65});
    objs.push_back(ConstantMedium::initFromColor(box2, 0.01f, {1, 1, 1}));
    return finishWorld(std::move(objs), bvh_rng, {});
}

World quadsWorld(HostRng& bvh_rng) {  // main.zig:127-143
    ObjectList objs;
    objs.push_back(Quad::init({-3, -2, 5}, {0, 0, -4}, {0, 4, 0}, Lambertian::fromColor({1, 0.2f, 0.2f})));
    objs.push_back(Quad::init({-2, -2, 0}, {4, 0, 0}, {0, 4, 0}, Lambertian::fromColor({0.2f, 1.0f, 0.2f})));
    objs.push_back(Quad::init({3, -2, 1}, {0, 0, 4}, {0, 4, 0}, Lambertian::fromColor({0.2f, 0.2f, 1.0f})));
    objs.push_back(Quad::init({-2, -3, 1}, {4, 0, 0}, {0, 0, 4}, Lambertian::fromColor({1.0f, 0.5f, 0})));
    objs.push_back(Quad::init({-2, -3, 5}, {4, 0, 0}, {0, 0, -4}, Lambertian::fromColor({0.2f, 0.8f, 0.8f})));
    return finishWorld(std::move(objs), bvh_rng, {});
}

World simpleLightWorld(HostRng& perlin_rng, HostRng& bvh_rng) {  // main.zig:145-166
    const Material material = Lambertian::init(NoiseTexture::init(4, perlin_rng));
    ObjectList objs;
    objs.push_back(Sphere::init({0, -1000, 0}, 1000, material));
    objs.push_back(Sphere::init({0, 2, 0}, 2, material));
    const Material difflight = DiffuseLight::fromColor({4, 4, 4});
    objs.push_back(Quad::init({3, 1, -2}, {2, 0, 0}, {0, 2, 0}, difflight));
    objs.push_back(Sphere::init({0, 7, 0}, 2, difflight));
    return finishWorld(std::move(objs), bvh_rng, {});
}

World texturedWorld(HostRng& perlin_rng, HostRng& bvh_rng, std::vector<Image> images) {
    const Material perlin = Lambertian::init(NoiseTexture::init(4, perlin_rng));
    const Texture checker =
        CheckerTexture::init(0.8f, SolidColor::init({0.2f, 0.3f, 0.1f}), SolidColor::init({0.9f, 0.9f, 0.9f}));
    ObjectList objs;
    objs.push_back(Sphere::init({0, -1000, 0}, 1000, perlin));               // twoPerlinWorld ground
    objs.push_back(Sphere::init({0, 2, 0}, 2, perlin));                      // twoPerlinWorld sphere
    if (!images.empty())
        objs.push_back(Sphere::init({-4.5f, 2, 0}, 2, Lambertian::init(ImageTexture::init(0))));  // earthWorld globe
    else
        objs.push_back(Sphere::init({-4.5f, 2, 0}, 2, Lambertian::fromColor({0.4f, 0.2f, 0.1f})));
    objs.push_back(Sphere::init({4.5f, 2, 0}, 2, Lambertian::init(checker)));  // twoSpheresWorld material
    objs.push_back(Sphere::init({2.2f, 0.7f, 3.0f}, 0.7f, Lambertian::init(checker)));
    objs.push_back(Sphere::init({-2.2f, 0.7f, 3.0f}, 0.7f, Dielectric::init(1.5f)));
    objs.push_back(Sphere::init({0.0f, 0.5f, 4.0f}, 0.5f, Metal::fromColor({0.8f, 0.8f, 0.9f}, 0.05f)));
    return finishWorld(std::move(objs), bvh_rng, std::move(images));
}

World randomSpheresWorld(HostRng& rng, HostRng& bvh_rng, uint32_t n) {
    ObjectList objs;
    objs.reserve((size_t)n + 1);
    objs.push_back(Sphere::init({0, -1000, 0}, 1000, Lambertian::fromColor({0.5f, 0.5f, 0.5f})));
    for (uint32_t i = 0; i < n; ++i) {
        const float choose_mat = rng.randomDouble();
        Vec3 center;
        center.x = rng.randomDoubleRange(-500, 500);
        center.y = rng.randomDoubleRange(0.2f, 40);
        center.z = rng.randomDoubleRange(-500, 500);
        const float radius = rng.randomDoubleRange(0.1f, 0.5f);
        if (choose_mat < 0.8f)
            objs.push_back(Sphere::init(center, radius, Lambertian::fromColor(rng.random() * rng.random())));
        else if (choose_mat < 0.95f)
            objs.push_back(Sphere::init(center, radius, Metal::fromColor(rng.randomRange(0.5f, 1), rng.randomDoubleRange(0, 0.5f))));
        else
            objs.push_back(Sphere::init(center, radius, Dielectric::init(rng.randomDoubleRange(1, 2))));
    }
    return finishWorld(std::move(objs), bvh_rng, {});
}

bool writePpm(const std::string& path, const uint8_t* rgba, uint32_t width, uint32_t height) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return false;
    std::fprintf(f, "P3\n%u %u\n255\n", width, height);  // stdout.zig:8
    for (uint32_t y = 0; y < height; ++y) {
        for (uint32_t x = 0; x < width; ++x) {
            const uint8_t* p = rgba + 4 * ((size_t)y * width + x);
            std::fprintf(f, "%u %u %u\n", p[0], p[1], p[2]);
        }
    }
    return std::fclose(f) == 0;
}

// 8-bit RGB PNG from the RGBA8 texture buffer ("save to file" is an open TODO of the reference, src/main.zig:47).
// Self-contained: stored (uncompressed) deflate blocks inside a zlib stream, CRC-32 per chunk, Adler-32 of the raw
// scanlines — every PNG reader accepts it, and the file is byte-for-byte reproducible.
namespace {
uint32_t crc32Update(uint32_t crc, const uint8_t* data, size_t n) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ data[i]) & 0xffu] ^ (crc >> 8);
    return crc;
}
void putBe32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24));
    v.push_back((uint8_t)(x >> 16));
    v.push_back((uint8_t)(x >> 8));
    v.push_back((uint8_t)x);
}
bool writeChunk(FILE* f, const char type[4], const std::vector<uint8_t>& data) {
    std::vector<uint8_t> head;
    putBe32(head, (uint32_t)data.size());
    head.insert(head.end(), type, type + 4);
    uint32_t crc = crc32Update(0xffffffffu, head.data() + 4, 4);
    crc = crc32Update(crc, data.data(), data.size()) ^ 0xffffffffu;
    std::vector<uint8_t> tail;
    putBe32(tail, crc);
    return std::fwrite(head.data(), 1, head.size(), f) == head.size() &&
           (data.empty() || std::fwrite(data.data(), 1, data.size(), f) == data.size()) &&
           std::fwrite(tail.data(), 1, 4, f) == 4;
}
}  // namespace

bool writePng(const std::string& path, const uint8_t* rgba, uint32_t width, uint32_t height) {
    if (width == 0 || height == 0) return false;
    // raw image data: per scanline a filter byte (0 = none) + RGB triples
    std::vector<uint8_t> raw;
    raw.reserve((size_t)height * (1 + 3 * (size_t)width));
    for (uint32_t y = 0; y < height; ++y) {
        raw.push_back(0);
        for (uint32_t x = 0; x < width; ++x) {
            const uint8_t* p = rgba + 4 * ((size_t)y * width + x);
            raw.insert(raw.end(), p, p + 3);
        }
    }
    std::vector<uint8_t> z;
    z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
    z.push_back(0x78);  // zlib header: deflate, 32 K window, no preset dictionary, check bits
    z.push_back(0x01);
    uint32_t a = 1, b = 0;  // Adler-32
    for (size_t off = 0; off < raw.size();) {
        const size_t n = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + n == raw.size() ? 1 : 0);  // BFINAL, BTYPE = 00 (stored)
        z.push_back((uint8_t)(n & 0xff));
        z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)(~n & 0xff));
        z.push_back((uint8_t)((~n >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + (ptrdiff_t)off, raw.begin() + (ptrdiff_t)(off + n));
        for (size_t i = off; i < off + n; ++i) {
            a = (a + raw[i]) % 65521u;
            b = (b + a) % 65521u;
        }
        off += n;
    }
    putBe32(z, (b << 16) | a);
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    static const uint8_t signature[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    bool ok = std::fwrite(signature, 1, 8, f) == 8;
    std::vector<uint8_t> ihdr;
    putBe32(ihdr, width);
    putBe32(ihdr, height);
    const uint8_t rest[5] = {8, 2, 0, 0, 0};  // bit depth 8, colour type 2 (RGB), deflate, adaptive filtering, no interlace
    ihdr.insert(ihdr.end(), rest, rest + 5);
    ok = ok && writeChunk(f, "IHDR", ihdr) && writeChunk(f, "IDAT", z) && writeChunk(f, "IEND", {});
    return (std::fclose(f) == 0) && ok;
}

}  // namespace rtw
