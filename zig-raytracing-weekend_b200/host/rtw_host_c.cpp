// rtw_host_c.cpp — C wrappers (include/rtw_host.h) over the C++ host mirror.
#include <cstring>
#include <map>
#include <memory>
#include <new>

#include "../../include/rtw_host.h"
#include "rtw_host.hpp"

using namespace rtw;

struct RtwWorld {
    World world;
    std::unique_ptr<LoweredScene> lowered;
    std::map<uint64_t, Texture> noise_by_seed;
    std::vector<Hittable> detached;  // objects built by rtw_obj_* and not (yet) attached to the world list
    bool built = false;
};

static Vec3 V(const float* p) { return Vec3{p[0], p[1], p[2]}; }

static Image make_image(const uint8_t* rgba, uint32_t w, uint32_t h) {
    Image im;
    im.width = w;
    im.height = h;
    im.bytes_per_row = w * 4;  // zstbi: bytes_per_row = width * components (zstbi.zig:138)
    im.data.assign(rgba, rgba + (size_t)w * h * 4);
    return im;
}

extern "C" int rtw_world_create(const RtwSceneParams* p, RtwWorld** out) {
    if (!p || !out) return RTB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    std::vector<Image> images;
    if (p->image_rgba && p->image_width && p->image_height)
        images.push_back(make_image(p->image_rgba, p->image_width, p->image_height));
    HostRng scene_rng(p->scene_seed), bvh_rng(p->bvh_seed), perlin_rng(p->perlin_seed);
    auto w = std::unique_ptr<RtwWorld>(new (std::nothrow) RtwWorld());
    if (!w) return RTB_ERR_OUT_OF_MEMORY;
    switch (p->kind) {
        case RTW_SCENE_BOOK1: {
            Book1Options o;
            o.checker_ground = (p->flags & RTW_BOOK1_CHECKER_GROUND) != 0;
            o.earth_sphere = (p->flags & RTW_BOOK1_EARTH_SPHERE) != 0;
            o.moving_spheres = (p->flags & RTW_BOOK1_STATIC_SPHERES) == 0;
            w->world = generateWorld(scene_rng, bvh_rng, o, std::move(images));
            break;
        }
        case RTW_SCENE_EARTH:
            if (images.empty()) return RTB_ERR_INVALID_ARGUMENT;
            w->world = earthWorld(bvh_rng, std::move(images));
            break;
        case RTW_SCENE_TWO_SPHERES:
            w->world = twoSpheresWorld(bvh_rng);
            break;
        case RTW_SCENE_TWO_PERLIN:
            w->world = twoPerlinWorld(perlin_rng, bvh_rng);
            break;
        case RTW_SCENE_TEXTURED:
            w->world = texturedWorld(perlin_rng, bvh_rng, std::move(images));
            break;
        case RTW_SCENE_RANDOM_SPHERES:
            w->world = randomSpheresWorld(scene_rng, bvh_rng, p->n_spheres);
            break;
        case RTW_SCENE_QUADS:
            w->world = quadsWorld(bvh_rng);
            break;
        case RTW_SCENE_CORNELL_BOX:
            w->world = cornellBox(bvh_rng);
            break;
        case RTW_SCENE_CORNELL_SMOKE:
            w->world = cornellBoxSmoke(bvh_rng);
            break;
        case RTW_SCENE_SIMPLE_LIGHT:
            w->world = simpleLightWorld(perlin_rng, bvh_rng);
            break;
        default:
            return RTB_ERR_INVALID_ARGUMENT;
    }
    w->lowered = w->world.lower();
    w->built = true;
    *out = w.release();
    return RTB_OK;
}

extern "C" int rtw_world_new(RtwWorld** out) {
    if (!out) return RTB_ERR_INVALID_ARGUMENT;
    *out = new (std::nothrow) RtwWorld();
    return *out ? RTB_OK : RTB_ERR_OUT_OF_MEMORY;
}

extern "C" int rtw_world_add_image(RtwWorld* w, const uint8_t* rgba, uint32_t width, uint32_t height) {
    if (!w || w->built || !rgba || !width || !height) return RTB_ERR_INVALID_ARGUMENT;
    w->world.images.push_back(make_image(rgba, width, height));
    return RTB_OK;
}

static int make_material(RtwWorld* w, const RtwMaterialSpec* s, Material* out) {
    if (!s) return RTB_ERR_INVALID_ARGUMENT;
    Texture tex;
    switch (s->texture) {
        case RTB_TEX_SOLID:
            tex = SolidColor::init(V(s->color));
            break;
        case RTB_TEX_CHECKER:
            tex = CheckerTexture::init(s->scale, SolidColor::init(V(s->color)), SolidColor::init(V(s->color2)));
            break;
        case RTB_TEX_IMAGE:
            if (s->image_index >= w->world.images.size()) return RTB_ERR_INVALID_ARGUMENT;
            tex = ImageTexture::init(s->image_index);
            break;
        case RTB_TEX_NOISE: {
            auto it = w->noise_by_seed.find(s->perlin_seed);
            if (it == w->noise_by_seed.end()) {
                HostRng rng(s->perlin_seed);
                it = w->noise_by_seed.emplace(s->perlin_seed, NoiseTexture::init(1.0f, rng)).first;
            }
            tex = it->second;
            tex.scale = s->scale;
            break;
        }
        default:
            return RTB_ERR_INVALID_ARGUMENT;
    }
    switch (s->material) {
        case RTB_MAT_LAMBERTIAN:
            *out = Lambertian::init(tex);
            break;
        case RTB_MAT_METAL:
            *out = Metal::fromColor(V(s->color), s->fuzz);
            break;
        case RTB_MAT_DIELECTRIC:
            *out = Dielectric::init(s->ir);
            break;
        case RTB_MAT_DIFFUSE_LIGHT:
            *out = DiffuseLight::init(tex);
            break;
        default:
            return RTB_ERR_UNSUPPORTED;
    }
    return RTB_OK;
}

extern "C" int rtw_world_add_sphere(RtwWorld* w, const float c1[3], const float* c2, float radius,
                                    const RtwMaterialSpec* spec) {
    if (!w || w->built || !c1) return RTB_ERR_INVALID_ARGUMENT;
    Material m;
    const int rc = make_material(w, spec, &m);
    if (rc != RTB_OK) return rc;
    w->world.objects.push_back(c2 ? Sphere::initMoving(V(c1), V(c2), radius, m) : Sphere::init(V(c1), radius, m));
    return RTB_OK;
}

extern "C" int rtw_world_add_quad(RtwWorld* w, const float q[3], const float u[3], const float v[3],
                                  const RtwMaterialSpec* spec) {
    if (!w || w->built || !q || !u || !v) return RTB_ERR_INVALID_ARGUMENT;
    Material m;
    const int rc = make_material(w, spec, &m);
    if (rc != RTB_OK) return rc;
    w->world.objects.push_back(Quad::init(V(q), V(u), V(v), m));
    return RTB_OK;
}

extern "C" int rtw_world_add_box(RtwWorld* w, const float a[3], const float b[3], int rotate, float angle_degrees,
                                 const float* offset_or_null, const RtwMaterialSpec* spec) {
    if (!w || w->built || !a || !b) return RTB_ERR_INVALID_ARGUMENT;
    Material m;
    const int rc = make_material(w, spec, &m);
    if (rc != RTB_OK) return rc;
    Hittable box = createBox(V(a), V(b), m);
    if (rotate) box = RotateY::init(box, angle_degrees);
    if (offset_or_null) box = Translate::init(box, V(offset_or_null));
    w->world.objects.push_back(box);
    return RTB_OK;
}

extern "C" int rtw_world_add_medium(RtwWorld* w, const float a[3], const float b[3], int rotate, float angle_degrees,
                                    const float* offset_or_null, float density, const float color[3]) {
    if (!w || w->built || !a || !b || !color || !(density > 0)) return RTB_ERR_INVALID_ARGUMENT;
    Hittable box = createBox(V(a), V(b), Lambertian::fromColor({0.73f, 0.73f, 0.73f}));  // boundary material is unused
    if (rotate) box = RotateY::init(box, angle_degrees);
    if (offset_or_null) box = Translate::init(box, V(offset_or_null));
    w->world.objects.push_back(ConstantMedium::initFromColor(box, density, V(color)));
    return RTB_OK;
}

// ---- detached objects + general wrappers ------------------------------------------------------------------------
static int detach(RtwWorld* w, const Hittable& h, uint32_t* handle_out) {
    if (!handle_out) return RTB_ERR_INVALID_ARGUMENT;
    w->detached.push_back(h);
    *handle_out = (uint32_t)w->detached.size() - 1u;
    return RTB_OK;
}
extern "C" int rtw_obj_sphere(RtwWorld* w, const float c1[3], const float* c2, float radius, const RtwMaterialSpec* spec,
                              uint32_t* handle_out) {
    if (!w || w->built || !c1) return RTB_ERR_INVALID_ARGUMENT;
    Material m;
    const int rc = make_material(w, spec, &m);
    if (rc != RTB_OK) return rc;
    return detach(w, c2 ? Sphere::initMoving(V(c1), V(c2), radius, m) : Sphere::init(V(c1), radius, m), handle_out);
}
extern "C" int rtw_obj_quad(RtwWorld* w, const float q[3], const float u[3], const float v[3], const RtwMaterialSpec* spec,
                            uint32_t* handle_out) {
    if (!w || w->built || !q || !u || !v) return RTB_ERR_INVALID_ARGUMENT;
    Material m;
    const int rc = make_material(w, spec, &m);
    if (rc != RTB_OK) return rc;
    return detach(w, Quad::init(V(q), V(u), V(v), m), handle_out);
}
extern "C" int rtw_obj_box(RtwWorld* w, const float a[3], const float b[3], const RtwMaterialSpec* spec, uint32_t* handle_out) {
    if (!w || w->built || !a || !b) return RTB_ERR_INVALID_ARGUMENT;
    Material m;
    const int rc = make_material(w, spec, &m);
    if (rc != RTB_OK) return rc;
    return detach(w, createBox(V(a), V(b), m), handle_out);
}
extern "C" int rtw_obj_list(RtwWorld* w, const uint32_t* handles, uint32_t n, uint32_t* handle_out) {
    if (!w || w->built || !handles || n == 0) return RTB_ERR_INVALID_ARGUMENT;
    std::vector<Hittable> members;
    for (uint32_t k = 0; k < n; ++k) {
        if (handles[k] >= w->detached.size()) return RTB_ERR_INVALID_ARGUMENT;
        members.push_back(w->detached[handles[k]]);
    }
    return detach(w, HittableList::init(members), handle_out);
}
extern "C" int rtw_obj_translate(RtwWorld* w, uint32_t handle, const float offset[3], uint32_t* handle_out) {
    if (!w || w->built || !offset || handle >= w->detached.size()) return RTB_ERR_INVALID_ARGUMENT;
    return detach(w, TranslateAny::init(w->detached[handle], V(offset)), handle_out);
}
extern "C" int rtw_obj_rotate_y(RtwWorld* w, uint32_t handle, float angle_degrees, uint32_t* handle_out) {
    if (!w || w->built || handle >= w->detached.size()) return RTB_ERR_INVALID_ARGUMENT;
    return detach(w, RotateYAny::init(w->detached[handle], angle_degrees), handle_out);
}
extern "C" int rtw_obj_medium(RtwWorld* w, uint32_t boundary, float density, const float color[3], uint32_t* handle_out) {
    if (!w || w->built || !color || !(density > 0) || boundary >= w->detached.size()) return RTB_ERR_INVALID_ARGUMENT;
    return detach(w, ConstantMediumOf::initFromColor(w->detached[boundary], density, V(color)), handle_out);
}
extern "C" int rtw_world_add_object(RtwWorld* w, uint32_t handle) {
    if (!w || w->built || handle >= w->detached.size()) return RTB_ERR_INVALID_ARGUMENT;
    w->world.objects.push_back(w->detached[handle]);
    return RTB_OK;
}

extern "C" int rtw_world_build(RtwWorld* w, uint64_t bvh_seed) {
    if (!w || w->built) return RTB_ERR_INVALID_ARGUMENT;
    HostRng rng(bvh_seed);
    w->world.tree = BVHTree::init(w->world.objects, 0, w->world.objects.size(), rng);
    w->lowered = w->world.lower();
    w->built = true;
    return RTB_OK;
}

extern "C" const RtbSceneDesc* rtw_world_desc(const RtwWorld* w) {
    return (w && w->built) ? &w->lowered->desc : nullptr;
}

extern "C" int rtw_world_object_box(const RtwWorld* w, uint32_t index, float box6[6]) {
    if (!w || !box6 || index >= w->world.objects.size()) return RTB_ERR_INVALID_ARGUMENT;
    const Aabb& b = w->world.objects[index].bounding_box;
    box6[0] = b.x.min;
    box6[1] = b.y.min;
    box6[2] = b.z.min;
    box6[3] = b.x.max;
    box6[4] = b.y.max;
    box6[5] = b.z.max;
    return RTB_OK;
}

extern "C" void rtw_world_destroy(RtwWorld* w) { delete w; }

static Camera to_camera(const RtwCameraOptions* o) {
    Camera c;
    c.aspect_ratio = o->aspect_ratio;
    c.image_width = (uint16_t)o->image_width;
    c.image_height = (uint16_t)o->image_height;
    c.samples_per_pixel = (uint16_t)o->samples_per_pixel;
    c.max_depth = (uint8_t)o->max_depth;
    c.background = V(o->background);
    c.vfov = o->vfov;
    c.lookfrom = V(o->lookfrom);
    c.lookat = V(o->lookat);
    c.vup = V(o->vup);
    c.defocus_angle = o->defocus_angle;
    c.focus_dist = o->focus_dist;
    c.background_mode = o->background_mode;
    return c;
}

extern "C" void rtw_camera_defaults(RtwCameraOptions* o) {
    if (!o) return;
    const Camera c;
    std::memset(o, 0, sizeof(*o));
    o->aspect_ratio = c.aspect_ratio;
    o->image_width = c.image_width;
    o->image_height = c.image_height;
    o->samples_per_pixel = c.samples_per_pixel;
    o->max_depth = c.max_depth;
    o->vfov = c.vfov;
    o->lookfrom[0] = c.lookfrom.x;
    o->lookfrom[1] = c.lookfrom.y;
    o->lookfrom[2] = c.lookfrom.z;
    o->vup[1] = 1.0f;
    o->defocus_angle = c.defocus_angle;
    o->focus_dist = c.focus_dist;
    o->background_mode = RTB_BACKGROUND_SOLID;
}

// The reference's fields are u16 / u16 / u16 / u8 (src/camera.zig:71-79).
static bool camera_in_range(const RtwCameraOptions* o) {
    return o && o->image_width >= 1 && o->image_width <= 65535u && o->image_height <= 65535u &&
           o->samples_per_pixel <= 65535u && o->max_depth <= 255u && o->aspect_ratio > 0;
}

extern "C" int rtw_camera_init(const RtwCameraOptions* o, RtbCamera* out) {
    if (!camera_in_range(o) || !out) return RTB_ERR_INVALID_ARGUMENT;
    Camera c = to_camera(o);
    c.init();
    *out = c.lowered();
    return RTB_OK;
}

extern "C" int rtw_camera_render(RtbScene* scene, const RtwCameraOptions* o, const RtbRenderOptions* ro, int scrub,
                                 float* buffer, uint8_t* texture_buffer, RtbRenderStats* stats) {
    if (!camera_in_range(o) || !buffer) return RTB_ERR_INVALID_ARGUMENT;
    Camera c = to_camera(o);
    c.init();
    if (scrub) {  // SharedStateImageWriter.scrub, src/camera.zig:41-45
        for (size_t i = 0; i < (size_t)c.size; ++i) {
            buffer[4 * i + 0] = buffer[4 * i + 1] = buffer[4 * i + 2] = 0.0f;
            buffer[4 * i + 3] = 1.0f;
        }
    }
    RtbRenderOptions opt{};
    if (ro) opt = *ro;
    const RtbCamera cam = c.lowered();
    return rtb_render(scene, &cam, &opt, buffer, texture_buffer, stats);
}

extern "C" int rtw_write_ppm(const char* path, const uint8_t* rgba, uint32_t width, uint32_t height) {
    if (!path || !rgba) return RTB_ERR_INVALID_ARGUMENT;
    return writePpm(path, rgba, width, height) ? RTB_OK : RTB_ERR_INVALID_ARGUMENT;
}

extern "C" int rtw_write_png(const char* path, const uint8_t* rgba, uint32_t width, uint32_t height) {
    if (!path || !rgba) return RTB_ERR_INVALID_ARGUMENT;
    return writePng(path, rgba, width, height) ? RTB_OK : RTB_ERR_INVALID_ARGUMENT;
}
