/*
 * rtw_host.h — C view of the host-side mirror of the reference's scene/camera API
 * (zig-raytracing-weekend_b200/host/rtw_host.hpp), for drivers that are not C++ (the Python
 * tests and bench.py use it through ctypes).  It builds worlds the way src/main.zig does
 * (scene builders, BVHTree.init), runs Camera.init (src/camera.zig:118-154) and lowers the result
 * into the RtbSceneDesc / RtbCamera PODs of rtb.h.  No per-pixel work happens here.
 */
#ifndef RTW_HOST_H
#define RTW_HOST_H

#include "rtb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RtwWorld RtwWorld; /* a built world (object list + BVH) and its lowered form */

/* Scene builders of src/main.zig plus the two synthetic BASELINE scenes. */
enum {
    RTW_SCENE_BOOK1 = 0,         /* generateWorld       src/main.zig:253-312 */
    RTW_SCENE_EARTH = 1,         /* earthWorld          src/main.zig:88-99   */
    RTW_SCENE_TWO_SPHERES = 2,   /* twoSpheresWorld     src/main.zig:101-113 */
    RTW_SCENE_TWO_PERLIN = 3,    /* twoPerlinWorld      src/main.zig:115-125 */
    RTW_SCENE_TEXTURED = 4,      /* BASELINE config 3: checker + earth + perlin in one world */
    RTW_SCENE_RANDOM_SPHERES = 5, /* BASELINE config 4: n random spheres + ground            */
    RTW_SCENE_QUADS = 6,          /* quadsWorld          src/main.zig:127-143 */
    RTW_SCENE_SIMPLE_LIGHT = 7,   /* simpleLightWorld    src/main.zig:145-166 (camera: lookfrom (26,3,6), lookat
                                     (0,2,0), depth 50, defocus 0, black background) */
    RTW_SCENE_CORNELL_BOX = 8,    /* cornellBox          src/main.zig:168-205 — HEAD's selected scene (:422);
                                     camera: 600 x 600, 200 spp, depth 200, vfov 40, lookfrom (278,278,-800),
                                     lookat (278,278,0), defocus 0, black background */
    RTW_SCENE_CORNELL_SMOKE = 9   /* cornellBoxSmoke     src/main.zig:207-251 (same camera, depth 50) */
};
enum {
    RTW_BOOK1_CHECKER_GROUND = 1u << 0, /* HEAD's checker ground (main.zig:257-260) instead of Book-1 grey */
    RTW_BOOK1_EARTH_SPHERE = 1u << 1,   /* HEAD's earthmap sphere (main.zig:299-303); needs an image       */
    RTW_BOOK1_STATIC_SPHERES = 1u << 2  /* Sphere.init instead of initMoving for the diffuse spheres        */
};

typedef struct RtwSceneParams {
    uint32_t kind;  /* RTW_SCENE_* */
    uint32_t flags; /* RTW_BOOK1_* */
    uint64_t scene_seed;
    uint64_t bvh_seed;
    uint64_t perlin_seed;
    uint32_t n_spheres;        /* RTW_SCENE_RANDOM_SPHERES */
    uint32_t image_width;      /* optional RGBA8 image 0 (earthmap), tightly packed */
    uint32_t image_height;
    uint32_t reserved;
    const uint8_t* image_rgba;
} RtwSceneParams;

int rtw_world_create(const RtwSceneParams* params, RtwWorld** world_out);

/* Incremental builder: the calls a Zig scene function makes (Sphere.init / initMoving / Quad.init
 * appended to world_objects, then BVHTree.init).  Material/texture spec in one POD. */
typedef struct RtwMaterialSpec {
    uint32_t material; /* RTB_MAT_*  */
    uint32_t texture;  /* RTB_TEX_* for lambertian / diffuse_light / isotropic */
    float color[3];    /* metal albedo | solid colour | checker even */
    float color2[3];   /* checker odd */
    float scale;       /* checker scale (NOT inverted) | noise scale */
    float fuzz;
    float ir;
    uint32_t image_index;
    uint64_t perlin_seed; /* noise: tables are generated from this seed (same seed = shared tables) */
} RtwMaterialSpec;

int rtw_world_new(RtwWorld** world_out);
int rtw_world_add_image(RtwWorld* world, const uint8_t* rgba, uint32_t width, uint32_t height);
int rtw_world_add_sphere(RtwWorld* world, const float center1[3], const float* center2_or_null, float radius,
                         const RtwMaterialSpec* material);
int rtw_world_add_quad(RtwWorld* world, const float q[3], const float u[3], const float v[3],
                       const RtwMaterialSpec* material);
/* Translate.init(RotateY.init(createBox(a, b, material), angle), offset) as one world object (src/main.zig:182-190);
 * rotate = 0 skips RotateY, offset = NULL skips Translate (src/objects.zig:314-319, :354-397, :510-532). */
int rtw_world_add_box(RtwWorld* world, const float a[3], const float b[3], int rotate, float angle_degrees,
                      const float* offset_or_null, const RtwMaterialSpec* material);
/* ConstantMedium.initFromColor(&box, density, color) with the same kind of box as the boundary
 * (src/objects.zig:450-452; cornellBoxSmoke src/main.zig:223-236). */
int rtw_world_add_medium(RtwWorld* world, const float a[3], const float b[3], int rotate, float angle_degrees,
                         const float* offset_or_null, float density, const float color[3]);
/* General instancing (src/objects.zig:264-443): objects are first built DETACHED — every call returns a handle —,
 * wrapped as often as wanted in any order (the wrapper copies the wrapped object, as the reference copies `obj` into
 * allocator.create(Hittable)), and finally attached to the world list with rtw_world_add_object. */
int rtw_obj_sphere(RtwWorld* world, const float center1[3], const float* center2_or_null, float radius,
                   const RtwMaterialSpec* material, uint32_t* handle_out);
int rtw_obj_quad(RtwWorld* world, const float q[3], const float u[3], const float v[3], const RtwMaterialSpec* material,
                 uint32_t* handle_out);
int rtw_obj_box(RtwWorld* world, const float a[3], const float b[3], const RtwMaterialSpec* material,
                uint32_t* handle_out);                                                  /* createBox */
int rtw_obj_list(RtwWorld* world, const uint32_t* handles, uint32_t n, uint32_t* handle_out); /* HittableList, in order */
int rtw_obj_translate(RtwWorld* world, uint32_t handle, const float offset[3], uint32_t* handle_out);
int rtw_obj_rotate_y(RtwWorld* world, uint32_t handle, float angle_degrees, uint32_t* handle_out);
int rtw_obj_medium(RtwWorld* world, uint32_t boundary, float density, const float color[3], uint32_t* handle_out);
int rtw_world_add_object(RtwWorld* world, uint32_t handle);
int rtw_world_build(RtwWorld* world, uint64_t bvh_seed); /* BVHTree.init + lowering */

const RtbSceneDesc* rtw_world_desc(const RtwWorld* world); /* valid until rtw_world_destroy */
/* Bounding box {min xyz, max xyz} the host computed for object i (after the BVH permutation). */
int rtw_world_object_box(const RtwWorld* world, uint32_t index, float box6[6]);
void rtw_world_destroy(RtwWorld* world);

/* Camera options = the public fields of Camera (src/camera.zig:70-91). */
typedef struct RtwCameraOptions {
    float aspect_ratio;
    uint32_t image_width;
    uint32_t image_height; /* 0 = derive from aspect_ratio (camera.zig:119-120) */
    uint32_t samples_per_pixel;
    uint32_t max_depth;
    float background[3];
    float vfov;
    float lookfrom[3];
    float lookat[3];
    float vup[3];
    float defocus_angle;
    float focus_dist;
    uint32_t background_mode; /* RTB_BACKGROUND_* */
} RtwCameraOptions;

void rtw_camera_defaults(RtwCameraOptions* options);                        /* camera.zig:70-91 defaults */
int rtw_camera_init(const RtwCameraOptions* options, RtbCamera* camera_out); /* Camera.init */

/* Camera.render drop-in: SharedStateImageWriter.init-style buffers (buffer = (0,0,0,1) per pixel
 * when `scrub` is non-zero, src/camera.zig:29-45), then one rtb_render call. */
int rtw_camera_render(RtbScene* scene, const RtwCameraOptions* options, const RtbRenderOptions* render_options,
                      int scrub, float* buffer, uint8_t* texture_buffer, RtbRenderStats* stats);

/* Output formats ("save to file" is an open TODO of the reference, src/main.zig:47):
 * P3 PPM as src/stdout.zig:5-18 prints it (values come from the RGBA8 buffer, so never 256), and an 8-bit RGB PNG. */
int rtw_write_ppm(const char* path, const uint8_t* rgba, uint32_t width, uint32_t height);
int rtw_write_png(const char* path, const uint8_t* rgba, uint32_t width, uint32_t height);

#ifdef __cplusplus
}
#endif
#endif
