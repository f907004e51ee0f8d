/* abi_smoke.c — include/rtb.h used from plain C99, outside ctypes: build a two-sphere scene by hand (the POD arrays a
 * Zig / C host would fill), create -> trace -> render -> group render -> destroy, and check the obvious.
 *
 *   gcc -std=c99 -pedantic -Wall -Werror -I include tests/abi_smoke.c -L zig-raytracing-weekend_b200/_lib -lrtb \
 *       -Wl,-rpath,zig-raytracing-weekend_b200/_lib -lm -o abi_smoke
 *
 * Exit code 0 = ok, 77 = no CUDA device (the library has no CPU fallback and says so), anything else = failure.
 * tests/test_host_and_abi.py compiles it (CPU suite) and runs it (GPU suite). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rtb.h"

#define CHECK(call)                                                                         \
    do {                                                                                    \
        int rc__ = (call);                                                                  \
        if (rc__ != RTB_OK) {                                                               \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, rtb_last_error());               \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

int main(void) {
    int n_dev = 0;
    RtbHittable hittables[2];
    RtbMaterial materials[2];
    RtbTexture textures[1];
    RtbBvhNode nodes[3];
    RtbSceneDesc desc;
    RtbCamera cam;
    RtbRenderOptions opt;
    RtbRenderStats stats;
    RtbScene* scene = NULL;
    RtbSceneGroup* group = NULL;
    RtbRay ray;
    RtbHit hit;
    float *accum, *accum2;
    unsigned char* rgba;
    const unsigned W = 64, H = 36;
    unsigned i;
    int devices[2] = {0, 0};
    double sum = 0.0;

    if (rtb_abi_version() != RTB_ABI_VERSION) return 2;
    if (sizeof(RtbHittable) != 64 || sizeof(RtbMaterial) != 32 || sizeof(RtbTexture) != 48 || sizeof(RtbBvhNode) != 40 ||
        sizeof(RtbRay) != 36 || sizeof(RtbHit) != 52)
        return 3;
    if (rtb_device_count(&n_dev) != RTB_OK) {
        /* the error path must be loud and must not crash */
        if (strstr(rtb_last_error(), "no CPU fallback") == NULL) return 4;
        memset(&desc, 0, sizeof(desc));
        desc.abi_version = RTB_ABI_VERSION;
        if (rtb_scene_create(&desc, 0, &scene) != RTB_ERR_NO_DEVICE || scene != NULL) return 5;
        printf("abi_smoke: no CUDA device (%s)\n", rtb_last_error());
        return 77;
    }

    /* world: a ground sphere (lambertian, solid grey) and a metal ball; BVH = root + two leaves */
    memset(hittables, 0, sizeof(hittables));
    memset(materials, 0, sizeof(materials));
    memset(textures, 0, sizeof(textures));
    memset(nodes, 0, sizeof(nodes));
    textures[0].type = RTB_TEX_SOLID;
    textures[0].color[0] = textures[0].color[1] = textures[0].color[2] = 0.5f;
    materials[0].type = RTB_MAT_LAMBERTIAN;
    materials[0].texture = 0;
    materials[1].type = RTB_MAT_METAL;
    materials[1].albedo[0] = 0.8f; materials[1].albedo[1] = 0.6f; materials[1].albedo[2] = 0.2f;
    materials[1].fuzz = 0.1f;
    hittables[0].type = RTB_HITTABLE_SPHERE; hittables[0].material = 0; hittables[0].radius = 100.0f;
    hittables[0].a[1] = -100.5f; hittables[0].a[2] = -1.0f;
    hittables[1].type = RTB_HITTABLE_SPHERE; hittables[1].material = 1; hittables[1].radius = 0.5f;
    hittables[1].a[2] = -1.0f;
    for (i = 0; i < 2; ++i) {
        int k;
        for (k = 0; k < 3; ++k) {
            nodes[1 + i].bmin[k] = hittables[i].a[k] - hittables[i].radius;
            nodes[1 + i].bmax[k] = hittables[i].a[k] + hittables[i].radius;
        }
        nodes[1 + i].left = nodes[1 + i].right = -1;
        nodes[1 + i].leaf = (int)i;
    }
    for (i = 0; i < 3; ++i) {
        nodes[0].bmin[i] = fminf(nodes[1].bmin[i], nodes[2].bmin[i]);
        nodes[0].bmax[i] = fmaxf(nodes[1].bmax[i], nodes[2].bmax[i]);
    }
    nodes[0].left = 1; nodes[0].right = 2; nodes[0].leaf = -1;
    memset(&desc, 0, sizeof(desc));
    desc.abi_version = RTB_ABI_VERSION;
    desc.n_nodes = 3; desc.n_hittables = 2; desc.n_materials = 2; desc.n_textures = 1; desc.root = 0;
    desc.nodes = nodes; desc.hittables = hittables; desc.materials = materials; desc.textures = textures;
    CHECK(rtb_scene_create(&desc, 0, &scene));

    /* one ray straight at the ball: t = 0.5, object 1, front face */
    memset(&ray, 0, sizeof(ray));
    ray.direction[2] = -1.0f; ray.t_min = 0.001f; ray.t_max = INFINITY;
    CHECK(rtb_trace_rays(scene, &ray, 1, RTB_TRAVERSAL_REFERENCE, &hit));
    if (hit.object != 1 || hit.front_face != 1u || fabsf(hit.t - 0.5f) > 1e-6f) return 6;
    CHECK(rtb_trace_rays(scene, &ray, 1, RTB_TRAVERSAL_SAH16, &hit));
    if (hit.object != 1 || fabsf(hit.t - 0.5f) > 1e-6f) return 7;

    /* camera as Camera.init derives it for lookfrom (0,0,0), lookat (0,0,-1), vfov 90, no defocus */
    memset(&cam, 0, sizeof(cam));
    cam.image_width = W; cam.image_height = H; cam.samples_per_pixel = 4; cam.max_depth = 10;
    cam.pixel_delta_u[0] = 2.0f * ((float)W / (float)H) / (float)W;
    cam.pixel_delta_v[1] = -2.0f / (float)H;
    cam.pixel00_loc[0] = -((float)W / (float)H) + 0.5f * cam.pixel_delta_u[0];
    cam.pixel00_loc[1] = 1.0f + 0.5f * cam.pixel_delta_v[1];
    cam.pixel00_loc[2] = -1.0f;
    cam.background_mode = RTB_BACKGROUND_SKY;
    memset(&opt, 0, sizeof(opt));
    opt.seed = 1234; opt.integrator = RTB_INTEGRATOR_WAVEFRONT; opt.traversal = RTB_TRAVERSAL_SAH16;
    accum = (float*)calloc((size_t)W * H * 4, sizeof(float));
    accum2 = (float*)calloc((size_t)W * H * 4, sizeof(float));
    rgba = (unsigned char*)calloc((size_t)W * H * 4, 1);
    if (!accum || !accum2 || !rgba) return 8;
    CHECK(rtb_render(scene, &cam, &opt, accum, rgba, &stats));
    if (stats.n_paths != (uint64_t)W * H * 4 || stats.n_launches == 0) return 9;
    for (i = 0; i < W * H; ++i) {
        if (accum[4 * i + 3] != 4.0f || rgba[4 * i + 3] != 255) return 10;
        sum += accum[4 * i] + accum[4 * i + 1] + accum[4 * i + 2];
    }
    if (!(sum > 0.0) || !(sum < 3.0 * 4.0 * W * H)) return 11;

    /* the same frame through the multi-GPU entry point (two replicas on device 0), tile partition: same bits */
    CHECK(rtb_group_create(&desc, devices, 2, &group));
    CHECK(rtb_group_render(group, &cam, &opt, RTB_PARTITION_TILES, accum2, NULL, NULL));
    if (memcmp(accum, accum2, (size_t)W * H * 16) != 0) return 12;
    CHECK(rtb_group_destroy(group));

    /* argument errors come back as codes + messages */
    opt.traversal = 99;
    if (rtb_render(scene, &cam, &opt, accum, rgba, NULL) != RTB_ERR_UNSUPPORTED || rtb_last_error()[0] == 0) return 13;
    CHECK(rtb_scene_destroy(scene));
    free(accum); free(accum2); free(rgba);
    printf("abi_smoke ok: %u x %u x 4 spp, %u launches, mean radiance %.4f\n", W, H, stats.n_launches, sum / (3.0 * 4.0 * W * H));
    return 0;
}
