// rtb_multi.cu — multi-GPU render behind ONE call of the C ABI (include/rtb.h: rtb_group_*).
//
// The reference renders one frame with 8 threads that share one SharedStateImageWriter buffer and are started by one
// call, startRender (src/main.zig:314-326; Camera.render per thread, src/camera.zig:93-116).  The drop-in equivalent
// across GPUs is one call as well: the caller hands over the scene once (rtb_group_create: one replica per device) and
// then calls rtb_group_render with its host buffers — no process launcher, rendezvous or IPC on the caller's side.
// One process, one host thread per device while rendering, the devices' accumulation buffers mapped into each other
// with cudaDeviceEnablePeerAccess (NVLink / NVSwitch), and the same fused exchange kernel as the multi-process path
// (rtb_exchange_resolve: every device sums its slice of the frame over all devices, resolves it and stores sums +
// RGBA8 into the root device's buffers).  Built on the public entry points of rtb.h only.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rtb.h"

extern "C" int rtb_set_last_error(int code, const char* message);  // rtb_api.cu (library-internal)

namespace {

int failf(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return rtb_set_last_error(code, buf);
}

// Balanced contiguous split of `total` samples over `n` ranks (the first total % n ranks get one more).
void sample_share(uint32_t total, uint32_t n, uint32_t r, uint32_t* begin, uint32_t* count) {
    const uint32_t base = total / n, rem = total % n;
    *count = base + (r < rem ? 1u : 0u);
    *begin = r * base + (r < rem ? r : rem);
}

}  // namespace

struct RtbSceneGroup {
    std::vector<int> devices;
    std::vector<RtbScene*> scenes;
    std::vector<float*> accum;  // one full-frame float4 buffer per member, on its device
    uint8_t* root_rgba = nullptr;
    uint64_t pixels = 0;  // capacity of the buffers
};

static void group_free_buffers(RtbSceneGroup* g) {
    for (size_t r = 0; r < g->accum.size(); ++r) {
        if (g->accum[r]) rtb_buffer_free(g->devices[r], g->accum[r]);
        g->accum[r] = nullptr;
    }
    if (g->root_rgba) rtb_buffer_free(g->devices[0], g->root_rgba);
    g->root_rgba = nullptr;
    g->pixels = 0;
}

extern "C" int rtb_group_create(const RtbSceneDesc* desc, const int* devices, uint32_t n_devices, RtbSceneGroup** group_out) {
    if (!group_out) return failf(RTB_ERR_INVALID_ARGUMENT, "group_out is NULL");
    *group_out = nullptr;
    if (!devices || n_devices == 0 || n_devices > 16) return failf(RTB_ERR_INVALID_ARGUMENT, "1..16 devices expected");
    RtbSceneGroup* g = new (std::nothrow) RtbSceneGroup();
    if (!g) return failf(RTB_ERR_OUT_OF_MEMORY, "host allocation failed");
    g->devices.assign(devices, devices + n_devices);
    g->scenes.assign(n_devices, nullptr);
    g->accum.assign(n_devices, nullptr);
    int rc = RTB_OK;
    for (uint32_t r = 0; r < n_devices && rc == RTB_OK; ++r) rc = rtb_scene_create(desc, devices[r], &g->scenes[r]);
    // every member reads every other member's buffer and writes the root's: map them into each other
    for (uint32_t a = 0; a < n_devices && rc == RTB_OK; ++a)
        for (uint32_t b = 0; b < n_devices && rc == RTB_OK; ++b) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            cudaError_t e = cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            if (e == cudaSuccess && !can)
                rc = failf(RTB_ERR_UNSUPPORTED, "device %d cannot map device %d's memory (no peer access)", devices[a], devices[b]);
            if (e == cudaSuccess && can) {
                e = cudaSetDevice(devices[a]);
                if (e == cudaSuccess) e = cudaDeviceEnablePeerAccess(devices[b], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) {
                    (void)cudaGetLastError();
                    e = cudaSuccess;
                }
            }
            if (e != cudaSuccess) rc = failf(RTB_ERR_CUDA, "peer access %d -> %d: %s", devices[a], devices[b], cudaGetErrorString(e));
        }
    if (rc != RTB_OK) {
        const std::string keep = rtb_last_error();
        for (RtbScene* s : g->scenes) rtb_scene_destroy(s);
        delete g;
        return rtb_set_last_error(rc, keep.c_str());
    }
    *group_out = g;
    return RTB_OK;
}

extern "C" int rtb_group_destroy(RtbSceneGroup* g) {
    if (!g) return RTB_OK;
    group_free_buffers(g);
    for (RtbScene* s : g->scenes) rtb_scene_destroy(s);
    delete g;
    return RTB_OK;
}

extern "C" int rtb_group_size(const RtbSceneGroup* g, uint32_t* n_out) {
    if (!g || !n_out) return failf(RTB_ERR_INVALID_ARGUMENT, "group / n_out is NULL");
    *n_out = (uint32_t)g->devices.size();
    return RTB_OK;
}

extern "C" int rtb_group_render(RtbSceneGroup* g, const RtbCamera* cam, const RtbRenderOptions* opt, uint32_t partition,
                                float* accum, uint8_t* rgba, RtbRenderStats* stats) {
    if (!g || !cam || !opt) return failf(RTB_ERR_INVALID_ARGUMENT, "group/camera/options is NULL");
    if (!accum) return failf(RTB_ERR_INVALID_ARGUMENT, "accum is NULL");
    if (partition != RTB_PARTITION_SAMPLES && partition != RTB_PARTITION_TILES)
        return failf(RTB_ERR_INVALID_ARGUMENT, "unknown partition %u", partition);
    if (opt->pixel_begin || opt->pixel_count || opt->tile_world > 1)
        return failf(RTB_ERR_INVALID_ARGUMENT, "rtb_group_render partitions the whole frame itself: leave the pixel / tile fields 0");
    const uint32_t n = (uint32_t)g->devices.size();
    const uint64_t npx = (uint64_t)cam->image_width * cam->image_height;
    if (npx == 0) return failf(RTB_ERR_INVALID_ARGUMENT, "empty image");
    const uint32_t total = opt->sample_count ? opt->sample_count : cam->samples_per_pixel;
    if (total == 0) return failf(RTB_ERR_INVALID_ARGUMENT, "no samples to render");
    if (g->pixels < npx) {  // (re)size the per-device frame buffers; they live as long as the group
        group_free_buffers(g);
        for (uint32_t r = 0; r < n; ++r) {
            void* p = nullptr;
            const int rc = rtb_buffer_alloc(g->devices[r], npx * 16, &p);
            if (rc != RTB_OK) {
                group_free_buffers(g);
                return rc;
            }
            g->accum[r] = static_cast<float*>(p);
        }
        void* p = nullptr;
        const int rc = rtb_buffer_alloc(g->devices[0], npx * 4, &p);
        if (rc != RTB_OK) {
            group_free_buffers(g);
            return rc;
        }
        g->root_rgba = static_cast<uint8_t*>(p);
        g->pixels = npx;
    }
    // render: one host thread per device (the enqueue of a wavefront render is thousands of launches)
    std::vector<int> status(n, RTB_OK);
    std::vector<std::string> message(n);
    std::vector<RtbRenderStats> st(n);
    std::vector<std::thread> workers;
    for (uint32_t r = 0; r < n; ++r) {
        workers.emplace_back([&, r]() {
            std::memset(&st[r], 0, sizeof(RtbRenderStats));
            cudaError_t e = cudaSetDevice(g->devices[r]);
            // the root starts from the caller's sums (samples are ADDED, as in rtb_render), the others from zero
            if (e == cudaSuccess)
                e = r == 0 ? cudaMemcpy(g->accum[0], accum, npx * 16, cudaMemcpyHostToDevice) : cudaMemset(g->accum[r], 0, npx * 16);
            if (e != cudaSuccess) {
                status[r] = RTB_ERR_CUDA;
                message[r] = std::string("frame buffer setup: ") + cudaGetErrorString(e);
                return;
            }
            RtbRenderOptions o = *opt;
            if (partition == RTB_PARTITION_SAMPLES) {
                uint32_t b = 0, c = 0;
                sample_share(total, n, r, &b, &c);
                if (c == 0) return;  // more devices than samples
                o.sample_begin = opt->sample_begin + b;
                o.sample_count = c;
            } else {
                o.sample_count = total;
                o.tile_rank = r;
                o.tile_world = n;
            }
            status[r] = rtb_render_device(g->scenes[r], cam, &o, g->accum[r], nullptr, &st[r]);
            if (status[r] != RTB_OK) message[r] = rtb_last_error();
        });
    }
    for (std::thread& w : workers) w.join();
    for (uint32_t r = 0; r < n; ++r)
        if (status[r] != RTB_OK) return rtb_set_last_error(status[r], (std::string("device ") + std::to_string(g->devices[r]) + ": " + message[r]).c_str());
    // exchange + resolve: every member combines its slice of the frame over all members into the root's buffers.
    // The renders above have completed (stats were read), so no barrier is needed before; the launches are
    // independent (disjoint slices) and are waited for below.
    std::vector<const float*> peers(n);
    for (uint32_t r = 0; r < n; ++r) peers[r] = g->accum[r];
    const float spp_total = (float)(opt->sample_begin + total);
    for (uint32_t r = 0; r < n; ++r) {
        const int rc = rtb_exchange_resolve(peers.data(), n, r, 0, g->accum[0], g->root_rgba, npx, spp_total, g->devices[r], nullptr);
        if (rc != RTB_OK) return rc;
    }
    for (uint32_t r = 0; r < n; ++r) {
        cudaError_t e = cudaSetDevice(g->devices[r]);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return failf(RTB_ERR_CUDA, "exchange on device %d: %s", g->devices[r], cudaGetErrorString(e));
    }
    cudaError_t e = cudaSetDevice(g->devices[0]);
    if (e == cudaSuccess) e = cudaMemcpy(accum, g->accum[0], npx * 16, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rgba) e = cudaMemcpy(rgba, g->root_rgba, npx * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return failf(RTB_ERR_CUDA, "rtb_group_render download: %s", cudaGetErrorString(e));
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        for (uint32_t r = 0; r < n; ++r) {
            stats->n_paths += st[r].n_paths;
            stats->n_rays += st[r].n_rays;
            stats->n_box_tests += st[r].n_box_tests;
            stats->n_object_tests += st[r].n_object_tests;
            stats->n_hits += st[r].n_hits;
            stats->n_launches += st[r].n_launches;
            if (st[r].device_ms > stats->device_ms) stats->device_ms = st[r].device_ms;  // the slowest member
        }
        stats->n_launches += n;
    }
    return RTB_OK;
}
