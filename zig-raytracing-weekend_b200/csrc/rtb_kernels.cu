// rtb_kernels.cu — CUDA kernels of the path-tracing hot path, sm_100a, compiled -fmad=false.
//
//   K1 render_megakernel : Camera.render + getRay + rayColor per pixel (src/camera.zig:93-208)
//   K3 trace_kernel      : world.hit for explicit ray batches           (src/bvh.zig:39-41)
//   K4 resolve_kernel    : toGamma2 + RGBA8 quantise                    (src/color.zig:43-62)
#include "rtb_kernels.cuh"

#include <cstdlib>

namespace rtb {

// ------------------------------------------------------------------------------------------
// K1 — megakernel.
//
// One thread owns one pixel for all the samples of this launch and keeps the running sum in
// registers, adding samples in sample order onto the value already in the accumulation buffer —
// exactly the order in which SharedStateImageWriter.writeColor adds them (src/camera.zig:54-56).
// Paths are regenerated in place: the loop body is "one ray segment"; when a lane's path ends it
// starts its next sample at once, so the warp stays converged on the traversal loop instead of
// idling until its longest path finishes.
// ------------------------------------------------------------------------------------------
// ORDERED: near-child-first per-octant layouts read from global memory (8 layouts do not fit in shared
// memory); Nearest.node is then an object index and the leaf records come from scene.prims.
template <bool SMEM_NODES, bool COUNT, bool QUADS, bool ORDERED>
__global__ void __launch_bounds__(kCtaThreads) render_megakernel(const RenderParams P) {
    float4* s_nodes = rtb_smem_nodes;
    const float4* __restrict__ nodes = P.scene.nodes;
    if (SMEM_NODES && !ORDERED) {
        for (uint32_t i = threadIdx.x; i < 2u * P.scene.n_nodes; i += kCtaThreads) s_nodes[i] = P.scene.nodes[i];
        __syncthreads();
        nodes = s_nodes;
    }

    const uint32_t tiles_x = (P.cam.width + kTileW - 1u) / kTileW;
    const uint32_t tile = blockIdx.x * P.tile_world + P.tile_rank;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t px = (tile % tiles_x) * kTileW + (warp & 3u) * 8u + (lane & 7u);
    const uint32_t py = (tile / tiles_x) * kTileH + (warp >> 2) * 4u + (lane >> 3);
    const uint32_t pixel = py * P.cam.width + px;
    const bool active = px < P.cam.width && py < P.cam.height && pixel >= P.pixel_begin && pixel < P.pixel_end;

    uint32_t n_rays = 0, n_box = 0, n_obj = 0, n_hits = 0;
    if (active && P.sample_count > 0u) {
        float4 acc = P.accum[pixel];
        const uint32_t s_end = P.sample_begin + P.sample_count;
        RngKey key;
        key.seed = P.seed;
        key.pixel = pixel;
        key.sample = P.sample_begin;
        if (P.cam.max_depth > 0u) {
            DRay ray = get_ray(P.cam, key);
            float3 T = f3(1.0f, 1.0f, 1.0f);
            float3 L = f3(0.0f, 0.0f, 0.0f);
            uint32_t segment = 1u;
            for (;;) {
                if (COUNT) ++n_rays;
                Nearest best;
                if (ORDERED && P.ordered == 3u) {  // RTB_TRAVERSAL_SAH16: packed half2 box nodes
                    const uint4* slots = P.scene.pk_nodes + (size_t)ray_octant_of_direction(ray.d) * P.scene.pk_slots;
                    const PackedRay pr = packed_ray_setup(P.scene, ray.o, ray.d);
                    best = traverse_packed<COUNT, QUADS, false>(slots, complex_tables(P.scene), ray.o, ray.d, ray.time, pr, 0.001f,
                                                                __int_as_float(0x7f800000), n_box, n_obj, 0u, key, segment);
                } else if (ORDERED) {
                    const float ix = 1.0f / ray.d.x, iy = 1.0f / ray.d.y, iz = 1.0f / ray.d.z;
                    const float4* oct =
                        P.scene.oct_nodes[P.ordered] + (size_t)ray_octant(ix, iy, iz) * 2u * (P.scene.oct_n_nodes[P.ordered] + 1u);
                    if (P.ordered == 2u)
                        best = traverse_octant<COUNT, QUADS, false, true>(oct, complex_tables(P.scene), ray.o, ray.d, ray.time, ix,
                                                                          iy, iz, 0.001f, __int_as_float(0x7f800000),
                                                                          n_box, n_obj, 0u, key, segment);
                    else
                        best = traverse_octant<COUNT, QUADS, false>(oct, complex_tables(P.scene), ray.o, ray.d, ray.time, ix, iy, iz,
                                                                    0.001f, __int_as_float(0x7f800000), n_box, n_obj, 0u,
                                                                    key, segment);
                } else {
                    best = traverse_reference<COUNT, QUADS>(nodes, P.scene.n_nodes, complex_tables(P.scene), ray, 0.001f,
                                                            __int_as_float(0x7f800000), n_box, n_obj, key, segment);
                }
                bool done;
                if (best.node == 0xffffffffu) {
                    L = L + T * miss_color(P.cam, ray);
                    done = true;
                } else {
                    if (COUNT) ++n_hits;
                    ShadeResult sr;
                    if (ORDERED) {
                        const float4* pr = P.scene.prims + 4u * (size_t)best.node;
                        sr = shade_rec<QUADS>(P.scene, pr[0], pr[1], pr[2], pr[3], ray, best.t, key, segment);
                    } else {
                        sr = shade<QUADS>(P.scene, nodes, ray, best, key, segment);
                    }
                    L = L + T * sr.emitted;
                    if (sr.scatters && segment < P.cam.max_depth) {
                        T = T * sr.attenuation;
                        ray = sr.scattered;
                        ++segment;
                        done = false;
                    } else {
                        done = true;
                    }
                }
                if (done) {
                    acc.x += L.x;
                    acc.y += L.y;
                    acc.z += L.z;
                    if (++key.sample == s_end) break;
                    ray = get_ray(P.cam, key);
                    T = f3(1.0f, 1.0f, 1.0f);
                    L = f3(0.0f, 0.0f, 0.0f);
                    segment = 1u;
                }
            }
        }
        acc.w = (float)s_end;  // buffer[i][3] = number_of_samples (src/camera.zig:56)
        P.accum[pixel] = acc;
    }
    if (COUNT) {
        unsigned long long v[4] = {n_rays, n_box, n_obj, n_hits};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned long long x = v[k];
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
            if (lane == 0 && x) atomicAdd(&P.counters[k], x);
        }
    }
}

template <bool SMEM, bool COUNT, bool QUADS, bool ORDERED = false>
static cudaError_t launch_mega_variant(const RenderParams& p, uint32_t grid, size_t smem, cudaStream_t stream) {
    auto kernel = render_megakernel<SMEM, COUNT, QUADS, ORDERED>;
    if (smem > 48u * 1024u) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<grid, kCtaThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

size_t megakernel_max_smem_nodes_bytes() { return 96u * 1024u; }

cudaError_t launch_megakernel(const RenderParams& p, bool nodes_in_smem, bool count_work, cudaStream_t stream,
                              LaunchInfo* info) {
    const uint32_t tiles_x = (p.cam.width + kTileW - 1u) / kTileW;
    const uint32_t tiles_y = (p.cam.height + kTileH - 1u) / kTileH;
    const uint32_t tiles = tiles_x * tiles_y;
    const uint32_t world = p.tile_world ? p.tile_world : 1u;
    if (p.tile_rank >= world) return cudaErrorInvalidValue;
    const uint32_t grid = (tiles > p.tile_rank) ? (tiles - p.tile_rank + world - 1u) / world : 0u;
    if (grid == 0u) return cudaSuccess;
    RenderParams q = p;
    q.tile_world = world;
    const size_t smem = nodes_in_smem ? (size_t)p.scene.n_nodes * 32u : 0u;
    const bool quads = p.scene.has_quads != 0u;
    cudaError_t e;
    if (p.ordered) {
        if (count_work) e = quads ? launch_mega_variant<false, true, true, true>(q, grid, 0, stream)
                                  : launch_mega_variant<false, true, false, true>(q, grid, 0, stream);
        else            e = quads ? launch_mega_variant<false, false, true, true>(q, grid, 0, stream)
                                  : launch_mega_variant<false, false, false, true>(q, grid, 0, stream);
        if (e == cudaSuccess && info) info->n_launches += 1;
        return e;
    }
#define RTB_DISPATCH(S, C, Q) e = launch_mega_variant<S, C, Q>(q, grid, smem, stream)
    if (nodes_in_smem) {
        if (count_work) { if (quads) RTB_DISPATCH(true, true, true); else RTB_DISPATCH(true, true, false); }
        else            { if (quads) RTB_DISPATCH(true, false, true); else RTB_DISPATCH(true, false, false); }
    } else {
        if (count_work) { if (quads) RTB_DISPATCH(false, true, true); else RTB_DISPATCH(false, true, false); }
        else            { if (quads) RTB_DISPATCH(false, false, true); else RTB_DISPATCH(false, false, false); }
    }
#undef RTB_DISPATCH
    if (e == cudaSuccess && info) info->n_launches += 1;
    return e;
}

// ------------------------------------------------------------------------------------------
// K3 — ray queries.  One thread per ray, reference visiting order, full HitRecord.
// ------------------------------------------------------------------------------------------
template <bool QUADS, bool ORDERED>
__global__ void __launch_bounds__(256) trace_kernel(const DevScene scene, const RtbRay* __restrict__ rays, uint64_t n,
                                                    RtbHit* __restrict__ hits, uint32_t layout) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RtbRay rr = rays[i];
    DRay r;
    r.o = f3(rr.origin[0], rr.origin[1], rr.origin[2]);
    r.d = f3(rr.direction[0], rr.direction[1], rr.direction[2]);
    r.time = rr.time;
    uint32_t n_box = 0, n_obj = 0;
    // ray queries key the constant-medium draw by the ray index (seed 0, sample 0, segment 1), see rtb.h
    RngKey qkey;
    qkey.seed = make_uint2(0u, 0u);
    qkey.pixel = (uint32_t)i;
    qkey.sample = 0u;
    Nearest best;
    if (ORDERED && layout == 3u) {
        const uint4* slots = scene.pk_nodes + (size_t)ray_octant_of_direction(r.d) * scene.pk_slots;
        const PackedRay pr = packed_ray_setup(scene, r.o, r.d);
        best = traverse_packed<true, QUADS, false>(slots, complex_tables(scene), r.o, r.d, r.time, pr, rr.t_min, rr.t_max, n_box,
                                                   n_obj, 0u, qkey, 1u);
    } else if (ORDERED) {
        const float ix = 1.0f / r.d.x, iy = 1.0f / r.d.y, iz = 1.0f / r.d.z;
        const float4* oct = scene.oct_nodes[layout] + (size_t)ray_octant(ix, iy, iz) * 2u * (scene.oct_n_nodes[layout] + 1u);
        if (layout == 2u)
            best = traverse_octant<true, QUADS, false, true>(oct, complex_tables(scene), r.o, r.d, r.time, ix, iy, iz, rr.t_min,
                                                             rr.t_max, n_box, n_obj, 0u, qkey, 1u);
        else
            best = traverse_octant<true, QUADS, false>(oct, complex_tables(scene), r.o, r.d, r.time, ix, iy, iz, rr.t_min, rr.t_max,
                                                       n_box, n_obj, 0u, qkey, 1u);
    } else {
        best = traverse_reference<true, QUADS>(scene.nodes, scene.n_nodes, complex_tables(scene), r, rr.t_min, rr.t_max, n_box,
                                               n_obj, qkey, 1u);
    }
    RtbHit h;
    h.object = -1;
    h.front_face = 0u;
    h.t = 0.0f;
    h.p[0] = h.p[1] = h.p[2] = 0.0f;
    h.normal[0] = h.normal[1] = h.normal[2] = 0.0f;
    h.u = h.v = 0.0f;
    if (best.node != 0xffffffffu) {
        const DHit d = ORDERED ? finish_hit_rec<QUADS, true>(scene.prims[4u * (size_t)best.node],
                                                             scene.prims[4u * (size_t)best.node + 1u], complex_tables(scene), r, best.t,
                                                             qkey, 1u)
                               : finish_hit<QUADS, true>(scene.nodes, complex_tables(scene), r, best, qkey, 1u);
        h.object = (int32_t)d.object;
        h.front_face = d.front_face ? 1u : 0u;
        h.t = d.t;
        h.p[0] = d.p.x; h.p[1] = d.p.y; h.p[2] = d.p.z;
        h.normal[0] = d.normal.x; h.normal[1] = d.normal.y; h.normal[2] = d.normal.z;
        h.u = d.u;
        h.v = d.v;
    }
    h.n_box_tests = n_box;
    h.n_object_tests = n_obj;
    hits[i] = h;
}

cudaError_t launch_trace(const DevScene& scene, const RtbRay* d_rays, uint64_t n, RtbHit* d_hits, uint32_t layout,
                         cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const uint32_t grid = (uint32_t)((n + 255u) / 256u);
    if (scene.has_quads) {
        if (layout) trace_kernel<true, true><<<grid, 256, 0, stream>>>(scene, d_rays, n, d_hits, layout);
        else        trace_kernel<true, false><<<grid, 256, 0, stream>>>(scene, d_rays, n, d_hits, 0u);
    } else {
        if (layout) trace_kernel<false, true><<<grid, 256, 0, stream>>>(scene, d_rays, n, d_hits, layout);
        else        trace_kernel<false, false><<<grid, 256, 0, stream>>>(scene, d_rays, n, d_hits, 0u);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K4 — resolve.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resolve_kernel(const float4* __restrict__ accum, uchar4* __restrict__ rgba,
                                                      uint64_t n_pixels, float n_override) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += stride) {
        const float4 a = accum[i];
        rgba[i] = quantise(a, n_override > 0.0f ? n_override : a.w);
    }
}

// SM count of the current device (grids are sized in multiples of it), queried once per device.
uint32_t current_sm_count() {
    static int cached[64] = {0};
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148u;
    if (cached[dev] == 0) {
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148u;
        cached[dev] = sms;
    }
    return (uint32_t)cached[dev];
}

cudaError_t launch_resolve(const float4* d_accum, uchar4* d_rgba, uint64_t n_pixels, float n_override,
                           cudaStream_t stream) {
    if (n_pixels == 0) return cudaSuccess;
    uint64_t blocks = (n_pixels + 255u) / 256u;
    const uint64_t cap = (uint64_t)current_sm_count() * 16u;
    if (blocks > cap) blocks = cap;
    resolve_kernel<<<(uint32_t)blocks, 256, 0, stream>>>(d_accum, d_rgba, n_pixels, n_override);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K5 — fused exchange + resolve over peer memory (multi-GPU).
//
// Every rank owns one contiguous slice of the frame.  For its slice it loads the float4 accumulators of ALL ranks
// (its own from local HBM, the others with plain loads through NVLink/NVSwitch peer mappings), adds them in rank
// order (deterministic), sets .w to the frame's samples per pixel, applies toGamma2 + quantisation and stores the
// sums and the RGBA8 pixels straight into the root rank's buffers (peer stores).  That is reduce-scatter + resolve +
// gather in one pass: every byte crosses NVLink once, no GPU receives more than (world-1)/world of a frame, and the
// resolve costs no extra trip through HBM.  The caller brackets the launch with a stream-ordered barrier.
// ------------------------------------------------------------------------------------------
template <int WORLD, int PPT>
__global__ void __launch_bounds__(256) exchange_resolve_kernel(const PeerAccums peers, uint32_t world,
                                                               float4* root_accum,  // may alias peers.p[root]
                                                               uchar4* __restrict__ root_rgba, uint64_t begin,
                                                               uint64_t end, float samples_per_pixel) {
    // PPT pixels per thread and iteration, 256 apart (coalesced), all WORLD * PPT loads issued before the first add:
    // the remote ones have NVLink latency, so bytes in flight are what sets the rate.
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * PPT;
    for (uint64_t i0 = begin + (uint64_t)blockIdx.x * blockDim.x * PPT + threadIdx.x; i0 < end; i0 += stride) {
        float4 v[WORLD > 0 ? WORLD : 1][PPT];
        float4 s[PPT];
        if (WORLD > 0) {
#pragma unroll
            for (int r = 0; r < WORLD; ++r)
#pragma unroll
                for (int k = 0; k < PPT; ++k) {
                    const uint64_t i = i0 + (uint64_t)k * 256u;
                    v[r][k] = i < end ? peers.p[r][i] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                s[k] = v[0][k];
#pragma unroll
                for (int r = 1; r < WORLD; ++r) {
                    s[k].x += v[r][k].x;
                    s[k].y += v[r][k].y;
                    s[k].z += v[r][k].z;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const uint64_t i = i0 + (uint64_t)k * 256u;
                s[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < end) {
                    s[k] = peers.p[0][i];
                    for (uint32_t r = 1; r < world; ++r) {
                        const float4 t = peers.p[r][i];
                        s[k].x += t.x;
                        s[k].y += t.y;
                        s[k].z += t.z;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const uint64_t i = i0 + (uint64_t)k * 256u;
            if (i < end) {
                s[k].w = samples_per_pixel;
                root_accum[i] = s[k];
                root_rgba[i] = quantise(s[k], samples_per_pixel);
            }
        }
    }
}

template <int PPT>
static void exchange_dispatch(const PeerAccums& peers, uint32_t world, float4* root_accum, uchar4* root_rgba,
                              uint64_t begin, uint64_t end, float spp, uint32_t g, cudaStream_t stream) {
    switch (world) {
        case 2: exchange_resolve_kernel<2, PPT><<<g, 256, 0, stream>>>(peers, world, root_accum, root_rgba, begin, end, spp); break;
        case 4: exchange_resolve_kernel<4, PPT><<<g, 256, 0, stream>>>(peers, world, root_accum, root_rgba, begin, end, spp); break;
        case 8: exchange_resolve_kernel<8, PPT><<<g, 256, 0, stream>>>(peers, world, root_accum, root_rgba, begin, end, spp); break;
        default: exchange_resolve_kernel<0, PPT><<<g, 256, 0, stream>>>(peers, world, root_accum, root_rgba, begin, end, spp); break;
    }
}

cudaError_t launch_exchange_resolve(const PeerAccums& peers, uint32_t world, float4* root_accum, uchar4* root_rgba,
                                    uint64_t begin, uint64_t end, float samples_per_pixel, cudaStream_t stream) {
    if (end <= begin) return cudaSuccess;
    static const int ppt = [] { const char* s = std::getenv("RTB_XCHG_PPT"); return s ? std::atoi(s) : 1; }();
    static const int bps = [] { const char* s = std::getenv("RTB_XCHG_BLOCKS_PER_SM"); return s ? std::atoi(s) : 8; }();
    const int p = ppt == 4 ? 4 : (ppt == 2 ? 2 : 1);
    uint64_t blocks = (end - begin + 256u * p - 1u) / (256u * p);
    const uint64_t cap = (uint64_t)current_sm_count() * (uint64_t)(bps > 0 ? bps : 8);
    if (blocks > cap) blocks = cap;
    const uint32_t g = (uint32_t)blocks;
    if (p == 4) exchange_dispatch<4>(peers, world, root_accum, root_rgba, begin, end, samples_per_pixel, g, stream);
    else if (p == 2) exchange_dispatch<2>(peers, world, root_accum, root_rgba, begin, end, samples_per_pixel, g, stream);
    else exchange_dispatch<1>(peers, world, root_accum, root_rgba, begin, end, samples_per_pixel, g, stream);
    return cudaGetLastError();
}

__global__ void philox_selftest_kernel(const uint4* __restrict__ ctr, uint2 key, uint32_t n, uint4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = philox4x32_10(ctr[i], key);
}

cudaError_t launch_philox_selftest(const uint4* d_ctr, uint2 key, uint32_t n, uint4* d_out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    philox_selftest_kernel<<<(n + 127u) / 128u, 128, 0, stream>>>(d_ctr, key, n, d_out);
    return cudaGetLastError();
}

// FP32 FMA peak: 8 independent dependent chains per thread, 3-register FFMA.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* __restrict__ out, uint32_t iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float b = 1.0f + blockIdx.x * 1e-9f, c = 1e-7f * (float)(threadIdx.x & 3u);
    for (uint32_t i = 0; i < iters; ++i) {
        a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
        a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

cudaError_t launch_ffma_peak(float* d_out, uint32_t grid, uint32_t iters, cudaStream_t stream) {
    ffma_peak_kernel<<<grid, 256, 0, stream>>>(d_out, iters);
    return cudaGetLastError();
}

}  // namespace rtb
