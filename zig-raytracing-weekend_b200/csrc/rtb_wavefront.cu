// rtb_wavefront.cu — K2: wavefront integrator, sm_100a, compiled -fmad=false.
//
// The same per-path arithmetic as the megakernel (rtb_device.cuh), split by stage with the live
// rays kept in compacted SoA queues in HBM:
//   wf_raygen      Camera.getRay for (pixel, sample) slots                (src/camera.zig:169-180)
//   wf_extend      world.hit for every queued ray                          (src/bvh.zig:122-136)
//   wf_shade       emitted + scatter, pushes surviving rays to the next queue, compacted
//                                                                          (src/camera.zig:191-207)
//   wf_accumulate  per-pixel sum of the finished samples IN SAMPLE ORDER   (src/camera.zig:54-56)
// All rays in a queue are at the same segment index (one bounce per iteration), so the segment —
// the RNG stream key — is a kernel argument, and nothing but the slot id has to travel with a ray.
#include "rtb_wavefront.cuh"

#include <new>

namespace rtb {

struct WfQueue {
    float4* o_time;  // origin.xyz, time
    float4* d_slot;  // direction.xyz, bits(slot)
    float4* T;       // throughput.xyz
    float4* L;       // radiance gathered so far.xyz
};

struct WavefrontState {
    size_t capacity = 0;  // slots
    WfQueue q[2]{};
    float2* hits = nullptr;    // t, bits(node)
    float4* stage = nullptr;   // finished radiance per slot
    uint32_t* counts = nullptr;  // [2] queue sizes
    uint32_t* h_count = nullptr;  // pinned readback
};

struct WfParams {
    RenderParams R;
    WfQueue in, out;
    float2* hits;
    float4* stage;
    const uint32_t* count_in;
    uint32_t* count_out;
    uint32_t slots_per_sample;  // owned tiles * 256
    uint32_t batch_begin;       // first sample of this batch
    uint32_t batch_samples;     // samples in flight per pixel in this batch
    uint32_t segment;           // 1-based segment index of the rays in `in`
};

__device__ __forceinline__ bool slot_pixel(const RenderParams& R, uint32_t r, uint32_t& pixel) {
    const uint32_t tiles_x = (R.cam.width + kTileW - 1u) / kTileW;
    const uint32_t tile = (r / kCtaThreads) * R.tile_world + R.tile_rank;
    const uint32_t t = r % kCtaThreads;
    const uint32_t warp = t >> 5, lane = t & 31u;
    const uint32_t px = (tile % tiles_x) * kTileW + (warp & 3u) * 8u + (lane & 7u);
    const uint32_t py = (tile / tiles_x) * kTileH + (warp >> 2) * 4u + (lane >> 3);
    pixel = py * R.cam.width + px;
    return px < R.cam.width && py < R.cam.height && pixel >= R.pixel_begin && pixel < R.pixel_end;
}

// Warp-aggregated queue push: one atomicAdd per warp.  Must be called by all 32 lanes.
__device__ __forceinline__ uint32_t queue_reserve(uint32_t* counter, bool push) {
    const uint32_t mask = __ballot_sync(0xffffffffu, push);
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == 0 && mask) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + __popc(mask & ((1u << lane) - 1u));
}

__device__ __forceinline__ void warp_add(unsigned long long* dst, unsigned long long x) {
    for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
    if ((threadIdx.x & 31u) == 0 && x) atomicAdd(dst, x);
}

__global__ void __launch_bounds__(256) wf_raygen(const WfParams P) {
    const uint32_t total = P.slots_per_sample * P.batch_samples;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += stride) {
        const uint32_t slot = base + threadIdx.x;
        bool push = false;
        DRay ray;
        if (slot < total) {
            uint32_t pixel;
            if (slot_pixel(P.R, slot % P.slots_per_sample, pixel)) {
                if (P.R.cam.max_depth == 0u) {
                    P.stage[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    RngKey key;
                    key.seed = P.R.seed;
                    key.pixel = pixel;
                    key.sample = P.batch_begin + slot / P.slots_per_sample;
                    ray = get_ray(P.R.cam, key);
                    push = true;
                }
            }
        }
        const uint32_t j = queue_reserve(P.count_out, push);
        if (push) {
            P.out.o_time[j] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.time);
            P.out.d_slot[j] = make_float4(ray.d.x, ray.d.y, ray.d.z, __uint_as_float(slot));
            P.out.T[j] = make_float4(1.f, 1.f, 1.f, 0.f);
            P.out.L[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

template <bool SMEM_NODES, bool COUNT, bool QUADS>
__global__ void __launch_bounds__(256) wf_extend(const WfParams P) {
    extern __shared__ float4 s_nodes[];
    const float4* __restrict__ nodes = P.R.scene.nodes;
    const uint32_t n = *P.count_in;
    if (blockIdx.x * blockDim.x >= n) return;
    if (SMEM_NODES) {
        for (uint32_t i = threadIdx.x; i < 2u * P.R.scene.n_nodes; i += blockDim.x) s_nodes[i] = P.R.scene.nodes[i];
        __syncthreads();
        nodes = s_nodes;
    }
    uint32_t n_box = 0, n_obj = 0, n_rays = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float4 a = P.in.o_time[i];
        const float4 b = P.in.d_slot[i];
        DRay r;
        r.o = f3(a);
        r.time = a.w;
        r.d = f3(b);
        if (COUNT) ++n_rays;
        const Nearest best = traverse_reference<COUNT, QUADS>(nodes, P.R.scene.n_nodes, P.R.scene.quads, r, 0.001f,
                                                              __int_as_float(0x7f800000), n_box, n_obj);
        P.hits[i] = make_float2(best.t, __uint_as_float(best.node));
    }
    if (COUNT) {
        warp_add(&P.R.counters[0], n_rays);
        warp_add(&P.R.counters[1], n_box);
        warp_add(&P.R.counters[2], n_obj);
    }
}

template <bool COUNT, bool QUADS>
__global__ void __launch_bounds__(256) wf_shade(const WfParams P) {
    const uint32_t n = *P.count_in;
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t n_hits = 0;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += stride) {
        const uint32_t i = base + threadIdx.x;
        bool push = false;
        DRay next;
        float3 T = f3(0.f, 0.f, 0.f), L = f3(0.f, 0.f, 0.f);
        uint32_t slot = 0;
        if (i < n) {
            const float4 a = P.in.o_time[i];
            const float4 b = P.in.d_slot[i];
            DRay r;
            r.o = f3(a);
            r.time = a.w;
            r.d = f3(b);
            slot = __float_as_uint(b.w);
            T = f3(P.in.T[i]);
            L = f3(P.in.L[i]);
            const float2 h = P.hits[i];
            Nearest best;
            best.t = h.x;
            best.node = __float_as_uint(h.y);
            if (best.node == 0xffffffffu) {
                L = L + T * miss_color(P.R.cam, r);
            } else {
                if (COUNT) ++n_hits;
                uint32_t pixel;
                slot_pixel(P.R, slot % P.slots_per_sample, pixel);
                RngKey key;
                key.seed = P.R.seed;
                key.pixel = pixel;
                key.sample = P.batch_begin + slot / P.slots_per_sample;
                const ShadeResult sr = shade<QUADS>(P.R.scene, P.R.scene.nodes, r, best, key, P.segment);
                L = L + T * sr.emitted;
                if (sr.scatters && P.segment < P.R.cam.max_depth) {
                    T = T * sr.attenuation;
                    next = sr.scattered;
                    push = true;
                }
            }
            if (!push) P.stage[slot] = make_float4(L.x, L.y, L.z, 0.f);
        }
        const uint32_t j = queue_reserve(P.count_out, push);
        if (push) {
            P.out.o_time[j] = make_float4(next.o.x, next.o.y, next.o.z, next.time);
            P.out.d_slot[j] = make_float4(next.d.x, next.d.y, next.d.z, __uint_as_float(slot));
            P.out.T[j] = make_float4(T.x, T.y, T.z, 0.f);
            P.out.L[j] = make_float4(L.x, L.y, L.z, 0.f);
        }
    }
    if (COUNT) warp_add(&P.R.counters[3], n_hits);
}

__global__ void __launch_bounds__(256) wf_accumulate(const WfParams P) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.slots_per_sample) return;
    uint32_t pixel;
    if (!slot_pixel(P.R, r, pixel)) return;
    float4 acc = P.R.accum[pixel];
    for (uint32_t b = 0; b < P.batch_samples; ++b) {
        const float4 l = P.stage[(size_t)b * P.slots_per_sample + r];
        acc.x += l.x;
        acc.y += l.y;
        acc.z += l.z;
    }
    acc.w = (float)(P.batch_begin + P.batch_samples);
    P.R.accum[pixel] = acc;
}

__global__ void wf_reset(uint32_t* counter) { *counter = 0u; }

// ------------------------------------------------------------------------------------------
WavefrontState* wavefront_create() { return new (std::nothrow) WavefrontState(); }

static void wf_free(WavefrontState* st) {
    for (int k = 0; k < 2; ++k) {
        cudaFree(st->q[k].o_time);
        cudaFree(st->q[k].d_slot);
        cudaFree(st->q[k].T);
        cudaFree(st->q[k].L);
        st->q[k] = WfQueue{};
    }
    cudaFree(st->hits);
    cudaFree(st->stage);
    st->hits = nullptr;
    st->stage = nullptr;
    st->capacity = 0;
}

void wavefront_destroy(WavefrontState* st) {
    if (!st) return;
    wf_free(st);
    cudaFree(st->counts);
    if (st->h_count) cudaFreeHost(st->h_count);
    delete st;
}

static cudaError_t wf_reserve(WavefrontState* st, size_t capacity) {
    cudaError_t e = cudaSuccess;
    if (!st->counts) {
        e = cudaMalloc(&st->counts, 2 * sizeof(uint32_t));
        if (e != cudaSuccess) return e;
        e = cudaMallocHost(&st->h_count, sizeof(uint32_t));
        if (e != cudaSuccess) return e;
    }
    if (capacity <= st->capacity) return cudaSuccess;
    wf_free(st);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        e = cudaMalloc(&st->q[k].o_time, capacity * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc(&st->q[k].d_slot, capacity * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc(&st->q[k].T, capacity * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc(&st->q[k].L, capacity * sizeof(float4));
    }
    if (e == cudaSuccess) e = cudaMalloc(&st->hits, capacity * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&st->stage, capacity * sizeof(float4));
    if (e != cudaSuccess) {
        wf_free(st);
        return e;
    }
    st->capacity = capacity;
    return cudaSuccess;
}

template <bool COUNT, bool QUADS>
static cudaError_t wf_launch_extend(const WfParams& P, bool smem_nodes, uint32_t grid, cudaStream_t stream) {
    if (smem_nodes) {
        const size_t smem = (size_t)P.R.scene.n_nodes * 32u;
        auto k = wf_extend<true, COUNT, QUADS>;
        if (smem > 48u * 1024u) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        k<<<grid, 256, smem, stream>>>(P);
    } else {
        wf_extend<false, COUNT, QUADS><<<grid, 256, 0, stream>>>(P);
    }
    return cudaGetLastError();
}

cudaError_t wavefront_render(WavefrontState* st, const RenderParams& p, bool count_work, cudaStream_t stream,
                             LaunchInfo* info) {
    const uint32_t tiles_x = (p.cam.width + kTileW - 1u) / kTileW;
    const uint32_t tiles_y = (p.cam.height + kTileH - 1u) / kTileH;
    const uint32_t tiles = tiles_x * tiles_y;
    const uint32_t world = p.tile_world ? p.tile_world : 1u;
    if (p.tile_rank >= world) return cudaErrorInvalidValue;
    const uint32_t owned = (tiles > p.tile_rank) ? (tiles - p.tile_rank + world - 1u) / world : 0u;
    if (owned == 0u || p.sample_count == 0u) return cudaSuccess;
    const uint32_t slots_per_sample = owned * kCtaThreads;
    // Samples in flight per pixel: enough paths to fill the machine (~4 M), bounded by memory.
    const uint64_t target_paths = 4ull << 20;
    uint32_t B = (uint32_t)((target_paths + slots_per_sample - 1) / slots_per_sample);
    if (B < 1u) B = 1u;
    if (B > p.sample_count) B = p.sample_count;
    if ((uint64_t)B * slots_per_sample > 0x7fffffffull) B = (uint32_t)(0x7fffffffull / slots_per_sample);
    if (B < 1u) return cudaErrorInvalidValue;
    cudaError_t e = wf_reserve(st, (size_t)B * slots_per_sample);
    if (e != cudaSuccess) return e;

    const bool smem_nodes = p.scene.n_nodes > 0 && (size_t)p.scene.n_nodes * 32u <= megakernel_max_smem_nodes_bytes();
    const bool quads = p.scene.has_quads != 0u;
    const uint32_t max_grid = 148u * 8u;

    WfParams P{};
    P.R = p;
    P.R.tile_world = world;
    P.hits = st->hits;
    P.stage = st->stage;
    P.slots_per_sample = slots_per_sample;
    uint32_t launches = 0;

    for (uint32_t s0 = 0; s0 < p.sample_count; s0 += B) {
        const uint32_t nb = (p.sample_count - s0 < B) ? p.sample_count - s0 : B;
        const uint32_t cap = nb * slots_per_sample;
        uint32_t grid = (cap + 255u) / 256u;
        if (grid > max_grid) grid = max_grid;
        P.batch_begin = p.sample_begin + s0;
        P.batch_samples = nb;
        e = cudaMemsetAsync(st->counts, 0, 2 * sizeof(uint32_t), stream);
        if (e != cudaSuccess) return e;
        int cur = 0;
        P.out = st->q[cur];
        P.count_out = st->counts + cur;
        wf_raygen<<<grid, 256, 0, stream>>>(P);
        ++launches;
        for (uint32_t bounce = 0; bounce < p.cam.max_depth; ++bounce) {
            P.in = st->q[cur];
            P.count_in = st->counts + cur;
            P.out = st->q[cur ^ 1];
            P.count_out = st->counts + (cur ^ 1);
            P.segment = bounce + 1u;
#define RTB_WF(C, Q)                                                        \
    do {                                                                    \
        e = wf_launch_extend<C, Q>(P, smem_nodes, grid, stream);            \
        if (e == cudaSuccess) {                                             \
            wf_shade<C, Q><<<grid, 256, 0, stream>>>(P);                    \
            e = cudaGetLastError();                                         \
        }                                                                   \
    } while (0)
            if (count_work) { if (quads) RTB_WF(true, true); else RTB_WF(true, false); }
            else            { if (quads) RTB_WF(false, true); else RTB_WF(false, false); }
#undef RTB_WF
            if (e != cudaSuccess) return e;
            wf_reset<<<1, 1, 0, stream>>>(st->counts + cur);
            launches += 3;
            cur ^= 1;
            // Every 4th bounce look at the queue size; stop when no ray is alive.
            if ((bounce & 3u) == 3u && bounce + 1u < p.cam.max_depth) {
                e = cudaMemcpyAsync(st->h_count, st->counts + cur, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
                if (e != cudaSuccess) return e;
                if (*st->h_count == 0u) break;
                const uint32_t live_grid = (*st->h_count + 255u) / 256u;
                grid = live_grid < max_grid ? live_grid : max_grid;
            }
        }
        wf_accumulate<<<(slots_per_sample + 255u) / 256u, 256, 0, stream>>>(P);
        ++launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (info) info->n_launches += launches;
    return cudaSuccess;
}

}  // namespace rtb
