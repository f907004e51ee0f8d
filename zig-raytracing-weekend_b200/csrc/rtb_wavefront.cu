// rtb_wavefront.cu — K2: wavefront integrator, sm_100a, compiled -fmad=false.
//
// The same per-path arithmetic as the megakernel (rtb_device.cuh), split by stage, with the live rays
// kept in compacted SoA queues in HBM, BINNED BY RAY OCTANT (the three `invD < 0` predicates of
// Aabb.hit, src/aabb.zig:97):
//   wf_raygen      Camera.getRay for (pixel, sample) slots                (src/camera.zig:169-180)
//   wf_extend      world.hit for every queued ray                          (src/bvh.zig:122-136)
//   wf_shade       emitted + scatter; surviving rays go to the next bounce's queues
//                                                                          (src/camera.zig:191-207)
//   wf_accumulate  per-pixel sum of the finished samples IN SAMPLE ORDER   (src/camera.zig:54-56)
// All rays in the queues are at the same segment index (one bounce per iteration), so the segment —
// the RNG stream key — is a kernel argument and only the slot id travels with a ray; throughput and
// gathered radiance stay in per-slot arrays.
//
// Why octant bins (evidence: profiles/r1a_*, r1b_*): ncu showed the extend kernel issue-bound on the
// half-rate ALU pipe (FSEL/FSETP/FMNMX), not on memory.  A CTA that only sees rays of one octant can
// stage THAT octant's node layout in shared memory, in which every slab is already stored as (entry
// plane, exit plane): the reference's per-visit swap `if (invD < 0)` (6 FSEL + 3 FSETP) disappears,
// and the layout may also bake in near-child-first order (RTB_TRAVERSAL_ORDERED) while staying a
// stackless skip-link walk.  Three persistent variants of the kernel (per-lane refill from the queue,
// with and without leaf batching / while-while) were measured on B200 and all lost to this plain
// one-thread-per-ray loop, so they are not kept.
#include "rtb_wavefront.cuh"

#include <new>

namespace rtb {

constexpr uint32_t kOctants = 8;
constexpr uint32_t kChunk = 256;  // rays per CTA pass

struct WfQueue {
    float4* o_time;  // [octant][capacity] origin.xyz, time
    float4* d_slot;  // [octant][capacity] direction.xyz, bits(slot)
};

// One pipeline lane: the queues of one in-flight batch and the stream its kernels run on.
struct WfLane {
    size_t capacity = 0;  // slots
    WfQueue q[2]{};
    float2* hits = nullptr;      // [octant][capacity] t, bits(object)
    float4* T = nullptr;         // [capacity] throughput
    float4* L = nullptr;         // [capacity] radiance gathered so far; final radiance once the path ends
    uint32_t* counts = nullptr;  // [2][8] queue sizes
    cudaStream_t stream = nullptr;
    cudaEvent_t accumulated = nullptr;  // recorded after this lane's wf_accumulate
};

constexpr int kLanes = 3;

struct WavefrontState {
    WfLane lanes[kLanes];
    cudaEvent_t begin = nullptr;
    int sm_count = 0;
    bool ready = false;
};

struct WfParams {
    RenderParams R;
    WfQueue in, out;
    float2* hits;
    float4* T;
    float4* L;
    const uint32_t* count_in;  // 8
    uint32_t* count_out;       // 8
    uint32_t capacity;          // stride between octant bins
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

// Warp-aggregated push into one of 8 bins: one atomicAdd per (warp, octant present in the warp).
// Must be called by all 32 lanes.
__device__ __forceinline__ uint32_t queue_reserve(uint32_t* counts8, bool push, uint32_t octant) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t peers = __match_any_sync(0xffffffffu, push ? octant : kOctants);
    const uint32_t leader = __ffs(peers) - 1u;
    uint32_t base = 0;
    if (push && lane == leader) base = atomicAdd(&counts8[octant], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(peers & ((1u << lane) - 1u));
}

__device__ __forceinline__ void warp_add(unsigned long long* dst, unsigned long long x) {
    for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
    if ((threadIdx.x & 31u) == 0 && x) atomicAdd(dst, x);
}

__device__ __forceinline__ void push_ray(const WfParams& P, const DRay& ray, uint32_t slot, bool push) {
    const uint32_t oct = push ? ray_octant(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z) : 0u;
    const uint32_t j = queue_reserve(P.count_out, push, oct);
    if (push) {
        const size_t at = (size_t)oct * P.capacity + j;
        P.out.o_time[at] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.time);
        P.out.d_slot[at] = make_float4(ray.d.x, ray.d.y, ray.d.z, __uint_as_float(slot));
    }
}

__global__ void __launch_bounds__(256) wf_raygen(const WfParams P) {
    const uint32_t total = P.slots_per_sample * P.batch_samples;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += stride) {
        const uint32_t slot = base + threadIdx.x;
        bool push = false;
        DRay ray;
        ray.o = ray.d = f3(0.f, 0.f, 0.f);
        ray.time = 0.f;
        if (slot < total) {
            uint32_t pixel;
            if (slot_pixel(P.R, slot % P.slots_per_sample, pixel)) {
                P.L[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (P.R.cam.max_depth != 0u) {
                    RngKey key;
                    key.seed = P.R.seed;
                    key.pixel = pixel;
                    key.sample = P.batch_begin + slot / P.slots_per_sample;
                    ray = get_ray(P.R.cam, key);
                    P.T[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
                    push = true;
                }
            }
        }
        push_ray(P, ray, slot, push);
    }
}

// One thread per ray, plain node loop.  A CTA walks 256-ray chunks of the concatenated bins; chunks
// it visits are in non-decreasing octant order, so it re-stages the node layout at most 8 times.
template <bool SMEM_NODES, bool COUNT, bool QUADS>
__global__ void __launch_bounds__(256) wf_extend(const WfParams P) {
    extern __shared__ float4 s_nodes[];
    __shared__ uint32_t s_count[kOctants], s_first_chunk[kOctants + 1];
    const uint32_t n_nodes = P.R.scene.n_nodes;
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (uint32_t o = 0; o < kOctants; ++o) {
            s_count[o] = P.count_in[o];
            s_first_chunk[o] = acc;
            acc += (s_count[o] + kChunk - 1u) / kChunk;
        }
        s_first_chunk[kOctants] = acc;
    }
    // The bins this bounce's shade kernel will push into are not read by anyone now: clear them here.
    if (blockIdx.x == 0 && threadIdx.x < kOctants) P.count_out[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t total_chunks = s_first_chunk[kOctants];
    const float4* __restrict__ layouts = P.R.scene.oct_nodes[P.R.ordered];
    uint32_t staged = kOctants;  // octant whose layout is in shared memory
    uint32_t n_box = 0, n_obj = 0, n_rays = 0;
    for (uint32_t c = blockIdx.x; c < total_chunks; c += gridDim.x) {
        uint32_t oct = 0;
        while (c >= s_first_chunk[oct + 1u]) ++oct;
        const float4* __restrict__ nodes = layouts + (size_t)oct * 2u * n_nodes;
        if (SMEM_NODES) {
            if (oct != staged) {
                __syncthreads();
                for (uint32_t i = threadIdx.x; i < 2u * n_nodes; i += blockDim.x) s_nodes[i] = nodes[i];
                __syncthreads();
                staged = oct;
            }
            nodes = s_nodes;
        }
        const uint32_t i = (c - s_first_chunk[oct]) * kChunk + threadIdx.x;
        if (i < s_count[oct]) {
            const size_t at = (size_t)oct * P.capacity + i;
            const float4 a = P.in.o_time[at];
            const float4 b = P.in.d_slot[at];
            const float3 o = f3(a), d = f3(b);
            if (COUNT) ++n_rays;
            const Nearest best = traverse_octant<COUNT, QUADS>(nodes, n_nodes, P.R.scene.quads, o, d, a.w, 1.0f / d.x,
                                                               1.0f / d.y, 1.0f / d.z, 0.001f,
                                                               __int_as_float(0x7f800000), n_box, n_obj);
            P.hits[at] = make_float2(best.t, __uint_as_float(best.node));
        }
    }
    if (COUNT) {
        warp_add(&P.R.counters[0], n_rays);
        warp_add(&P.R.counters[1], n_box);
        warp_add(&P.R.counters[2], n_obj);
    }
}

template <bool COUNT, bool QUADS>
__global__ void __launch_bounds__(256) wf_shade(const WfParams P) {
    __shared__ uint32_t s_first[kOctants + 1];
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (uint32_t o = 0; o < kOctants; ++o) {
            s_first[o] = acc;
            acc += P.count_in[o];
        }
        s_first[kOctants] = acc;
    }
    __syncthreads();
    const uint32_t n = s_first[kOctants];
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t n_hits = 0;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += stride) {
        const uint32_t g = base + threadIdx.x;
        bool push = false;
        DRay next;
        next.o = next.d = f3(0.f, 0.f, 0.f);
        next.time = 0.f;
        uint32_t slot = 0;
        if (g < n) {
            uint32_t oct = 0;
            while (g >= s_first[oct + 1u]) ++oct;
            const size_t at = (size_t)oct * P.capacity + (g - s_first[oct]);
            const float4 a = P.in.o_time[at];
            const float4 b = P.in.d_slot[at];
            DRay r;
            r.o = f3(a);
            r.time = a.w;
            r.d = f3(b);
            slot = __float_as_uint(b.w);
            const float2 h = P.hits[at];
            Nearest best;
            best.t = h.x;
            best.node = __float_as_uint(h.y);
            float3 L = f3(P.L[slot]);
            const float3 T = f3(P.T[slot]);
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
                const ShadeResult sr = shade<QUADS>(P.R.scene, P.R.scene.prims, r, best, key, P.segment);
                L = L + T * sr.emitted;
                if (sr.scatters && P.segment < P.R.cam.max_depth) {
                    const float3 Tn = T * sr.attenuation;
                    P.T[slot] = make_float4(Tn.x, Tn.y, Tn.z, 0.f);
                    next = sr.scattered;
                    push = true;
                }
            }
            P.L[slot] = make_float4(L.x, L.y, L.z, 0.f);
        }
        push_ray(P, next, slot, push);
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
        const float4 l = P.L[(size_t)b * P.slots_per_sample + r];
        acc.x += l.x;
        acc.y += l.y;
        acc.z += l.z;
    }
    acc.w = (float)(P.batch_begin + P.batch_samples);
    P.R.accum[pixel] = acc;
}

// ------------------------------------------------------------------------------------------
WavefrontState* wavefront_create() { return new (std::nothrow) WavefrontState(); }

static void lane_free(WfLane* ln) {
    for (int k = 0; k < 2; ++k) {
        cudaFree(ln->q[k].o_time);
        cudaFree(ln->q[k].d_slot);
        ln->q[k] = WfQueue{};
    }
    cudaFree(ln->hits);
    cudaFree(ln->T);
    cudaFree(ln->L);
    ln->hits = nullptr;
    ln->T = ln->L = nullptr;
    ln->capacity = 0;
}

void wavefront_destroy(WavefrontState* st) {
    if (!st) return;
    for (WfLane& ln : st->lanes) {
        lane_free(&ln);
        cudaFree(ln.counts);
        if (ln.stream) cudaStreamDestroy(ln.stream);
        if (ln.accumulated) cudaEventDestroy(ln.accumulated);
    }
    if (st->begin) cudaEventDestroy(st->begin);
    delete st;
}

static cudaError_t wf_init(WavefrontState* st) {
    if (st->ready) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&st->sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&st->begin, cudaEventDisableTiming);
    for (WfLane& ln : st->lanes) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ln.accumulated, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc(&ln.counts, 2 * kOctants * sizeof(uint32_t));
    }
    st->ready = e == cudaSuccess;
    return e;
}

static cudaError_t lane_reserve(WfLane* ln, size_t capacity) {
    if (capacity <= ln->capacity) return cudaSuccess;
    cudaError_t e = cudaStreamSynchronize(ln->stream);  // nothing may still be using the old buffers
    if (e != cudaSuccess) return e;
    lane_free(ln);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        e = cudaMalloc(&ln->q[k].o_time, kOctants * capacity * sizeof(float4));
        if (e == cudaSuccess) e = cudaMalloc(&ln->q[k].d_slot, kOctants * capacity * sizeof(float4));
    }
    if (e == cudaSuccess) e = cudaMalloc(&ln->hits, kOctants * capacity * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&ln->T, capacity * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&ln->L, capacity * sizeof(float4));
    if (e != cudaSuccess) {
        lane_free(ln);
        return e;
    }
    ln->capacity = capacity;
    return cudaSuccess;
}

template <bool COUNT, bool QUADS>
static cudaError_t wf_launch_extend(const WfParams& P, bool smem_nodes, uint32_t grid, cudaStream_t stream) {
    if (smem_nodes) {
        const size_t smem = (size_t)P.R.scene.n_nodes * 32u;
        auto k = wf_extend<true, COUNT, QUADS>;
        if (smem > 40u * 1024u) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        k<<<grid, 256, smem, stream>>>(P);
    } else {
        wf_extend<false, COUNT, QUADS><<<grid, 256, 0, stream>>>(P);
    }
    return cudaGetLastError();
}

// Batches of ~8 M paths are pipelined over kLanes streams.  After the first ~10 bounces a batch is a
// thin, latency-bound tail (a few long paths; measured ~58 us per bounce for 40 bounces = 20 % of a
// batch when run alone); on its own stream that tail overlaps the next batches' full-width bounces.
// wf_accumulate calls are chained with events so every pixel still receives its samples in sample
// order, i.e. the result stays bit-identical to the megakernel's and to a single-stream run.
cudaError_t wavefront_render(WavefrontState* st, const RenderParams& p, bool count_work, cudaStream_t stream,
                             LaunchInfo* info) {
    const uint32_t tiles_x = (p.cam.width + kTileW - 1u) / kTileW;
    const uint32_t tiles_y = (p.cam.height + kTileH - 1u) / kTileH;
    const uint32_t tiles = tiles_x * tiles_y;
    const uint32_t world = p.tile_world ? p.tile_world : 1u;
    if (p.tile_rank >= world) return cudaErrorInvalidValue;
    const uint32_t owned = (tiles > p.tile_rank) ? (tiles - p.tile_rank + world - 1u) / world : 0u;
    if (owned == 0u || p.sample_count == 0u) return cudaSuccess;
    cudaError_t e = wf_init(st);
    if (e != cudaSuccess) return e;
    const uint32_t slots_per_sample = owned * kCtaThreads;
    // Samples in flight per pixel and batch (608 B of queue/state per path: 4.9 GB per lane at 8 M).
    const uint64_t target_paths = 8ull << 20;
    uint32_t B = (uint32_t)((target_paths + slots_per_sample - 1) / slots_per_sample);
    if (B < 1u) B = 1u;
    if (B > p.sample_count) B = p.sample_count;
    if ((uint64_t)B * slots_per_sample > 0x3fffffffull) B = (uint32_t)(0x3fffffffull / slots_per_sample);
    if (B < 1u) return cudaErrorInvalidValue;
    const uint32_t n_batches = (p.sample_count + B - 1u) / B;
    const int lanes_used = n_batches < (uint32_t)kLanes ? (int)n_batches : kLanes;
    for (int l = 0; l < lanes_used; ++l) {
        e = lane_reserve(&st->lanes[l], (size_t)B * slots_per_sample);
        if (e != cudaSuccess) return e;
    }

    const bool smem_nodes = p.scene.n_nodes > 0 && (size_t)p.scene.n_nodes * 32u <= megakernel_max_smem_nodes_bytes();
    const bool quads = p.scene.has_quads != 0u;
    const uint32_t max_grid = (uint32_t)st->sm_count * 8u;

    // fork: the lanes start after everything already queued on the caller's stream
    e = cudaEventRecord(st->begin, stream);
    for (int l = 0; l < lanes_used && e == cudaSuccess; ++l) e = cudaStreamWaitEvent(st->lanes[l].stream, st->begin, 0);
    if (e != cudaSuccess) return e;

    uint32_t launches = 0;
    int prev_lane = -1;
    for (uint32_t batch = 0; batch < n_batches; ++batch) {
        const uint32_t s0 = batch * B;
        const uint32_t nb = (p.sample_count - s0 < B) ? p.sample_count - s0 : B;
        const uint32_t cap = nb * slots_per_sample;
        const int lane_id = (int)(batch % (uint32_t)kLanes);
        WfLane& ln = st->lanes[lane_id];
        WfParams P{};
        P.R = p;
        P.R.tile_world = world;
        P.hits = ln.hits;
        P.T = ln.T;
        P.L = ln.L;
        P.capacity = (uint32_t)ln.capacity;
        P.slots_per_sample = slots_per_sample;
        P.batch_begin = p.sample_begin + s0;
        P.batch_samples = nb;
        e = cudaMemsetAsync(ln.counts, 0, 2 * kOctants * sizeof(uint32_t), ln.stream);
        if (e != cudaSuccess) return e;
        int cur = 0;
        P.out = ln.q[cur];
        P.count_out = ln.counts + cur * kOctants;
        uint32_t grid = (cap + 255u) / 256u + kOctants;  // chunks: every bin may end with a partial one
        if (grid > max_grid) grid = max_grid;
        wf_raygen<<<grid, 256, 0, ln.stream>>>(P);
        ++launches;
        for (uint32_t bounce = 0; bounce < p.cam.max_depth; ++bounce) {
            P.in = ln.q[cur];
            P.count_in = ln.counts + cur * kOctants;
            P.out = ln.q[cur ^ 1];
            P.count_out = ln.counts + (cur ^ 1) * kOctants;
            P.segment = bounce + 1u;
#define RTB_WF(C, Q)                                                      \
    do {                                                                  \
        e = wf_launch_extend<C, Q>(P, smem_nodes, grid, ln.stream);       \
        if (e == cudaSuccess) {                                           \
            wf_shade<C, Q><<<grid, 256, 0, ln.stream>>>(P);               \
            e = cudaGetLastError();                                       \
        }                                                                 \
    } while (0)
            if (count_work) { if (quads) RTB_WF(true, true); else RTB_WF(true, false); }
            else            { if (quads) RTB_WF(false, true); else RTB_WF(false, false); }
#undef RTB_WF
            if (e != cudaSuccess) return e;
            launches += 2;
            cur ^= 1;
            // Later bounces hold a small fraction of the rays: a smaller grid keeps the (mostly empty)
            // launches cheap.  Correct for any count: the kernels stride over the whole queue.
            if (bounce == 7u && grid > (uint32_t)st->sm_count * 2u) grid = (uint32_t)st->sm_count * 2u;
        }
        if (prev_lane >= 0 && prev_lane != lane_id) {
            e = cudaStreamWaitEvent(ln.stream, st->lanes[prev_lane].accumulated, 0);
            if (e != cudaSuccess) return e;
        }
        wf_accumulate<<<(slots_per_sample + 255u) / 256u, 256, 0, ln.stream>>>(P);
        ++launches;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventRecord(ln.accumulated, ln.stream);
        if (e != cudaSuccess) return e;
        prev_lane = lane_id;
    }
    // join: the caller's stream continues after the last accumulate (which follows all earlier ones);
    // the other lanes have nothing queued after their own accumulate, which precedes it in the chain.
    e = cudaStreamWaitEvent(stream, st->lanes[prev_lane].accumulated, 0);
    if (e != cudaSuccess) return e;
    if (info) info->n_launches += launches;
    return cudaSuccess;
}

}  // namespace rtb
