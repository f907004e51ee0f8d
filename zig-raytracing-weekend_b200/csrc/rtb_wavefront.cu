// rtb_wavefront.cu — K2: wavefront integrator, sm_100a, compiled -fmad=false.
//
// The same per-path arithmetic as the megakernel (rtb_device.cuh), split by stage, with the live rays
// kept in compacted SoA queues in HBM:
//   wf_raygen      Camera.getRay for (pixel, sample) slots                (src/camera.zig:169-180)
//   wf_extend      world.hit for every queued ray                          (src/bvh.zig:122-136)
//   wf_shade       emitted + scatter; surviving rays go to the next bounce's queues
//                                                                          (src/camera.zig:191-207)
//   wf_accumulate  per-pixel sum of the finished samples IN SAMPLE ORDER   (src/camera.zig:54-56)
// All rays in the queues are at the same segment index (one bounce per iteration), so the segment —
// the RNG stream key — is a kernel argument and only the slot id travels with a ray; throughput and
// gathered radiance stay in per-slot arrays.
//
// Two levels of binning keep warps coherent (evidence: profiles/r1a_*, r1c_*, r1d_*):
//   * RAY queues are binned by ray OCTANT (the three `invD < 0` predicates of Aabb.hit,
//     src/aabb.zig:97).  A CTA that only sees rays of one octant stages THAT octant's node layout in
//     shared memory, in which every slab is already stored as (entry plane, exit plane): the
//     reference's per-visit swap (6 FSEL + 3 FSETP on the half-rate ALU pipe) disappears, and the
//     layout may bake in near-child-first order while staying a stackless skip-link walk.
//   * HIT queues are binned by SHADING CLASS (miss / lambertian / metal / dielectric / other): ncu
//     showed the shade kernel executing every material's code in every warp (890 warp-instructions per
//     warp, 11-18 of 32 lanes active); with one class per 256-ray chunk a warp runs one material path.
// Once a batch has become thin (a fraction of a percent of its paths alive, bounce ~14 of 50 on Book-1) the
// remaining paths are finished by ONE launch of wf_tail instead of dozens of nearly empty kernel pairs.
// Five dynamic-ray-fetch variants of the extend kernel (per-lane and thresholded refill, with and without
// leaf batching / while-while, with a guard-free self-looping sentinel) were measured on B200 and all lost
// to the plain one-thread-per-ray loop, so they are not kept.  Two later attempts at the idle lanes ARE
// kept, off by default, because their images are bit-identical and their ncu counters are the evidence DESIGN.md §5.17
// and §5.23 argue from: wf_extend_stream (lane refill + parked leaves) and wf_extend_evict (straggler eviction).
// Batches of ~4 Mi paths are pipelined over eight streams ("lanes"), one wf_extend CTA per SM per launch, so that the
// kernels of different batches — at different bounces — share every SM.
#include "rtb_wavefront.cuh"

#include <new>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace rtb {

constexpr uint32_t kOctants = 8;
constexpr uint32_t kBins = 8;     // counters per binned queue set (8 octants; 5 shading classes, padded)
constexpr uint32_t kChunk = 256;  // hit records per CTA pass of wf_shade
// Threads (= rays per CTA pass) of wf_extend.  The octant layout staged in shared memory is shared by the whole CTA,
// so a larger CTA amortises it: with the 47 KB Book-1 SAH layout 256 threads allow 4 CTAs = 1024 threads per SM,
// 512 threads allow 3 CTAs = 1536.
#ifndef RTB_EXTEND_THREADS
#define RTB_EXTEND_THREADS 512
#endif
constexpr uint32_t kExtendThreads = RTB_EXTEND_THREADS;
// CTAs of ONE wf_extend launch per SM when the layout is staged in shared memory.  Three fit (registers), and a launch
// that takes all three slots was the default until r2r: but then every SM drains and refills at each kernel boundary
// (a CTA walks a contiguous share of the chunks; shares of different octants take different times).  With one CTA per SM
// per launch and twice as many lanes in flight, the three slots of an SM are held by launches of DIFFERENT batches at
// different bounces, whose ends do not coincide: 3 530 -> 4 030 Mpaths/s on Book-1 (profiles/r2r_ab.log, r2s_ab.log;
// capping the resident extend CTAs per SM to leave room for wf_shade instead made it slower, r2t_ab.log).
#ifndef RTB_EXTEND_GRID_PER_SM
#define RTB_EXTEND_GRID_PER_SM 1
#endif
// The same for layouts that stay in global memory (the 1 M-sphere scene: an L2-latency-bound walk that wants every
// warp it can get: 32 registers and four 512-thread CTAs per SM instead of 40 and three measured +15 %,
// profiles/r3i_million_ab.log).
#ifndef RTB_EXTEND_GRID_PER_SM_GLOBAL
#define RTB_EXTEND_GRID_PER_SM_GLOBAL 4
#endif
#ifndef RTB_SHADE_GRID_PER_SM
#define RTB_SHADE_GRID_PER_SM 5
#endif
// wf_shade's register budget.  Until r3b it was capped at 48 registers (5 CTAs per SM: it is bound by DRAM latency, and
// when it ran alone on an SM the extra warps measured +7 %), at the price of 120 B of spills in its hot paths.  With
// eight batches in flight the SM is shared with other launches anyway and issue slots are what the step is short of:
// 80 registers (3 CTAs, no spills) measured +2.5 % on Book-1 and +3.3 % on the textured scene, 64 registers +1 %,
// 128 registers -13 % (profiles/r3b_shade_occ_ab.log, r3c_shade_occ_ab.log); likewise 5 instead of 8 CTAs per SM
// in its grid (+1 %).
// RTB_SHADE_COOP=1: wf_shade draws the random unit vectors of a lambertian / metal chunk warp-cooperatively
// (random_unit_vector_coop: idle lanes evaluate the next tries of the paths whose rejection loop is still running).
// Bit-identical; measured (profiles/r3m_coop_ab.log, r3n_coop_*.csv): 29 instead of 21-24 of 32 lanes per instruction
// in wf_shade, but 95.5 M instead of 88.9 M warp-instructions on bounce 0 (the votes, the scratch exchange and the
// shuffles of a round cost more than the rounds saved) and the step is flat (-0.5 %).  Off by default.
#ifndef RTB_SHADE_COOP
#define RTB_SHADE_COOP 0
#endif
#ifndef RTB_SHADE_MINBLOCKS
#define RTB_SHADE_MINBLOCKS 3
#endif
#ifndef RTB_EXTEND_MINBLOCKS
#define RTB_EXTEND_MINBLOCKS 3
#endif

// 32-byte records everywhere (= one DRAM sector), because the shade kernel GATHERS them: ncu (r1e) showed
// it DRAM-bound at ~50 % of HBM peak fetching 16 B pieces out of separate arrays, two sectors per 32 B used.
struct WfQueue {
    float4* rays;  // [octant][capacity][2]: {origin.xyz, time}, {direction.xyz, bits(slot)}
};

// One pipeline lane: the queues of one in-flight batch and the stream its kernels run on.
struct WfLane {
    size_t capacity = 0;  // slots
    WfQueue q[2]{};
    uint4* hitq = nullptr;       // [class][capacity] {ray position, bits(t), object, path slot}
    float4* TL = nullptr;        // [capacity][2]: {T.xyz, L.x}, {L.y, L.z, -, -}: throughput and radiance so far
                                 // (the final radiance once the path has ended)
    uint32_t* counts = nullptr;  // [2][8] ray-queue sizes, then [2][8] hit-queue sizes
    cudaStream_t stream = nullptr;
    cudaEvent_t accumulated = nullptr;  // recorded after this lane's wf_accumulate
    // survivors[b] = rays entering bounce b of the lane's latest batch, written by wf_extend into mapped pinned host
    // memory (no copy, no synchronisation): how the host learns where the thin tail of a batch begins
    uint32_t* survivors_host = nullptr;
    uint32_t* survivors_dev = nullptr;
};
constexpr uint32_t kSurvivorSlots = 256;  // max_depth is a u8 in the reference (src/camera.zig:79)

#ifndef RTB_WF_LANES
#define RTB_WF_LANES 8
#endif
#ifndef RTB_WF_BATCH_LOG2
#define RTB_WF_BATCH_LOG2 22
#endif
constexpr int kLanes = RTB_WF_LANES;

struct WavefrontState {
    WfLane lanes[kLanes];
    // bounce at which a batch of the last render became thin enough for wf_tail (0 = not known yet) and what it was
    // learned for
    uint32_t tail_bounce = 0, tail_depth = 0, tail_slots = 0;
    const void* tail_scene = nullptr;
    cudaEvent_t begin = nullptr;
    int sm_count = 0;
    bool ready = false;
    // RTB_TIMELINE=<file>: per-warp residency records of the next render (development aid, tools/timeline.py)
    uint4* timeline = nullptr;
    uint32_t* timeline_count = nullptr;
};
constexpr uint32_t kTimelineCap = 24u << 20;

struct WfParams {
    RenderParams R;
    WfQueue in, out;
    uint4* hitq;
    float4* TL;
    const uint32_t* count_in;   // 8: ray bins of this bounce
    uint32_t* count_out;        // 8: ray bins of the next bounce
    uint32_t* hit_count;        // 8: hit bins of this bounce
    uint32_t* hit_count_next;   // 8: hit bins of the next bounce (cleared by this bounce's extend)
    uint32_t capacity;          // stride between bins
    uint32_t slots_per_sample;  // owned tiles * 256
    uint32_t batch_begin;       // first sample of this batch
    uint32_t batch_samples;     // samples in flight per pixel in this batch
    uint32_t segment;           // 1-based segment index of the rays in `in`
    uint32_t zero;              // always 0; only there to make an address opaque to ptxas (see traverse_octant)
    uint32_t* survivors;        // mapped host memory, [bounce] = rays entering that bounce (may be NULL)
    uint32_t perlin_smem;       // wf_shade: number of Perlin tables to stage in shared memory (0 = read them from global)
    // RTB_TIMELINE (development aid, NULL otherwise): every warp appends {start ns lo, hi, duration ns, kind | sm << 8 |
    // lane << 16 | bounce << 24} so that tools/timeline.py can reconstruct what was resident on each SM over time
    uint4* timeline;
    uint32_t* timeline_count;
    uint32_t timeline_cap;
    uint32_t timeline_tag;      // lane << 16 | bounce << 24
};

enum : uint32_t { TL_RAYGEN = 0u, TL_EXTEND = 1u, TL_SHADE = 2u, TL_TAIL = 3u };
// The start time waits in shared memory (one slot per warp), not in a register the kernel would carry through its loops.
struct TimelineScope {
    __device__ __forceinline__ explicit TimelineScope(const WfParams& P) {
        if (P.timeline && (threadIdx.x & 31u) == 0u) {
            unsigned long long t0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            slot()[threadIdx.x >> 5] = t0;
        }
    }
    static __device__ __forceinline__ unsigned long long* slot() {
        __shared__ unsigned long long t0s[32];
        return t0s;
    }
    __device__ __forceinline__ void finish(const WfParams& P, uint32_t kind) const {
        if (P.timeline && (threadIdx.x & 31u) == 0u) {
            unsigned long long t1;
            uint32_t sm;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            const unsigned long long t0 = slot()[threadIdx.x >> 5];
            const uint32_t at = atomicAdd(P.timeline_count, 1u);
            if (at < P.timeline_cap)
                P.timeline[at] = make_uint4((uint32_t)t0, (uint32_t)(t0 >> 32), (uint32_t)(t1 - t0), kind | (sm << 8) | P.timeline_tag);
        }
    }
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

// Warp-aggregated push into one of up to 8 bins: one atomicAdd per (warp, bin present in the warp).
// Must be called by all 32 lanes.
__device__ __forceinline__ uint32_t queue_reserve(uint32_t* counts8, bool push, uint32_t bin) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t peers = __match_any_sync(0xffffffffu, push ? bin : kBins);
    const uint32_t leader = __ffs(peers) - 1u;
    uint32_t base = 0;
    if (push && lane == leader) base = atomicAdd(&counts8[bin], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(peers & ((1u << lane) - 1u));
}

__device__ __forceinline__ void warp_add(unsigned long long* dst, unsigned long long x) {
    for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
    if ((threadIdx.x & 31u) == 0 && x) atomicAdd(dst, x);
}

__device__ __forceinline__ void push_ray(const WfParams& P, const DRay& ray, uint32_t slot, bool push) {
    const uint32_t oct = push ? ray_octant_of_direction(ray.d) : 0u;  // == ray_octant(1/d.x, 1/d.y, 1/d.z)
    const uint32_t j = queue_reserve(P.count_out, push, oct);
    if (push) {
        const size_t at = (size_t)oct * P.capacity + j;
        P.out.rays[2u * at] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.time);
        P.out.rays[2u * at + 1u] = make_float4(ray.d.x, ray.d.y, ray.d.z, __uint_as_float(slot));
    }
}

// Splits the concatenation of `n_bins` queues into 256-entry chunks; chunk c of a CTA's walk
// (c = blockIdx.x, += gridDim.x) lies in exactly one bin, and the bins it meets are non-decreasing.
struct ChunkMap {
    uint32_t count[kBins];
    uint32_t first_chunk[kBins + 1];
};
__device__ __forceinline__ void chunk_map_init(ChunkMap& m, const uint32_t* counts, uint32_t n_bins,
                                               uint32_t chunk = kChunk) {
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (uint32_t b = 0; b < kBins; ++b) {
            m.count[b] = b < n_bins ? counts[b] : 0u;
            m.first_chunk[b] = acc;
            acc += (m.count[b] + chunk - 1u) / chunk;
        }
        m.first_chunk[kBins] = acc;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) wf_raygen(const WfParams P) {
    const TimelineScope tl(P);
    const uint32_t total = P.slots_per_sample * P.batch_samples;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t tiles_x = (P.R.cam.width + kTileW - 1u) / kTileW;
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += stride) {
        // slots_per_sample is a multiple of the 256 slots of a tile and so is `base`: the whole block is in ONE tile of ONE
        // sample, and every integer division of slot -> (sample, tile, pixel) is block-uniform (seven per thread before).
        const uint32_t sample = base / P.slots_per_sample;
        const uint32_t tile = ((base - sample * P.slots_per_sample) / kCtaThreads) * P.R.tile_world + P.R.tile_rank;
        const uint32_t tile_y = tile / tiles_x, tile_x = tile - tile_y * tiles_x;
        const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
        const uint32_t px = tile_x * kTileW + (warp & 3u) * 8u + (lane & 7u);
        const uint32_t py = tile_y * kTileH + (warp >> 2) * 4u + (lane >> 3);
        const uint32_t pixel = py * P.R.cam.width + px;  // == slot_pixel(P.R, slot % slots_per_sample, .)
        const uint32_t slot = base + threadIdx.x;        // < total: total is a multiple of 256 too
        bool push = false;
        DRay ray;
        ray.o = ray.d = f3(0.f, 0.f, 0.f);
        ray.time = 0.f;
        if (px < P.R.cam.width && py < P.R.cam.height && pixel >= P.R.pixel_begin && pixel < P.R.pixel_end) {
            // the path's Philox key (pixel, sample) rides in the two spare words of its (T, L) record, so that
            // wf_shade does not have to recompute it from the slot (four 32-bit integer divisions per hit)
            const uint32_t sample_index = P.batch_begin + sample;
            P.TL[2u * (size_t)slot] = make_float4(1.f, 1.f, 1.f, 0.f);
            P.TL[2u * (size_t)slot + 1u] = make_float4(0.f, 0.f, __uint_as_float(pixel), __uint_as_float(sample_index));
            if (P.R.cam.max_depth != 0u) {
                RngKey key;
                key.seed = P.R.seed;
                key.pixel = pixel;
                key.sample = sample_index;
                ray = get_ray(P.R.cam, key, px + 1u, py + 1u);
                push = true;
            }
        }
        push_ray(P, ray, slot, push);
    }
    tl.finish(P, TL_RAYGEN);
}

// One thread per ray, plain node loop over the octant's layout; the result goes to the hit queue of
// the hit object's shading class.
// SLAB = how a box node is tested: kSlabExact the reference's arithmetic on pre-swapped planes (layouts 0 and 1),
// kSlabFma one FMA per plane on the library's own padded SAH tree (layout 2), kSlabPacked the 16-byte half2 nodes of
// RTB_TRAVERSAL_SAH16 (traverse_packed).
enum : int { kSlabExact = 0, kSlabFma = 1, kSlabPacked = 2 };
template <bool SMEM_NODES, bool COUNT, bool QUADS, int SLAB>
#ifndef RTB_EXTEND_MINBLOCKS_GLOBAL
#define RTB_EXTEND_MINBLOCKS_GLOBAL 4
#endif
__global__ void __launch_bounds__(kExtendThreads, SMEM_NODES ? RTB_EXTEND_MINBLOCKS : RTB_EXTEND_MINBLOCKS_GLOBAL)
wf_extend(const WfParams P) {
    __shared__ ChunkMap map;
    const TimelineScope tl(P);
    constexpr bool PACKED = SLAB == kSlabPacked;
    const uint32_t layout = PACKED ? 2u : P.R.ordered;
    const uint32_t n_nodes = P.R.scene.oct_n_nodes[layout];
    chunk_map_init(map, P.count_in, kOctants, kExtendThreads);
    // Bins nobody reads during this kernel: the ray bins this bounce's shade kernel will push into and
    // the hit bins of the next bounce.
    if (blockIdx.x == 0 && threadIdx.x < kBins) {
        P.count_out[threadIdx.x] = 0u;
        P.hit_count_next[threadIdx.x] = 0u;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && P.survivors && P.segment <= kSurvivorSlots) {
        uint32_t total = 0;
        for (uint32_t b = 0; b < kOctants; ++b) total += map.count[b];
        P.survivors[P.segment - 1u] = total;
    }
    const uint32_t total_chunks = map.first_chunk[kBins];
    // 16-byte units per octant: 2 per node (+ sentinel) for the float4 layouts, pk_slots for the packed one
    const size_t oct_stride = PACKED ? (size_t)P.R.scene.pk_slots : 2u * ((size_t)n_nodes + 1u);
    const float4* __restrict__ layouts =
        PACKED ? reinterpret_cast<const float4*>(P.R.scene.pk_nodes) : P.R.scene.oct_nodes[layout];
    uint32_t staged = kOctants;  // octant whose layout is in shared memory
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(rtb_smem_nodes) + P.zero;
    uint32_t n_box = 0, n_obj = 0, n_rays = 0;
    // A CTA walks a CONTIGUOUS range of chunks: the chunks are ordered by octant, so it stages one or two layouts per
    // launch instead of all eight (a strided walk meets every octant).  Layouts that stay in global memory
    // (RTB_EXTEND_STRIDED_GLOBAL) are walked strided instead: then all CTAs are in the same one or two octants at any
    // time, and those layouts — not all eight — are what the L2 has to hold.
#ifndef RTB_EXTEND_STRIDED_GLOBAL
#define RTB_EXTEND_STRIDED_GLOBAL 1
#endif
    constexpr bool STRIDED = !SMEM_NODES && RTB_EXTEND_STRIDED_GLOBAL;
    const uint32_t per_cta = (total_chunks + gridDim.x - 1u) / gridDim.x;
    const uint32_t c_last = (blockIdx.x + 1u) * per_cta < total_chunks ? (blockIdx.x + 1u) * per_cta : total_chunks;
    const uint32_t c_end = STRIDED ? total_chunks : c_last;
    const uint32_t c_step = STRIDED ? gridDim.x : 1u;
    for (uint32_t c = STRIDED ? blockIdx.x : blockIdx.x * per_cta; c < c_end; c += c_step) {
        uint32_t oct = 0;
        while (c >= map.first_chunk[oct + 1u]) ++oct;
        const float4* __restrict__ nodes = layouts + (size_t)oct * oct_stride;
        if (SMEM_NODES && oct != staged) {
            __syncthreads();
            // copy, turning the skip links (node indices) of interior nodes into shared-window addresses
            if (PACKED) {
                for (uint32_t i = threadIdx.x; i < (uint32_t)oct_stride; i += blockDim.x) {
                    float4 v = nodes[i];
                    const uint32_t meta = __float_as_uint(v.w);
                    // every slot: word 3 < 2^30 <=> box node (leaves set bit 30/31 in both of their slots); the skip links
                    // (slot indices) become byte distances from the node itself (see traverse_packed)
                    if (meta < (1u << 30)) v.w = __uint_as_float((meta - i) * 16u);  // distance to the skip target in bytes
                    rtb_smem_nodes[i] = v;
                }
            } else {  // a node becomes (entry, exit) pairs per axis for the packed-FP32 slab test (see traverse_octant)
                for (uint32_t n = threadIdx.x; n < (uint32_t)oct_stride / 2u; n += blockDim.x) {
                    const float4 f0 = nodes[2u * n], f1 = nodes[2u * n + 1u];
                    uint32_t meta = __float_as_uint(f0.w);
                    if (meta < (1u << 30)) meta = smem_base + meta * 32u;
                    rtb_smem_nodes[2u * n] = make_float4(f0.x, f1.x, f0.y, f1.y);
                    rtb_smem_nodes[2u * n + 1u] = make_float4(f0.z, f1.z, __uint_as_float(meta), f1.w);
                }
            }
            __syncthreads();
            staged = oct;
        }
        const uint32_t i = (c - map.first_chunk[oct]) * kExtendThreads + threadIdx.x;
        const bool valid = i < map.count[oct];
        uint32_t cls = CLASS_MISS;
        uint4 entry = make_uint4(0u, 0u, 0u, 0u);
        if (valid) {
            const size_t at = (size_t)oct * P.capacity + i;
            const float4 a = P.in.rays[2u * at];
            const float4 b = P.in.rays[2u * at + 1u];
            const float3 o = f3(a), d = f3(b);
            if (COUNT) ++n_rays;
            RngKey key = RngKey{};
            if (QUADS) {  // only complex objects (constant media) draw random numbers inside hit()
                const uint32_t slot = __float_as_uint(b.w);
                key.seed = P.R.seed;
                slot_pixel(P.R, slot % P.slots_per_sample, key.pixel);
                key.sample = P.batch_begin + slot / P.slots_per_sample;
            }
            Nearest best;
            if (PACKED) {
                const PackedRay pr = packed_ray_setup(P.R.scene, o, d);
                best = traverse_packed<COUNT, QUADS, SMEM_NODES>(reinterpret_cast<const uint4*>(nodes), complex_tables(P.R.scene), o,
                                                                 d, a.w, pr, 0.001f, __int_as_float(0x7f800000), n_box,
                                                                 n_obj, smem_base, key, P.segment);
            } else {
                best = traverse_octant<COUNT, QUADS, SMEM_NODES, SLAB == kSlabFma>(
                    nodes, complex_tables(P.R.scene), o, d, a.w, 1.0f / d.x, 1.0f / d.y, 1.0f / d.z, 0.001f,
                    __int_as_float(0x7f800000), n_box, n_obj, smem_base, key, P.segment);
            }
            if (best.node != 0xffffffffu) cls = P.R.scene.object_class[best.node];
            entry = make_uint4((uint32_t)at, __float_as_uint(best.t), best.node, __float_as_uint(b.w));
        }
        const uint32_t j = queue_reserve(P.hit_count, valid, cls);
        if (valid) P.hitq[(size_t)cls * P.capacity + j] = entry;
    }
    if (COUNT) {
        warp_add(&P.R.counters[0], n_rays);
        warp_add(&P.R.counters[1], n_box);
        warp_add(&P.R.counters[2], n_obj);
    }
    tl.finish(P, TL_EXTEND);
}

// ------------------------------------------------------------------------------------------------------------------
// wf_extend_evict — wf_extend over the packed SAH16 layout with STRAGGLER EVICTION.
//
// The compiler turns the node loop into phases — every lane walks to its next leaf (or to the end), the warp
// reconverges, the lanes at a leaf test it — and a warp stays until its longest walk has ended: on bounce rays the
// slab loop runs ~78 times per warp for rays that need 23 visits (tools/simt_model.cpp on real Book-1 rays), 11-16 of
// 32 lanes busy.  Refilling idle lanes was tried three times and lost to its per-visit overheads (DESIGN.md 5.14c,
// 5.17).  This kernel does the opposite and pays nothing per visit: at a phase boundary — where the warp is converged
// anyway — ONE vote counts the lanes still walking; once they are few (<= kEvictAt) they write their walk state
// {ray, node address, nearest hit so far} into a 32-entry buffer the warp owns in shared memory and the warp moves on
// to its next 32 rays.  When the buffer is nearly full its stragglers are finished together as one dense group (a
// straggler resumes exactly where it stopped — same node, same nearest hit, the same per-ray constants recomputed from
// the same ray — so every hit and every work counter is identical to wf_extend's).
// MEASURED (profiles/r3g_evict_*.csv, r3f_evict_ab.log): on bounce 1 the node loop runs 14 % fewer times, exactly as
// the SIMT model predicts (8.69 M instead of 10.09 M warp-level LDS), and 19.6 instead of 15.9 lanes are active per
// instruction — but the bookkeeping around it (the vote and the walking flags of every phase, the buffer, the second
// ray set-up and queue push of a resumed group) adds more warp-instructions than the loop saves: 176 M instead of 172 M
// on bounce 1, 171 M instead of 153 M on the coherent camera rays, and the whole step is 7 % SLOWER (3 995 -> 3 720
// Mpaths/s) whatever the threshold (5, 8, 12).  Off by default (RTB_EXTEND_EVICT=1 enables it); kept as the measured
// answer to "raise the lanes per instruction of wf_extend": it can be done, and it does not pay.
#ifndef RTB_EVICT_AT
#define RTB_EVICT_AT 8
#endif
constexpr uint32_t kEvictAt = RTB_EVICT_AT;
constexpr uint32_t kEvictFlushAt = 32u - kEvictAt;            // a buffer this full cannot take another eviction for sure
constexpr uint32_t kEvictWarpBytes = 32u * 48u;               // {o, time} {d, bits(slot)} {node address, t, object, queue position}

template <bool COUNT, bool SMEM>
__device__ __forceinline__ void extend_group(const WfParams& P, const uint4* __restrict__ slots, float4* __restrict__ pool,
                                             uint32_t& buffered, bool valid, float3 o, float3 d, float time,
                                             uint32_t slot_bits, uint32_t at, uint32_t i, float best_t, uint32_t best_node,
                                             bool may_evict, uint32_t& n_box, uint32_t& n_obj) {
    const uint32_t lane = threadIdx.x & 31u;
    const float t_min = 0.001f;
    constexpr uint32_t kStep = SMEM ? 16u : 1u;  // SMEM: `i` is a shared-window address, else a slot index
    const PackedRay pr = packed_ray_setup(P.R.scene, o, d);
    const __half2 ix = as_h2(pr.ix), iy = as_h2(pr.iy), iz = as_h2(pr.iz);
    const __half2 nx = as_h2(pr.nx), ny = as_h2(pr.ny), nz = as_h2(pr.nz);
    __half2 K = packed_interval(t_min, best_t, pr.sigma);
    bool walking = valid, evicted = false;
    const uint32_t n_start = __popc(__ballot_sync(0xffffffffu, valid));
    for (;;) {
        uint4 n = make_uint4(0u, 0u, 0u, RTB_META_END);
        if (walking) {
            for (;;) {  // to the next leaf, or to the end
                if (SMEM) {
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(n.x), "=r"(n.y), "=r"(n.z), "=r"(n.w) : "r"(i));
                } else {
                    n = slots[i];
                }
                if (n.w >= (1u << 30)) break;
                if (COUNT) ++n_box;
                const __half2 tx = __hfma2(as_h2(n.x), ix, nx);
                const __half2 ty = __hfma2(as_h2(n.y), iy, ny);
                const __half2 tz = __hfma2(as_h2(n.z), iz, nz);
                const __half2 r = __hmax2(__hmax2(tx, ty), __hmax2(tz, K));
                const bool miss = __hge(__high2half(r), __hneg(__low2half(r)));
                if (SMEM) i += miss ? n.w : kStep;  // staged links are byte distances from the node (see traverse_packed)
                else      i = miss ? n.w : i + kStep;
            }
            if (n.w == RTB_META_END) walking = false;
        }
        const uint32_t alive = __ballot_sync(0xffffffffu, walking);
        if (alive == 0u) break;
        const uint32_t n_alive = __popc(alive);
        // (a group that has lost no lane yet is not evicted: every eviction must come with progress)
        if (may_evict && n_alive <= kEvictAt && n_alive < n_start && buffered + n_alive <= 32u) {
            if (walking) {
                const uint32_t e = buffered + __popc(alive & ((1u << lane) - 1u));
                pool[3u * e] = make_float4(o.x, o.y, o.z, time);
                pool[3u * e + 1u] = make_float4(d.x, d.y, d.z, __uint_as_float(slot_bits));
                pool[3u * e + 2u] = make_float4(__uint_as_float(i), best_t, __uint_as_float(best_node), __uint_as_float(at));
                evicted = true;
            }
            buffered += n_alive;
            break;
        }
        if (walking) {  // the leaf the lane stopped at: the reference's sphere test (src/objects.zig:116-149)
            if (COUNT) ++n_obj;
            uint4 m;
            if (SMEM) {
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+16];" : "=r"(m.x), "=r"(m.y), "=r"(m.z), "=r"(m.w) : "r"(i));
            } else {
                m = slots[i + 1u];
            }
            const float3 c1 = f3(__uint_as_float(n.x), __uint_as_float(n.y), __uint_as_float(n.z));
            const float3 cv = f3(__uint_as_float(m.x), __uint_as_float(m.y), __uint_as_float(m.z));
            const float3 center = ((n.w >> 30) == KIND_MOVING_SPHERE) ? c1 + splat3(time) * cv : c1;
            float root;
            if (sphere_root_a(o, d, length_squared(d), center, __uint_as_float(m.w), t_min, best_t, root)) {
                best_t = root;
                best_node = n.w & RTB_META_INDEX_MASK;
                K = packed_interval(t_min, root, pr.sigma);
            }
            i += 2u * kStep;
        }
    }
    __syncwarp();  // the buffer's entries are read by other lanes when the stragglers resume
    const bool done = valid && !evicted;
    uint32_t cls = CLASS_MISS;
    if (done && best_node != 0xffffffffu) cls = P.R.scene.object_class[best_node];
    const uint32_t j = queue_reserve(P.hit_count, done, cls);
    if (done) P.hitq[(size_t)cls * P.capacity + j] = make_uint4(at, __float_as_uint(best_t), best_node, slot_bits);
}

// The stragglers in the warp's buffer as one dense group.  `again`: its own last few walkers may be evicted once more
// (they then wait for the next group); otherwise it runs to the end (before the layout changes, and at the very end).
template <bool COUNT, bool SMEM>
__device__ __forceinline__ void extend_resume(const WfParams& P, const uint4* __restrict__ slots, float4* __restrict__ pool,
                                              uint32_t& buffered, bool again, uint32_t& n_box, uint32_t& n_obj) {
    const uint32_t lane = threadIdx.x & 31u;
    const bool valid = lane < buffered;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
    if (valid) {
        a = pool[3u * lane];
        b = pool[3u * lane + 1u];
        c = pool[3u * lane + 2u];
    }
    __syncwarp();
    buffered = 0u;
    extend_group<COUNT, SMEM>(P, slots, pool, buffered, valid, f3(a), f3(b), a.w, __float_as_uint(b.w), __float_as_uint(c.w),
                              __float_as_uint(c.x), c.y, __float_as_uint(c.z), again, n_box, n_obj);
}

#ifndef RTB_EVICT_AGAIN
#define RTB_EVICT_AGAIN 1
#endif
template <bool COUNT, bool SMEM>
__global__ void __launch_bounds__(kExtendThreads, SMEM ? RTB_EXTEND_MINBLOCKS : RTB_EXTEND_MINBLOCKS_GLOBAL)
wf_extend_evict(const WfParams P) {
    __shared__ ChunkMap map;
    const TimelineScope tl(P);
    chunk_map_init(map, P.count_in, kOctants, kExtendThreads);
    if (blockIdx.x == 0 && threadIdx.x < kBins) {
        P.count_out[threadIdx.x] = 0u;
        P.hit_count_next[threadIdx.x] = 0u;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && P.survivors && P.segment <= kSurvivorSlots) {
        uint32_t total = 0;
        for (uint32_t b = 0; b < kOctants; ++b) total += map.count[b];
        P.survivors[P.segment - 1u] = total;
    }
    const uint32_t total_chunks = map.first_chunk[kBins];
    const uint32_t oct_stride = P.R.scene.pk_slots;
    const float4* __restrict__ layouts = reinterpret_cast<const float4*>(P.R.scene.pk_nodes);
    // dynamic shared memory: [the staged layout (SMEM only)] [one 32-entry straggler buffer per warp]
    float4* __restrict__ pool = rtb_smem_nodes + (SMEM ? oct_stride : 0u) + (threadIdx.x >> 5) * (kEvictWarpBytes / 16u);
    uint32_t current = kOctants, buffered = 0u;  // octant of the layout the buffered stragglers (and the staged copy) belong to
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(rtb_smem_nodes) + P.zero;
    const uint4* __restrict__ slots = nullptr;
    uint32_t n_box = 0, n_obj = 0, n_rays = 0;
    // SMEM: a contiguous range of chunks per CTA (one or two layouts to stage); global layouts: strided, so that all CTAs
    // are in the same one or two octants at any time (see wf_extend)
    const uint32_t per_cta = (total_chunks + gridDim.x - 1u) / gridDim.x;
    const uint32_t c_last = (blockIdx.x + 1u) * per_cta < total_chunks ? (blockIdx.x + 1u) * per_cta : total_chunks;
    const uint32_t c_end = SMEM ? c_last : total_chunks;
    const uint32_t c_step = SMEM ? 1u : gridDim.x;
    for (uint32_t c = SMEM ? blockIdx.x * per_cta : blockIdx.x; c < c_end; c += c_step) {
        uint32_t oct = 0;
        while (c >= map.first_chunk[oct + 1u]) ++oct;
        if (oct != current) {
            // the buffered stragglers hold positions in the layout that is about to be left
            if (buffered) extend_resume<COUNT, SMEM>(P, slots, pool, buffered, false, n_box, n_obj);
            slots = reinterpret_cast<const uint4*>(layouts + (size_t)oct * oct_stride);
            if (SMEM) {
                __syncthreads();
                const float4* __restrict__ nodes = layouts + (size_t)oct * oct_stride;
                for (uint32_t i = threadIdx.x; i < oct_stride; i += blockDim.x) {
                    float4 v = nodes[i];
                    const uint32_t meta = __float_as_uint(v.w);
                    if (meta < (1u << 30)) v.w = __uint_as_float((meta - i) * 16u);  // distance to the skip target in bytes
                    rtb_smem_nodes[i] = v;
                }
                __syncthreads();
            }
            current = oct;
        }
        const uint32_t i = (c - map.first_chunk[oct]) * kExtendThreads + threadIdx.x;
        const bool valid = i < map.count[oct];
        const size_t at = (size_t)oct * P.capacity + i;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (valid) {
            a = P.in.rays[2u * at];
            b = P.in.rays[2u * at + 1u];
            if (COUNT) ++n_rays;
        }
        extend_group<COUNT, SMEM>(P, slots, pool, buffered, valid, f3(a), f3(b), a.w, __float_as_uint(b.w), (uint32_t)at,
                                  SMEM ? smem_base : 0u, __int_as_float(0x7f800000), 0xffffffffu, true, n_box, n_obj);
        if (buffered >= kEvictFlushAt) extend_resume<COUNT, SMEM>(P, slots, pool, buffered, RTB_EVICT_AGAIN != 0, n_box, n_obj);
    }
    if (buffered) extend_resume<COUNT, SMEM>(P, slots, pool, buffered, false, n_box, n_obj);
    if (COUNT) {
        warp_add(&P.R.counters[0], n_rays);
        warp_add(&P.R.counters[1], n_box);
        warp_add(&P.R.counters[2], n_obj);
    }
    tl.finish(P, TL_EXTEND);
}

// ------------------------------------------------------------------------------------------------------------------
// wf_extend_stream — the extend kernel for INCOHERENT rays (bounces >= 1) over the packed SAH16 layout.
//
// ncu on the plain kernel (profiles/r2d_*): bounce rays keep 14 of 32 lanes busy.  The compiler turns its node loop
// into "walk to the next leaf, reconverge, test the leaf", and a warp pays the LONGEST leaf-to-leaf walk of its 32
// rays every time (~100 node iterations per warp for rays that need 24 each), then idles until its longest ray is done.
// A SIMT cost model on real Book-1 bounce rays (tools/simt_model.cpp) said two things have to change together:
//   * lanes must not wait at a leaf: a lane PARKS the leaf (up to kParked of them) and keeps walking with the old
//     t_max — the walk visits a superset, leaves are still tested in walk order, so the result is the same bits — and
//     the warp tests parked leaves together when some lane's parking is full (or too few lanes can still walk);
//   * a lane whose ray is done takes the next ray at once.  Earlier refill schemes lost to their own overhead
//      because the per-ray prologue (loads, packed_ray_setup) and epilogue (class lookup, queue
//     reservation) then ran with a handful of lanes; here both stay at FULL occupancy: a warp prepares 32 rays at a
//     time into a shared-memory staging area (all lanes take part, busy or not) and buffers finished rays there too,
//     flushing 32 results at a time.  Taking a ray = four LDS.128, retiring one = one STS.128.
// One CTA = 16 warps sharing one staged octant layout; warps claim 32-ray groups of the CTA's chunk range from a
// shared-memory counter.
#ifndef RTB_STREAM_MIN_WALKERS
#define RTB_STREAM_MIN_WALKERS 24
#endif
constexpr uint32_t kParked = 2;                                   // leaves a lane may park before it has to wait
constexpr uint32_t kStreamStageBytes = 32u * 64u;                 // per warp: 32 prepared rays x 4 quads (SoA by quad)
constexpr uint32_t kStreamResultBytes = 32u * 16u;                // per warp: 32 finished rays
constexpr uint32_t kStreamWarpBytes = kStreamStageBytes + kStreamResultBytes;

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Prepares rays [g, g + n) of octant bin `oct` (n <= 32) into the warp's staging area: quad q of ray j at
// stage + (q * 32 + j) * 16 = {o.xyz, time} {d.xyz, slot} {ix, iy, iz, nx} {ny, nz, sigma, queue position}.
// (Arguments by value: a reference to the kernel's parameter block would force a copy of it into local memory.)
struct PackedFrame {
    float center[3], inv_scale[3], scale[3];
};
static __device__ __noinline__ void stream_stage_fill(const float4* __restrict__ rays, PackedFrame fr, uint32_t stage,
                                                      uint32_t first, uint32_t n) {
    const uint32_t lane = threadIdx.x & 31u;
    if (lane < n) {
        const size_t at = (size_t)first + lane;
        const float4 a = rays[2u * at];
        const float4 b = rays[2u * at + 1u];
        float ix, iy, iz, nx, ny, nz, mx, my, mz;
        packed_axis_terms(a.x, b.x, fr.center[0], fr.inv_scale[0], fr.scale[0], ix, nx, mx);
        packed_axis_terms(a.y, b.y, fr.center[1], fr.inv_scale[1], fr.scale[1], iy, ny, my);
        packed_axis_terms(a.z, b.z, fr.center[2], fr.inv_scale[2], fr.scale[2], iz, nz, mz);
        const float sigma = packed_sigma(mx, my, mz);
        uint32_t pix, piy, piz, pnx, pny, pnz;
        packed_axis_pack(ix, nx, mx, sigma, pix, pnx);
        packed_axis_pack(iy, ny, my, sigma, piy, pny);
        packed_axis_pack(iz, nz, mz, sigma, piz, pnz);
        sts128(stage + (0u * 32u + lane) * 16u, make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w)));
        sts128(stage + (1u * 32u + lane) * 16u, make_uint4(__float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w)));
        sts128(stage + (2u * 32u + lane) * 16u, make_uint4(pix, piy, piz, pnx));
        sts128(stage + (3u * 32u + lane) * 16u, make_uint4(pny, pnz, __float_as_uint(sigma), (uint32_t)at));
    }
    __syncwarp();
}

// Pushes the warp's n buffered results {queue position, bits(t), object, slot} to the hit queues of their classes.
static __device__ __noinline__ void stream_flush(uint32_t* hit_count, uint4* __restrict__ hitq, uint32_t capacity,
                                                 const uint8_t* __restrict__ object_class, uint32_t results, uint32_t n) {
    __syncwarp();
    const uint32_t lane = threadIdx.x & 31u;
    const bool valid = lane < n;
    uint4 e = make_uint4(0u, 0u, 0u, 0u);
    uint32_t cls = CLASS_MISS;
    if (valid) {
        e = lds128(results + lane * 16u);
        if (e.z != 0xffffffffu) cls = object_class[e.z];
    }
    const uint32_t j = queue_reserve(hit_count, valid, cls);
    if (valid) hitq[(size_t)cls * capacity + j] = e;
    __syncwarp();
}

template <bool COUNT>
__global__ void __launch_bounds__(kExtendThreads, RTB_EXTEND_MINBLOCKS) wf_extend_stream(const WfParams P) {
    __shared__ ChunkMap map;
    __shared__ uint32_t seg_next;
    chunk_map_init(map, P.count_in, kOctants, kExtendThreads);
    if (blockIdx.x == 0 && threadIdx.x < kBins) {
        P.count_out[threadIdx.x] = 0u;
        P.hit_count_next[threadIdx.x] = 0u;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && P.survivors && P.segment <= kSurvivorSlots) {
        uint32_t total = 0;
        for (uint32_t b = 0; b < kOctants; ++b) total += map.count[b];
        P.survivors[P.segment - 1u] = total;
    }
    const uint32_t total_chunks = map.first_chunk[kBins];
    const uint32_t pk_slots = P.R.scene.pk_slots;
    const uint4* __restrict__ layouts = P.R.scene.pk_nodes;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(rtb_smem_nodes) + P.zero;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t stage = smem_base + pk_slots * 16u + warp * kStreamWarpBytes;
    const uint32_t results = stage + kStreamStageBytes;
    const uint32_t lt_mask = (1u << lane) - 1u;
    constexpr uint32_t kNone = 0xffffffffu;
    // lane state machine in one register: 0..kParked-1 = walking with that many leaves parked, kParked = parking full
    // (waiting for a leaf phase), + kDone once the walk has reached the end sentinel, kIdle = no ray
    constexpr uint32_t kDone = 4u, kIdle = 8u;
    static_assert(kParked < kDone, "state encoding");
    PackedFrame fr;
    for (int k = 0; k < 3; ++k) {
        fr.center[k] = P.R.scene.pk_center[k];
        fr.inv_scale[k] = P.R.scene.pk_inv_scale[k];
        fr.scale[k] = P.R.scene.pk_scale[k];
    }
    uint32_t staged = kOctants;
    uint32_t n_box = 0, n_obj = 0, n_rays = 0;

    const uint32_t per_cta = (total_chunks + gridDim.x - 1u) / gridDim.x;
    uint32_t c = blockIdx.x * per_cta;
    const uint32_t c_end = (blockIdx.x + 1u) * per_cta < total_chunks ? (blockIdx.x + 1u) * per_cta : total_chunks;
    while (c < c_end) {  // one iteration per octant this CTA's chunk range touches (block-uniform)
        uint32_t oct = 0;
        while (c >= map.first_chunk[oct + 1u]) ++oct;
        const uint32_t c_hi = c_end < map.first_chunk[oct + 1u] ? c_end : map.first_chunk[oct + 1u];
        const uint32_t r0 = (c - map.first_chunk[oct]) * kExtendThreads;
        uint32_t r1 = (c_hi - map.first_chunk[oct]) * kExtendThreads;
        if (r1 > map.count[oct]) r1 = map.count[oct];
        c = c_hi;
        __syncthreads();  // every warp has left the previous segment
        if (oct != staged) {
            const uint4* __restrict__ src = layouts + (size_t)oct * pk_slots;
            for (uint32_t k = threadIdx.x; k < pk_slots; k += blockDim.x) {
                uint4 v = src[k];
                if (v.w < (1u << 30)) v.w = smem_base + v.w * 16u;  // skip links become shared-window addresses
                sts128(smem_base + k * 16u, v);
            }
            staged = oct;
        }
        if (threadIdx.x == 0) seg_next = r0;
        __syncthreads();

        // ---- per-warp stream over the segment's rays ----
        uint32_t wstate = kIdle;
        float ox = 0.f, oy = 0.f, oz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, time = 0.f, sigma = 1.f;
        uint32_t slot = 0u, at = 0u, pix = 0u, piy = 0u, piz = 0u, pnx = 0u, pny = 0u, pnz = 0u;
        uint32_t i = smem_base, pend0 = 0u, pend1 = 0u, best_obj = kNone;
        float best_t = __int_as_float(0x7f800000);
        __half2 K = packed_interval(0.001f, best_t, 1.0f);
        uint32_t n_staged = 0u, taken = 0u, rcount = 0u;  // warp-uniform
        bool exhausted = false;
        const uint32_t bin_base = oct * P.capacity;
        for (;;) {
            // -- hand staged rays to idle lanes (refilling the staging area first when it is empty) --
            const uint32_t idle = __ballot_sync(0xffffffffu, wstate == kIdle);
            if (idle != 0u) {
                if (taken == n_staged && !exhausted) {
                    uint32_t g = 0u;
                    if (lane == 0u) g = atomicAdd(&seg_next, 32u);
                    g = __shfl_sync(0xffffffffu, g, 0);
                    taken = 0u;
                    if (g >= r1) {
                        exhausted = true;
                        n_staged = 0u;
                    } else {
                        n_staged = r1 - g < 32u ? r1 - g : 32u;
                        stream_stage_fill(P.in.rays, fr, stage, bin_base + g, n_staged);
                    }
                }
                const uint32_t avail = n_staged - taken;
                if (avail != 0u) {
                    const uint32_t rank = __popc(idle & lt_mask);
                    if (wstate == kIdle && rank < avail) {
                        const uint32_t rec = stage + (taken + rank) * 16u;
                        const uint4 q0 = lds128(rec), q1 = lds128(rec + 512u), q2 = lds128(rec + 1024u), q3 = lds128(rec + 1536u);
                        ox = __uint_as_float(q0.x); oy = __uint_as_float(q0.y); oz = __uint_as_float(q0.z); time = __uint_as_float(q0.w);
                        dx = __uint_as_float(q1.x); dy = __uint_as_float(q1.y); dz = __uint_as_float(q1.z); slot = q1.w;
                        pix = q2.x; piy = q2.y; piz = q2.z; pnx = q2.w;
                        pny = q3.x; pnz = q3.y; sigma = __uint_as_float(q3.z); at = q3.w;
                        wstate = 0u;
                        i = smem_base;
                        best_t = __int_as_float(0x7f800000);
                        best_obj = kNone;
                        K = packed_interval(0.001f, best_t, sigma);
                        if (COUNT) ++n_rays;
                    }
                    const uint32_t n_idle = __popc(idle);
                    taken += n_idle < avail ? n_idle : avail;
                }
            }
            const uint32_t n_active = 32u - __popc(__ballot_sync(0xffffffffu, wstate == kIdle));
            if (n_active == 0u) break;  // nothing staged, nothing in flight: the segment is done
            // -- walk: every lane that can, two nodes per vote; stop when too few lanes can still walk (the others
            //    are waiting for a leaf phase or for a new ray).  While rays can still be handed out all 32 lanes are
            //    active and the bar is RTB_STREAM_MIN_WALKERS; when the segment drains it follows the lanes left. --
            uint32_t min_walkers = (n_active * 3u) / 4u;
            if (min_walkers > (uint32_t)RTB_STREAM_MIN_WALKERS) min_walkers = (uint32_t)RTB_STREAM_MIN_WALKERS;
            if (min_walkers == 0u) min_walkers = 1u;
            for (;;) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (wstate < kParked) {
                        const uint4 n = lds128(i);
                        if (n.w < (1u << 30)) {
                            if (COUNT) ++n_box;
                            const __half2 tx = __hfma2(as_h2(n.x), as_h2(pix), as_h2(pnx));
                            const __half2 ty = __hfma2(as_h2(n.y), as_h2(piy), as_h2(pny));
                            const __half2 tz = __hfma2(as_h2(n.z), as_h2(piz), as_h2(pnz));
                            const __half2 r = __hmax2(__hmax2(tx, ty), __hmax2(tz, K));
                            const bool miss = __hge(__high2half(r), __hneg(__low2half(r)));
                            i = miss ? n.w : i + 16u;
                        } else if (n.w == RTB_META_END) {
                            wstate += kDone;
                        } else {  // a leaf: park it, keep walking
                            if (wstate == 0u) pend0 = i; else pend1 = i;
                            ++wstate;
                            i += 32u;
                        }
                    }
                }
                const uint32_t walkers = __ballot_sync(0xffffffffu, wstate < kParked);
                if ((uint32_t)__popc(walkers) < min_walkers) break;
            }
            // -- test the oldest parked leaf of every lane that has one (walk order is kept: first in, first tested) --
            if ((wstate & 3u) != 0u && wstate != kIdle) {
                if (COUNT) ++n_obj;
                const uint4 n = lds128(pend0), m = lds128(pend0 + 16u);
                const float3 o = f3(ox, oy, oz), d = f3(dx, dy, dz);
                const float3 c1 = f3(__uint_as_float(n.x), __uint_as_float(n.y), __uint_as_float(n.z));
                const float3 cv = f3(__uint_as_float(m.x), __uint_as_float(m.y), __uint_as_float(m.z));
                const float3 center = ((n.w >> 30) == KIND_MOVING_SPHERE) ? c1 + splat3(time) * cv : c1;
                float root;
                if (sphere_root_a(o, d, length_squared(d), center, __uint_as_float(m.w), 0.001f, best_t, root)) {
                    best_t = root;
                    best_obj = n.w & RTB_META_INDEX_MASK;
                    K = packed_interval(0.001f, root, sigma);
                }
                pend0 = pend1;
                --wstate;
            }
            // -- retire finished rays into the warp's result buffer (flushed 32 at a time, at full occupancy) --
            const bool fin = wstate == kDone;
            const uint32_t finm = __ballot_sync(0xffffffffu, fin);
            if (finm != 0u) {
                const uint32_t nf = __popc(finm);
                if (rcount + nf > 32u) {
                    stream_flush(P.hit_count, P.hitq, P.capacity, P.R.scene.object_class, results, rcount);
                    rcount = 0u;
                }
                if (fin) {
                    sts128(results + (rcount + __popc(finm & lt_mask)) * 16u, make_uint4(at, __float_as_uint(best_t), best_obj, slot));
                    wstate = kIdle;
                }
                rcount += nf;
            }
        }
        stream_flush(P.hit_count, P.hitq, P.capacity, P.R.scene.object_class, results, rcount);
    }
    if (COUNT) {
        warp_add(&P.R.counters[0], n_rays);
        warp_add(&P.R.counters[1], n_box);
        warp_add(&P.R.counters[2], n_obj);
    }
}

// One thread per hit record; every 256-record chunk belongs to one shading class.
template <bool COUNT, bool QUADS>
__global__ void __launch_bounds__(256, RTB_SHADE_MINBLOCKS) wf_shade(const WfParams P) {
    __shared__ ChunkMap map;
    __shared__ uint4 coop_scratch[8][32];  // random_unit_vector_coop: one row per warp
    const TimelineScope tl(P);
    chunk_map_init(map, P.hit_count, kShadeClasses);
    const uint32_t total_chunks = map.first_chunk[kBins];
    uint32_t n_hits = 0;
    // Perlin tables (src/perlin.zig:76-81: 256 gradients + 3 x 256 permutation entries per NoiseTexture, 4 864 B packed)
    // are gathered 7 x 8 x 4 times per noise-textured hit with data-dependent indices: a CTA that meets a chunk of
    // textured hits (CLASS_OTHER) stages them in shared memory once (north_star: "perlin tables in constant or shared
    // memory"; __constant__ would serialise the divergent indices).
    DevScene scene = P.R.scene;
    bool perlin_staged = false;
    for (uint32_t c = blockIdx.x; c < total_chunks; c += gridDim.x) {
        uint32_t cls = 0;
        while (c >= map.first_chunk[cls + 1u]) ++cls;
        if (P.perlin_smem != 0u && cls == CLASS_OTHER && !perlin_staged) {  // block-uniform
            const uint32_t n16 = P.perlin_smem * (uint32_t)(sizeof(DevPerlin) / 16u);
            const float4* __restrict__ src = reinterpret_cast<const float4*>(P.R.scene.perlins);
            for (uint32_t k = threadIdx.x; k < n16; k += blockDim.x) rtb_smem_nodes[k] = src[k];
            __syncthreads();
            scene.perlins = reinterpret_cast<const DevPerlin*>(rtb_smem_nodes);
            perlin_staged = true;
        }
        const uint32_t i = (c - map.first_chunk[cls]) * kChunk + threadIdx.x;
        bool push = false;
        DRay next;
        next.o = next.d = f3(0.f, 0.f, 0.f);
        next.time = 0.f;
        uint32_t slot = 0;
        const bool valid = i < map.count[cls];
        uint4 e = make_uint4(0u, 0u, 0u, 0u);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, tl0 = a, tl1 = a;
        if (valid) {
            e = P.hitq[(size_t)cls * P.capacity + i];
            a = P.in.rays[2u * (size_t)e.x];
            b = P.in.rays[2u * (size_t)e.x + 1u];
            slot = e.w;  // == bits(b.w); carried in the hit record so that the (T, L) gather does not wait for the ray
            tl0 = P.TL[2u * (size_t)slot];
            tl1 = P.TL[2u * (size_t)slot + 1u];
        }
        RngKey key;
        key.seed = P.R.seed;
        key.pixel = __float_as_uint(tl1.z);
        key.sample = __float_as_uint(tl1.w);
        // a chunk of solid-lambertian or metal hits: every lane needs a randomUnitVector; the warp draws them together
        float3 unit = f3(0.f, 0.f, 0.f);
        const bool coop = RTB_SHADE_COOP && (cls == CLASS_LAMBERT_SOLID || cls == CLASS_METAL);  // block-uniform
        if (coop) unit = random_unit_vector_coop(valid, key, P.segment, coop_scratch[threadIdx.x >> 5]);
        if (valid) {
            DRay r;
            r.o = f3(a);
            r.time = a.w;
            r.d = f3(b);
            float3 T = f3(tl0);
            float3 L = f3(tl0.w, tl1.x, tl1.y);
            if (cls == CLASS_MISS) {  // block-uniform
                L = L + T * miss_color(P.R.cam, r);
            } else {
                if (COUNT) ++n_hits;
                const float4* __restrict__ pr = P.R.scene.prims + 4u * (size_t)e.z;
                const float4 f0 = pr[0], f1 = pr[1], m0 = pr[2], m1 = pr[3];
                const ShadeResult sr =
                    shade_rec<QUADS>(scene, f0, f1, m0, m1, r, __uint_as_float(e.y), key, P.segment, coop ? &unit : nullptr);
                L = L + T * sr.emitted;
                if (sr.scatters && P.segment < P.R.cam.max_depth) {
                    T = T * sr.attenuation;
                    next = sr.scattered;
                    push = true;
                }
            }
            P.TL[2u * (size_t)slot] = make_float4(T.x, T.y, T.z, L.x);
            P.TL[2u * (size_t)slot + 1u] = make_float4(L.y, L.z, tl1.z, tl1.w);
        }
        push_ray(P, next, slot, push);
    }
    if (COUNT) warp_add(&P.R.counters[3], n_hits);
    tl.finish(P, TL_SHADE);
}

// The thin tail of a batch.  After ~12 bounces of the Book-1 scene fewer than 1 % of the paths are alive, but the
// per-bounce kernel pair still costs its launch + staging latency 38 more times (measured: bounces 12..50 = 10 % of the
// step for < 1 % of the work).  wf_tail finishes every path that is still alive in ONE launch: a thread takes one ray of
// the current queue and runs the rest of that path like the megakernel does — same traversal, same shading, same Philox
// streams, so the result is bit-identical whichever bounce the switch happens at.
template <bool COUNT, bool QUADS, int SLAB>
__global__ void __launch_bounds__(256) wf_tail(const WfParams P) {
    __shared__ ChunkMap map;
    const TimelineScope tl(P);
    chunk_map_init(map, P.count_in, kOctants);
    constexpr bool PACKED = SLAB == kSlabPacked;
    const uint32_t total_chunks = map.first_chunk[kBins];
    const size_t oct_stride = PACKED ? (size_t)P.R.scene.pk_slots : 2u * ((size_t)P.R.scene.oct_n_nodes[P.R.ordered] + 1u);
    const float4* __restrict__ layouts =
        PACKED ? reinterpret_cast<const float4*>(P.R.scene.pk_nodes) : P.R.scene.oct_nodes[P.R.ordered];
    uint32_t n_rays = 0, n_box = 0, n_obj = 0, n_hits = 0;
    for (uint32_t c = blockIdx.x; c < total_chunks; c += gridDim.x) {
        uint32_t bin = 0;
        while (c >= map.first_chunk[bin + 1u]) ++bin;
        const uint32_t i = (c - map.first_chunk[bin]) * kChunk + threadIdx.x;
        if (i >= map.count[bin]) continue;
        const size_t at = (size_t)bin * P.capacity + i;
        const float4 a = P.in.rays[2u * at];
        const float4 b = P.in.rays[2u * at + 1u];
        DRay ray;
        ray.o = f3(a);
        ray.time = a.w;
        ray.d = f3(b);
        const uint32_t slot = __float_as_uint(b.w);
        const float4 tl0 = P.TL[2u * (size_t)slot];
        const float4 tl1 = P.TL[2u * (size_t)slot + 1u];
        float3 T = f3(tl0);
        float3 L = f3(tl0.w, tl1.x, tl1.y);
        RngKey key;
        key.seed = P.R.seed;
        key.pixel = __float_as_uint(tl1.z);
        key.sample = __float_as_uint(tl1.w);
        uint32_t segment = P.segment;
        for (;;) {
            if (COUNT) ++n_rays;
            Nearest best;
            if (PACKED) {
                const uint4* __restrict__ slots =
                    reinterpret_cast<const uint4*>(layouts) + (size_t)ray_octant_of_direction(ray.d) * oct_stride;
                const PackedRay pr = packed_ray_setup(P.R.scene, ray.o, ray.d);
                best = traverse_packed<COUNT, QUADS, false>(slots, complex_tables(P.R.scene), ray.o, ray.d, ray.time, pr, 0.001f,
                                                            __int_as_float(0x7f800000), n_box, n_obj, 0u, key, segment);
            } else {
                const float ix = 1.0f / ray.d.x, iy = 1.0f / ray.d.y, iz = 1.0f / ray.d.z;
                const float4* __restrict__ nodes = layouts + (size_t)ray_octant(ix, iy, iz) * oct_stride;
                best = traverse_octant<COUNT, QUADS, false, SLAB == kSlabFma>(nodes, complex_tables(P.R.scene), ray.o, ray.d, ray.time,
                                                                             ix, iy, iz, 0.001f, __int_as_float(0x7f800000),
                                                                             n_box, n_obj, 0u, key, segment);
            }
            if (best.node == 0xffffffffu) {
                L = L + T * miss_color(P.R.cam, ray);
                break;
            }
            if (COUNT) ++n_hits;
            const float4* __restrict__ pr = P.R.scene.prims + 4u * (size_t)best.node;
            const ShadeResult sr = shade_rec<QUADS>(P.R.scene, pr[0], pr[1], pr[2], pr[3], ray, best.t, key, segment);
            L = L + T * sr.emitted;
            if (!(sr.scatters && segment < P.R.cam.max_depth)) break;
            T = T * sr.attenuation;
            ray = sr.scattered;
            ++segment;
        }
        P.TL[2u * (size_t)slot] = make_float4(T.x, T.y, T.z, L.x);
        P.TL[2u * (size_t)slot + 1u] = make_float4(L.y, L.z, tl1.z, tl1.w);
    }
    if (COUNT) {
        warp_add(&P.R.counters[0], n_rays);
        warp_add(&P.R.counters[1], n_box);
        warp_add(&P.R.counters[2], n_obj);
        warp_add(&P.R.counters[3], n_hits);
    }
    tl.finish(P, TL_TAIL);
}

__global__ void __launch_bounds__(256) wf_accumulate(const WfParams P) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.slots_per_sample) return;
    uint32_t pixel;
    if (!slot_pixel(P.R, r, pixel)) return;
    float4 acc = P.R.accum[pixel];
    for (uint32_t b = 0; b < P.batch_samples; ++b) {
        const size_t s = (size_t)b * P.slots_per_sample + r;
        const float4 tl0 = P.TL[2u * s];
        const float4 tl1 = P.TL[2u * s + 1u];
        acc.x += tl0.w;
        acc.y += tl1.x;
        acc.z += tl1.y;
    }
    acc.w = (float)(P.batch_begin + P.batch_samples);
    P.R.accum[pixel] = acc;
}

// ------------------------------------------------------------------------------------------
WavefrontState* wavefront_create() { return new (std::nothrow) WavefrontState(); }

static void lane_free(WfLane* ln) {
    for (int k = 0; k < 2; ++k) {
        cudaFree(ln->q[k].rays);
        ln->q[k] = WfQueue{};
    }
    cudaFree(ln->hitq);
    cudaFree(ln->TL);
    ln->hitq = nullptr;
    ln->TL = nullptr;
    ln->capacity = 0;
}

void wavefront_destroy(WavefrontState* st) {
    if (!st) return;
    for (WfLane& ln : st->lanes) {
        lane_free(&ln);
        cudaFree(ln.counts);
        if (ln.survivors_host) cudaFreeHost(ln.survivors_host);
        if (ln.stream) cudaStreamDestroy(ln.stream);
        if (ln.accumulated) cudaEventDestroy(ln.accumulated);
    }
    if (st->begin) cudaEventDestroy(st->begin);
    cudaFree(st->timeline);
    cudaFree(st->timeline_count);
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
        if (e == cudaSuccess) e = cudaMalloc(&ln.counts, 4 * kBins * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaHostAlloc(&ln.survivors_host, kSurvivorSlots * sizeof(uint32_t), cudaHostAllocMapped);
        if (e == cudaSuccess) {
            std::memset(ln.survivors_host, 0xff, kSurvivorSlots * sizeof(uint32_t));
            e = cudaHostGetDevicePointer(&ln.survivors_dev, ln.survivors_host, 0);
        }
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
        e = cudaMalloc(&ln->q[k].rays, 2 * kOctants * capacity * sizeof(float4));
    }
    if (e == cudaSuccess) e = cudaMalloc(&ln->hitq, kShadeClasses * capacity * sizeof(uint4));
    if (e == cudaSuccess) e = cudaMalloc(&ln->TL, 2 * capacity * sizeof(float4));
    if (e != cudaSuccess) {
        lane_free(ln);
        return e;
    }
    ln->capacity = capacity;
    return cudaSuccess;
}

// Bytes of one octant's layout as the extend kernel stages it.
static size_t wf_layout_bytes(const DevScene& sc, uint32_t ordered) {
    return ordered == 3u ? (size_t)sc.pk_slots * 16u : ((size_t)sc.oct_n_nodes[ordered] + 1u) * 32u;
}
template <bool COUNT, bool QUADS, int SLAB>
static cudaError_t wf_launch_extend_slab(const WfParams& P, bool smem_nodes, uint32_t grid, cudaStream_t stream) {
    if (smem_nodes) {
        const size_t smem = wf_layout_bytes(P.R.scene, P.R.ordered);
        auto k = wf_extend<true, COUNT, QUADS, SLAB>;
        if (smem > 40u * 1024u) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        k<<<grid, kExtendThreads, smem, stream>>>(P);
    } else {
        wf_extend<false, COUNT, QUADS, SLAB><<<grid, kExtendThreads, 0, stream>>>(P);
    }
    return cudaGetLastError();
}
// RTB_EXTEND_STREAM: 0 = never use wf_extend_stream (default: measured slower, see its header and DESIGN.md section 5),
// 1 = for bounces >= 1, 2 = for every bounce.
static int wf_stream_mode() {
    static const int v = [] { const char* s = std::getenv("RTB_EXTEND_STREAM"); return s && s[0] ? std::atoi(s) : 0; }();
    return v;
}
template <bool COUNT>
static cudaError_t wf_launch_extend_stream(const WfParams& P, uint32_t grid, cudaStream_t stream) {
    const size_t smem = (size_t)P.R.scene.pk_slots * 16u + (kExtendThreads / 32u) * kStreamWarpBytes;
    auto k = wf_extend_stream<COUNT>;
    if (smem > 40u * 1024u) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k<<<grid, kExtendThreads, smem, stream>>>(P);
    return cudaGetLastError();
}
// RTB_EXTEND_EVICT=1 renders SAH16 scenes (spheres only) with wf_extend_evict — over the staged layout, or over the global
// one for large scenes.  Off by default: measured slower in both (see the kernel's header; the 1 M-sphere scene, whose
// walk is bound by L2 latency at 5.6 lanes per instruction: 290 -> 275 Mpaths/s for thresholds 4 ... 20,
// profiles/r3k_million_evict_ab.log).
static int wf_evict_mode() {
    static const int v = [] { const char* s = std::getenv("RTB_EXTEND_EVICT"); return s && s[0] ? std::atoi(s) : -1; }();
    return v;
}
static size_t wf_evict_smem_bytes(const DevScene& sc, bool smem_nodes) {
    return (smem_nodes ? (size_t)sc.pk_slots * 16u : 0u) + (kExtendThreads / 32u) * kEvictWarpBytes;
}
template <bool COUNT, bool SMEM>
static cudaError_t wf_launch_extend_evict(const WfParams& P, uint32_t grid, cudaStream_t stream) {
    const size_t smem = wf_evict_smem_bytes(P.R.scene, SMEM);
    auto k = wf_extend_evict<COUNT, SMEM>;
    if (smem > 40u * 1024u) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k<<<grid, kExtendThreads, smem, stream>>>(P);
    return cudaGetLastError();
}
template <bool COUNT, bool QUADS>
static cudaError_t wf_launch_extend(const WfParams& P, bool smem_nodes, uint32_t grid, cudaStream_t stream) {
    if (P.R.ordered == 3u && !QUADS && wf_stream_mode() == 0) {
        const int mode = wf_evict_mode();
        // three CTAs of the shared-memory variant (layout + 24 KB of straggler buffers each) must still fit an SM
        if (smem_nodes && mode == 1 && wf_evict_smem_bytes(P.R.scene, true) <= 72u * 1024u)
            return wf_launch_extend_evict<COUNT, true>(P, grid, stream);
        if (!smem_nodes && mode == 1) return wf_launch_extend_evict<COUNT, false>(P, grid, stream);
    }
    if (P.R.ordered == 3u && smem_nodes && !QUADS && wf_stream_mode() >= (P.segment == 1u ? 2 : 1))
        return wf_launch_extend_stream<COUNT>(P, grid, stream);
    if (P.R.ordered == 3u) return wf_launch_extend_slab<COUNT, QUADS, kSlabPacked>(P, smem_nodes, grid, stream);
    return P.R.ordered == 2u ? wf_launch_extend_slab<COUNT, QUADS, kSlabFma>(P, smem_nodes, grid, stream)
                             : wf_launch_extend_slab<COUNT, QUADS, kSlabExact>(P, smem_nodes, grid, stream);
}

// Batches of ~4 M paths are pipelined over kLanes (8) streams.  After the first ~10 bounces a batch is a
// thin, latency-bound tail (a few long paths; measured ~58 us per bounce for 40 bounces = 20 % of a
// batch when run alone); on its own stream that tail overlaps the next batches' full-width bounces.
// wf_accumulate calls are chained with events so every pixel still receives its samples in sample
// order, i.e. the result stays bit-identical to the megakernel's and to a single-stream run.
// First bounce at which at most 1/128 of the batch's paths (and at least a few thousand rays' worth of launches) are
// still alive; max_depth + 1 if the batch never gets that thin.  survivors[b] = rays that entered bounce b.
static uint32_t choose_tail_bounce(const uint32_t* survivors, uint32_t max_depth) {
    const uint32_t first = survivors[0];
    if (first == 0xffffffffu) return max_depth + 1u;
    const uint32_t thresh = first / 128u > 4096u ? first / 128u : 4096u;
    for (uint32_t b = 2; b < max_depth && b < kSurvivorSlots; ++b)
        if (survivors[b] != 0xffffffffu && survivors[b] <= thresh) return b;
    return max_depth + 1u;
}
// RTB_WF_TAIL_BOUNCE=k forces the switch bounce (k = 0: never switch, every bounce stays a kernel pair); used for A/B
// measurements and by the test that the result does not depend on where the switch happens.  -1 = not forced.
static int wf_forced_tail_bounce() {
    static const int v = [] { const char* s = std::getenv("RTB_WF_TAIL_BOUNCE"); return s && s[0] ? std::atoi(s) : -1; }();
    return v;
}

// Most path slots one pass may hold per sample (624 B each).  A frame with more owned pixels than this is rendered
// in several passes over interleaved subsets of its tiles (every pixel belongs to exactly one pass, so the per-pixel
// sample order is unchanged).  RTB_WF_MAX_SLOTS overrides it (tests force the multi-pass path on a small frame).
static uint32_t wf_max_slots_per_sample() {
    static const uint32_t v = [] {
        const char* s = std::getenv("RTB_WF_MAX_SLOTS");
        const unsigned long long x = s ? std::strtoull(s, nullptr, 10) : 0ull;
        return (x >= 256ull && x <= 0x1fffffffull) ? (uint32_t)x : (10u << 20);  // x 624 B x 8 lanes = 52 GB
    }();
    return v;
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
    if ((uint64_t)owned * kCtaThreads > wf_max_slots_per_sample()) {
        const uint64_t k64 = ((uint64_t)owned * kCtaThreads + wf_max_slots_per_sample() - 1u) / wf_max_slots_per_sample();
        if (k64 * world > 0xffffffffull) return cudaErrorInvalidValue;
        const uint32_t k = (uint32_t)k64;
        for (uint32_t j = 0; j < k; ++j) {  // pass j owns the tiles t with t % (world * k) == tile_rank + j * world
            RenderParams q = p;
            q.tile_world = world * k;
            q.tile_rank = p.tile_rank + j * world;
            const cudaError_t ej = wavefront_render(st, q, count_work, stream, info);
            if (ej != cudaSuccess) return ej;
        }
        return cudaSuccess;
    }
    cudaError_t e = wf_init(st);
    if (e != cudaSuccess) return e;
    const uint32_t slots_per_sample = owned * kCtaThreads;
    // Samples in flight per pixel and batch (624 B of queue/state per path slot: 2.6 GB per lane at 4 Mi).
    const uint64_t target_paths = 1ull << RTB_WF_BATCH_LOG2;
    uint32_t B = (uint32_t)((target_paths + slots_per_sample - 1) / slots_per_sample);
    if (B < 1u) B = 1u;
    if (B > p.sample_count) B = p.sample_count;
    if ((uint64_t)B * slots_per_sample > 0x1fffffffull) B = (uint32_t)(0x1fffffffull / slots_per_sample);
    if (B < 1u) return cudaErrorInvalidValue;
    const uint32_t n_batches = (p.sample_count + B - 1u) / B;
    int lanes_used = n_batches < (uint32_t)kLanes ? (int)n_batches : kLanes;
    {   // RTB_WF_LANES_MAX=n caps the batches in flight (A/B measurements)
        static const int cap = [] { const char* s = std::getenv("RTB_WF_LANES_MAX"); return s && s[0] ? std::atoi(s) : 0; }();
        if (cap > 0 && lanes_used > cap) lanes_used = cap;
    }
    for (int l = 0; l < lanes_used; ++l) {
        e = lane_reserve(&st->lanes[l], (size_t)B * slots_per_sample);
        if (e == cudaErrorMemoryAllocation && l > 0) {  // not enough memory for every lane: pipeline over fewer
            (void)cudaGetLastError();
            lanes_used = l;
            break;
        }
        if (e != cudaSuccess) return e;
    }

    const bool smem_nodes = p.scene.n_nodes > 0 && wf_layout_bytes(p.scene, p.ordered) <= megakernel_max_smem_nodes_bytes();
    const bool quads = p.scene.has_quads != 0u;
    const uint32_t max_grid = (uint32_t)st->sm_count * 8u;
    // Perlin tables staged by wf_shade: up to 8 NoiseTextures (38 KB of dynamic shared memory, below the 48 KB that
    // needs no opt-in); RTB_PERLIN_SMEM=0 reads them from global memory instead (A/B measurements).
    static const bool perlin_smem_on = [] { const char* s = std::getenv("RTB_PERLIN_SMEM"); return !(s && s[0] == '0'); }();
    const uint32_t perlin_count = (perlin_smem_on && p.scene.n_perlins > 0u && p.scene.n_perlins <= 8u) ? p.scene.n_perlins : 0u;
    const size_t perlin_bytes = (size_t)perlin_count * sizeof(DevPerlin);

    static const char* timeline_path = std::getenv("RTB_TIMELINE");
    if (timeline_path && timeline_path[0] && !st->timeline) {
        e = cudaMalloc(&st->timeline, (size_t)kTimelineCap * sizeof(uint4));
        if (e == cudaSuccess) e = cudaMalloc(&st->timeline_count, sizeof(uint32_t));
        if (e != cudaSuccess) return e;
    }
    if (st->timeline) {
        e = cudaMemsetAsync(st->timeline_count, 0, sizeof(uint32_t), stream);
        if (e != cudaSuccess) return e;
    }

    // fork: the lanes start after everything already queued on the caller's stream
    e = cudaEventRecord(st->begin, stream);
    for (int l = 0; l < lanes_used && e == cudaSuccess; ++l) e = cudaStreamWaitEvent(st->lanes[l].stream, st->begin, 0);
    if (e != cudaSuccess) return e;

    uint32_t launches = 0;
    int prev_lane = -1;
    uint32_t tail_k = 0;
    if (st->tail_scene == p.scene.nodes && st->tail_depth == p.cam.max_depth && st->tail_slots == slots_per_sample)
        tail_k = st->tail_bounce;
    if (wf_forced_tail_bounce() == 0) tail_k = p.cam.max_depth + 1u;  // never switches
    else if (wf_forced_tail_bounce() > 0) tail_k = (uint32_t)wf_forced_tail_bounce();
    for (uint32_t batch = 0; batch < n_batches; ++batch) {
        const uint32_t s0 = batch * B;
        const uint32_t nb = (p.sample_count - s0 < B) ? p.sample_count - s0 : B;
        const uint32_t cap = nb * slots_per_sample;
        const int lane_id = (int)(batch % (uint32_t)lanes_used);
        WfLane& ln = st->lanes[lane_id];
        WfParams P{};
        P.R = p;
        P.R.tile_world = world;
        P.hitq = ln.hitq;
        P.TL = ln.TL;
        P.capacity = (uint32_t)ln.capacity;
        P.slots_per_sample = slots_per_sample;
        P.batch_begin = p.sample_begin + s0;
        P.batch_samples = nb;
        P.perlin_smem = perlin_count;
        P.timeline = st->timeline;
        P.timeline_count = st->timeline_count;
        P.timeline_cap = kTimelineCap;
        P.timeline_tag = (uint32_t)lane_id << 16;
        e = cudaMemsetAsync(ln.counts, 0, 4 * kBins * sizeof(uint32_t), ln.stream);
        if (e != cudaSuccess) return e;
        int cur = 0;
        P.out = ln.q[cur];
        P.count_out = ln.counts + cur * kBins;
        uint32_t grid = (cap + 255u) / 256u + kBins;  // chunks: every bin may end with a partial one
        if (grid > max_grid) grid = max_grid;
        uint32_t grid_e = (cap + kExtendThreads - 1u) / kExtendThreads + kBins;
        const uint32_t extend_per_sm = smem_nodes ? RTB_EXTEND_GRID_PER_SM : RTB_EXTEND_GRID_PER_SM_GLOBAL;
        if (grid_e > (uint32_t)st->sm_count * extend_per_sm) grid_e = (uint32_t)st->sm_count * extend_per_sm;
        uint32_t grid_s = (cap + 255u) / 256u + kBins;
        if (grid_s > (uint32_t)st->sm_count * RTB_SHADE_GRID_PER_SM) grid_s = (uint32_t)st->sm_count * RTB_SHADE_GRID_PER_SM;
        // Where does the thin tail begin?  Known from the last render of this scene, or learned now: batch 0 runs
        // every bounce as a kernel pair and reports how many rays entered each; when its lane comes up for reuse the
        // host waits for it (the other lanes keep the GPU busy meanwhile) and picks the switch bounce for the rest.
        if (tail_k == 0u && batch == (uint32_t)lanes_used) {
            e = cudaEventSynchronize(st->lanes[0].accumulated);
            if (e != cudaSuccess) return e;
            tail_k = choose_tail_bounce(st->lanes[0].survivors_host, p.cam.max_depth);
            st->tail_bounce = tail_k;
            st->tail_depth = p.cam.max_depth;
            st->tail_scene = p.scene.nodes;
            st->tail_slots = slots_per_sample;
        }
        P.survivors = (tail_k == 0u && batch == 0u) ? ln.survivors_dev : nullptr;
        if (P.survivors) std::memset(ln.survivors_host, 0xff, kSurvivorSlots * sizeof(uint32_t));
        wf_raygen<<<grid, 256, 0, ln.stream>>>(P);
        ++launches;
        for (uint32_t bounce = 0; bounce < p.cam.max_depth; ++bounce) {
            P.in = ln.q[cur];
            P.count_in = ln.counts + cur * kBins;
            P.timeline_tag = ((uint32_t)lane_id << 16) | ((bounce & 0xffu) << 24);
            if (tail_k != 0u && bounce >= tail_k) {  // everything still alive finishes in one launch
                P.segment = bounce + 1u;
                const uint32_t grid_t = (uint32_t)st->sm_count * 4u;
#define RTB_TAIL(C, Q)                                                                         \
    do {                                                                                       \
        if (p.ordered == 3u)      wf_tail<C, Q, kSlabPacked><<<grid_t, 256, 0, ln.stream>>>(P); \
        else if (p.ordered == 2u) wf_tail<C, Q, kSlabFma><<<grid_t, 256, 0, ln.stream>>>(P);    \
        else                      wf_tail<C, Q, kSlabExact><<<grid_t, 256, 0, ln.stream>>>(P);  \
    } while (0)
                if (count_work) { if (quads) RTB_TAIL(true, true); else RTB_TAIL(true, false); }
                else            { if (quads) RTB_TAIL(false, true); else RTB_TAIL(false, false); }
#undef RTB_TAIL
                e = cudaGetLastError();
                if (e != cudaSuccess) return e;
                ++launches;
                break;
            }
            P.out = ln.q[cur ^ 1];
            P.count_out = ln.counts + (cur ^ 1) * kBins;
            P.hit_count = ln.counts + (2u + (bounce & 1u)) * kBins;
            P.hit_count_next = ln.counts + (2u + ((bounce + 1u) & 1u)) * kBins;
            P.segment = bounce + 1u;
#define RTB_WF(C, Q)                                                      \
    do {                                                                  \
        e = wf_launch_extend<C, Q>(P, smem_nodes, grid_e, ln.stream);     \
        if (e == cudaSuccess) {                                           \
            wf_shade<C, Q><<<grid_s, 256, perlin_bytes, ln.stream>>>(P);  \
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
            if (bounce == 7u) {
                if (grid_s > (uint32_t)st->sm_count * 2u) grid_s = (uint32_t)st->sm_count * 2u;
                if (grid_e > (uint32_t)st->sm_count) grid_e = (uint32_t)st->sm_count;
            }
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
    if (st->timeline) {  // development aid: blocks, then writes the records of this render
        e = cudaStreamSynchronize(stream);
        uint32_t n = 0;
        if (e == cudaSuccess) e = cudaMemcpy(&n, st->timeline_count, sizeof(n), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return e;
        if (n > kTimelineCap) n = kTimelineCap;
        std::vector<uint4> rec(n);
        e = cudaMemcpy(rec.data(), st->timeline, (size_t)n * sizeof(uint4), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return e;
        if (FILE* f = std::fopen(timeline_path, "wb")) {
            std::fwrite(rec.data(), sizeof(uint4), n, f);
            std::fclose(f);
        }
    }
    return cudaSuccess;
}

}  // namespace rtb
