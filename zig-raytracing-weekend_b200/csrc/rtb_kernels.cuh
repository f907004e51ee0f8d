// rtb_kernels.cuh — kernel parameter blocks and host-callable launchers (implemented in
// rtb_kernels.cu, compiled for sm_100a with -fmad=false).
#pragma once

#include "rtb_device.cuh"

namespace rtb {

// Pixel tiling: one CTA of 256 threads owns a 32 x 8 pixel tile; each warp an 8 x 4 sub-tile (so
// the 32 primary rays of a warp are spatially coherent and the float4 accumulator rows coalesce).
constexpr uint32_t kTileW = 32, kTileH = 8, kCtaThreads = 256;

struct RenderParams {
    DevScene scene;
    DevCamera cam;
    float4* accum;  // W*H, (sum r, sum g, sum b, n)
    uint2 seed;
    uint32_t sample_begin, sample_count;
    uint32_t pixel_begin, pixel_end;  // flat pixel range [begin, end)
    uint32_t tile_rank, tile_world;
    uint32_t ordered;  // layout index = RTB_TRAVERSAL_* (0 reference order, 1 ordered, 2 SAH re-partition)
    unsigned long long* counters;  // [rays, box tests, object tests, hits] or nullptr
};

struct LaunchInfo {
    uint32_t n_launches = 0;
};

// K1: megakernel path tracer.  nodes_in_smem selects the variant that stages the node array in
// shared memory (small scenes); count_work fills params.counters.
cudaError_t launch_megakernel(const RenderParams& p, bool nodes_in_smem, bool count_work, cudaStream_t stream,
                              LaunchInfo* info);

// K3: nearest-hit query for a batch of rays (parity harness).
cudaError_t launch_trace(const DevScene& scene, const RtbRay* d_rays, uint64_t n, RtbHit* d_hits, uint32_t layout,
                         cudaStream_t stream);

// K4: resolve (toGamma2 + truncation).
cudaError_t launch_resolve(const float4* d_accum, uchar4* d_rgba, uint64_t n_pixels, float n_override,
                           cudaStream_t stream);

// K5: fused multi-GPU exchange + resolve over peer memory (one launch per rank, its slice [begin, end)).
constexpr uint32_t kMaxPeers = 16;
struct PeerAccums {
    const float4* p[kMaxPeers];  // accumulators of rank 0..world-1 as mapped into THIS process
};
cudaError_t launch_exchange_resolve(const PeerAccums& peers, uint32_t world, float4* root_accum, uchar4* root_rgba,
                                    uint64_t begin, uint64_t end, float samples_per_pixel, cudaStream_t stream);

cudaError_t launch_philox_selftest(const uint4* d_ctr, uint2 key, uint32_t n, uint4* d_out, cudaStream_t stream);

// FFMA-chain microbenchmark (roofline denominator): out must hold grid*256 floats.
cudaError_t launch_ffma_peak(float* d_out, uint32_t grid, uint32_t iters, cudaStream_t stream);

// Number of SMs of the current device.
uint32_t current_sm_count();

// Largest dynamic shared memory the megakernel may use for the node array.
size_t megakernel_max_smem_nodes_bytes();

}  // namespace rtb
