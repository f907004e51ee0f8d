// rtb_wavefront.cuh — K2: wavefront integrator (raygen / extend / shade+compact / accumulate).
#pragma once

#include "rtb_kernels.cuh"

namespace rtb {

struct WavefrontState;  // device queues, owned by an RtbScene, grown on demand

WavefrontState* wavefront_create();
void wavefront_destroy(WavefrontState* st);

// Renders samples [p.sample_begin, +p.sample_count) of the pixels selected by p into p.accum,
// bit-identically to the megakernel (same per-path arithmetic, per-pixel sums in sample order).
cudaError_t wavefront_render(WavefrontState* st, const RenderParams& p, bool count_work, cudaStream_t stream,
                             LaunchInfo* info);

}  // namespace rtb
