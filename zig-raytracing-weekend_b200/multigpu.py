"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous and the single exchange step.

The reference's only parallelism is 8 threads over contiguous pixel strips (src/main.zig:318-324,
src/camera.zig:93-99).  Here the scene is replicated on every GPU and the work is partitioned either by
  * SAMPLE ranges ("samples"): rank r renders global samples [r*spp_r, (r+1)*spp_r) of the whole frame —
    Philox is keyed by the global sample index, so the union of the ranks' paths is exactly the set a
    single GPU would trace; perfectly balanced; or
  * interleaved 32x8-pixel TILES ("tiles"): rank r renders tiles t with t % world == r — disjoint pixels.
Either way every rank ends up with a float4 accumulation buffer of the full frame (zero where it rendered
nothing) and ONE collective combines them: a sum-reduce to rank 0 over NCCL/NVLink (gloo on CPU in tests),
BEFORE gamma and quantisation, which then run once on rank 0 (rtb_resolve_device).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Partition:
    mode: str          # "samples" | "tiles"
    rank: int
    world: int
    sample_begin: int  # global index of this rank's first sample
    sample_count: int  # samples this rank renders
    tile_rank: int
    tile_world: int
    total_samples: int  # samples per pixel in the combined frame


def plan(mode: str, rank: int, world: int, samples: int, sample_base: int = 0, weak: bool = False) -> Partition:
    """Partition `samples` samples per pixel over `world` ranks.

    weak=True (sample mode only): every rank renders `samples` samples, so the combined frame holds
    world*samples per pixel (fixed work per GPU — the bench's weak-scaling workload).
    """
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if mode == "samples":
        if weak:
            return Partition(mode, rank, world, sample_base + rank * samples, samples, 0, 1, world * samples)
        base, rem = divmod(samples, world)
        count = base + (1 if rank < rem else 0)
        begin = rank * base + min(rank, rem)
        return Partition(mode, rank, world, sample_base + begin, count, 0, 1, samples)
    if mode == "tiles":
        if weak:
            raise ValueError("weak scaling is defined for the sample partition only")
        return Partition(mode, rank, world, sample_base, samples, rank, world, samples)
    raise ValueError(f"unknown partition mode {mode!r}")


def apply(part: Partition, options):
    """Write a Partition into an RtbRenderOptions."""
    options.sample_begin = part.sample_begin
    options.sample_count = part.sample_count
    options.tile_rank = part.tile_rank
    options.tile_world = part.tile_world
    return options


def combine(accum, part: Partition, group=None, dst: int = 0, fix_w: bool = True):
    """The exchange step: sum-reduce the per-rank float4 accumulators (torch tensor [N,4]) onto `dst`.

    After the reduce the .w column of `dst` holds the sum of the ranks' sample counters; it is
    overwritten with the frame's true samples-per-pixel (north_star: ".w = spp"), so that the resolve
    divides by the right n.  Returns the tensor (meaningful on `dst` only).
    """
    import torch.distributed as dist

    if part.world > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if fix_w and part.rank == dst:
        accum[:, 3] = float(part.total_samples)
    return accum
