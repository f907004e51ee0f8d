"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous and the single exchange step.

The reference's only parallelism is 8 threads over contiguous pixel strips (src/main.zig:318-324,
src/camera.zig:93-99).  Here the scene is replicated on every GPU and the work is partitioned either by
  * SAMPLE ranges ("samples"): rank r renders global samples [r*spp_r, (r+1)*spp_r) of the whole frame —
    Philox is keyed by the global sample index, so the union of the ranks' paths is exactly the set a
    single GPU would trace; perfectly balanced; or
  * interleaved 32x8-pixel TILES ("tiles"): rank r renders tiles t with t % world == r — disjoint pixels.
Either way every rank ends up with a float4 accumulation buffer of the full frame (zero where it rendered
nothing) and ONE exchange step combines them BEFORE gamma and quantisation:
  * `PeerExchange` (the product path on a multi-GPU node): the ranks map each other's buffers (CUDA IPC over
    NVLink/NVSwitch) and each runs ONE kernel (rtb_exchange_resolve) that sums its slice of the frame over all
    ranks, resolves it and stores sums + RGBA8 straight into rank 0's buffers — reduce-scatter + resolve + gather
    fused; torch.distributed only carries the 64-byte handles and the barrier around the kernel;
  * `combine` (NCCL sum-reduce to rank 0, then rtb_resolve_device there; gloo on CPU in the tests): the baseline the
    fused kernel is measured against, and the fallback when peer mapping is unavailable.
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


def split_samples(total: int, world: int, weights=None):
    """Split `total` samples over `world` ranks: equal shares, or proportional to `weights` (e.g. the rate every rank
    showed in warm-up renders — a slower GPU then gets fewer samples instead of making the others wait at the
    exchange).  Largest-remainder rounding, at least one sample per rank; deterministic, so every rank computes the
    same table from the same (all-gathered) weights.  Returns the per-rank counts."""
    if total < world:
        raise ValueError(f"{total} samples cannot be shared by {world} ranks")
    if weights is None:
        base, rem = divmod(total, world)
        return [base + (1 if r < rem else 0) for r in range(world)]
    if len(weights) != world or any(not (w > 0.0) for w in weights):
        raise ValueError("weights must be one positive number per rank")
    s = float(sum(weights))
    ideal = [total * w / s for w in weights]
    counts = [max(1, int(x)) for x in ideal]
    # hand out / take back the difference by largest (smallest) remainder
    order = sorted(range(world), key=lambda r: ideal[r] - int(ideal[r]), reverse=True)
    k = 0
    while sum(counts) < total:
        counts[order[k % world]] += 1
        k += 1
    k = 0
    while sum(counts) > total:
        r = order[-1 - (k % world)]
        if counts[r] > 1:
            counts[r] -= 1
        k += 1
    return counts


def plan(mode: str, rank: int, world: int, samples: int, sample_base: int = 0, weak: bool = False,
         weights=None) -> Partition:
    """Partition `samples` samples per pixel over `world` ranks.

    weak=True (sample mode only): the combined frame holds world*samples per pixel (fixed work per GPU on average —
    the bench's weak-scaling workload); otherwise the ranks share `samples`.  weights (sample mode only): per-rank
    rates for a proportional split (see split_samples); None = equal shares.
    """
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if mode == "samples":
        total = world * samples if weak else samples
        if weak and weights is None:
            return Partition(mode, rank, world, sample_base + rank * samples, samples, 0, 1, total)
        if total < world:   # fewer samples than ranks: the trailing ranks get none (callers skip their render)
            counts = [1 if r < total else 0 for r in range(world)]
        else:
            counts = split_samples(total, world, weights)
        begin = sum(counts[:rank])
        return Partition(mode, rank, world, sample_base + begin, counts[rank], 0, 1, total)
    if mode == "tiles":
        if weak:
            raise ValueError("weak scaling is defined for the sample partition only")
        return Partition(mode, rank, world, sample_base, samples, rank, world, samples)
    raise ValueError(f"unknown partition mode {mode!r}")


def apply(part: Partition, options):
    """Write a Partition into an RtbRenderOptions."""
    if part.sample_count == 0:   # RtbRenderOptions.sample_count == 0 means "the camera's spp", not "nothing"
        raise ValueError("this rank has no samples to render (more ranks than samples): skip its render call")
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


class _DeviceArray:
    """Zero-copy view descriptor (``__cuda_array_interface__``) of a library-owned device buffer."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerExchange:
    """Per-rank accumulation buffers mapped into every rank + the fused exchange/resolve kernel.

    Mirrors what the reference gets for free from shared memory: its 8 threads write strips of ONE
    `SharedStateImageWriter` buffer (src/camera.zig:22-27, src/main.zig:318-323).  Here every rank owns a full-frame
    float4 buffer allocated by the library (cudaMalloc, so it can be exported), all ranks map all buffers, and after
    the render each rank combines + resolves its slice of the frame into rank `root`'s buffers in one kernel.

    group: a torch.distributed process group (None = default).  With the NCCL backend the barrier is a 1-element
    all-reduce on the current CUDA stream (stream-ordered, no host synchronisation); with gloo (tests) it is a
    device synchronise + host barrier.
    """

    def __init__(self, n_pixels: int, rank: int, world: int, device: int, group=None, root: int = 0):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _ffi, _check

        self._ffi, self._check, self._C = _ffi, _check, C
        self.n_pixels, self.rank, self.world, self.device, self.group, self.root = n_pixels, rank, world, device, group, root
        lib = _ffi.rtb()
        self._own = []
        self._opened = []
        self.accum = self.rgba = None

        def alloc(nbytes):
            p = C.c_void_p()
            _check(lib.rtb_buffer_alloc(device, nbytes, C.byref(p)), "rtb_buffer_alloc")
            self._own.append(p.value)
            return p.value

        def agree(ok, what, error):
            """Every rank learns whether every rank succeeded, so that a local failure (out of memory, no IPC
            permission) makes ALL ranks raise instead of leaving the others waiting in the next collective."""
            if world == 1:
                flags = [(ok, error)]
            else:
                flags = [None] * world
                dist.all_gather_object(flags, (ok, error), group=group)
            bad = [(r, e) for r, (k, e) in enumerate(flags) if not k]
            if bad:
                self._release()
                raise RuntimeError(f"PeerExchange unavailable ({what}): " + "; ".join(f"rank {r}: {e}" for r, e in bad))

        # 1. local buffers + their export handles
        mine, error = [], None
        try:
            self.accum_ptr = alloc(n_pixels * 16)
            self.rgba_ptr = alloc(n_pixels * 4)
            self.accum = torch.as_tensor(_DeviceArray(self.accum_ptr, (n_pixels, 4), "<f4"), device=f"cuda:{device}")
            self.rgba = torch.as_tensor(_DeviceArray(self.rgba_ptr, (n_pixels, 4), "|u1"), device=f"cuda:{device}")
            self.accum.zero_()
            self.rgba.zero_()
            if world > 1:
                for ptr in (self.accum_ptr, self.rgba_ptr):
                    h = _ffi.RtbIpcHandle()
                    _check(lib.rtb_ipc_export(device, ptr, C.byref(h)), "rtb_ipc_export")
                    mine.append(bytes(h.bytes))
        except Exception as e:  # noqa: BLE001 - reported to every rank below
            error = str(e)
        agree(error is None, "allocating / exporting the buffers", error)
        self.peer_accum = [None] * world
        self.peer_accum[rank] = self.accum_ptr
        self.root_accum, self.root_rgba = self.accum_ptr, self.rgba_ptr
        self._nccl = world > 1 and dist.get_backend(group) == "nccl"
        self._flag = torch.zeros(1, device=f"cuda:{device}") if self._nccl else None
        # 2. map everybody else's buffers
        if world > 1:
            everyone = [None] * world
            dist.all_gather_object(everyone, mine, group=group)

            def open_(raw):
                h = _ffi.RtbIpcHandle()
                C.memmove(h.bytes, raw, 64)
                p = C.c_void_p()
                _check(lib.rtb_ipc_open(device, C.byref(h), C.byref(p)), "rtb_ipc_open")
                self._opened.append(p.value)
                return p.value

            error = None
            try:
                for r in range(world):
                    if r != rank:
                        self.peer_accum[r] = open_(everyone[r][0])
                if rank != root:
                    self.root_accum = self.peer_accum[root]
                    self.root_rgba = open_(everyone[root][1])
            except Exception as e:  # noqa: BLE001
                error = str(e)
            agree(error is None, "mapping the peers' buffers", error)
            self.barrier()
        self._peers = (C.c_void_p * world)(*self.peer_accum)

    def _release(self):
        lib = self._ffi.rtb()
        self.accum = self.rgba = None
        for p in self._opened:
            lib.rtb_ipc_close(self.device, p)
        self._opened = []
        for p in self._own:
            lib.rtb_buffer_free(self.device, p)
        self._own = []

    def slice(self):
        C = self._C
        b, e = C.c_uint64(), C.c_uint64()
        self._check(self._ffi.rtb().rtb_exchange_slice(self.n_pixels, self.world, self.rank, self.root, C.byref(b),
                                                       C.byref(e)),
                    "rtb_exchange_slice")
        return b.value, e.value

    def barrier(self, stream: int | None = None):
        """Orders the CUDA work of all ranks: nothing queued after it on `stream` starts before everything queued
        before it, on every rank, has finished.  `stream` = raw cudaStream_t of the stream the exchange kernel runs on
        (None = torch's current stream); the NCCL all-reduce is issued ON that stream, so it orders the kernel."""
        if self.world == 1:
            return
        import torch
        import torch.distributed as dist
        if self._nccl:
            if stream is None or stream == torch.cuda.current_stream(self.device).cuda_stream:
                dist.all_reduce(self._flag, group=self.group)
            else:
                with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=self.device)):
                    dist.all_reduce(self._flag, group=self.group)
        else:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)

    def exchange(self, samples_per_pixel: float, stream: int | None = None, n_pixels: int | None = None):
        """barrier -> rtb_exchange_resolve on every rank -> barrier, all on `stream` (raw cudaStream_t; None = torch's
        current stream).  Afterwards rank `root`'s `accum` holds the combined sums (.w = samples_per_pixel) and its
        `rgba` the resolved frame.  NOTE: root's `accum` then holds the COMBINED sums — zero every rank's `accum`
        before rendering the next frame into it."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        self.barrier(stream)
        n = self.n_pixels if n_pixels is None else int(n_pixels)   # a frame smaller than the buffers: their first n pixels
        if not (0 < n <= self.n_pixels):
            raise ValueError("n_pixels out of range")
        self._check(self._ffi.rtb().rtb_exchange_resolve(self._peers, self.world, self.rank, self.root, self.root_accum,
                                                         self.root_rgba, n, float(samples_per_pixel),
                                                         self.device, stream), "rtb_exchange_resolve")
        self.barrier(stream)

    def close(self):
        if self.world > 1:
            try:
                self.barrier()
            except Exception:  # noqa: BLE001 - closing anyway
                pass
        self._release()
