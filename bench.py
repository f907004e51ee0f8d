#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 path-tracing hot path (contract: see the task brief / DESIGN.md §6).

  python bench.py [--gpus N] [--steps K] [--warmup W]             our arm (CUDA through the C ABI)
  python bench.py --impl reference [--gpus N] [--steps K] ...      the reference's CPU algorithm (oracle port)

Metric (BASELINE.json): Mpaths/s on the Book-1 final scene.  A *step* is one pass of the hot path over one
batch: `spp` samples per pixel of the 1200x675 Book-1 frame (BASELINE config 2: 500 spp, depth 50 = 405 M paths),
followed by the exchange step (NCCL sum-reduce to rank 0 when N > 1) and the resolve kernel (gamma + RGBA8).
N > 1 is WEAK scaling: every rank renders its own range of `spp` samples of the same frame (sample-partitioned,
Philox keyed by the global sample index), so the combined frame holds N*spp samples per pixel.

  value  : device-resident — accumulators live in HBM, timed with CUDA events on the launch stream, max over ranks.
  e2e    : the same work through the reference-facing call with HOST buffers (rtb_render: H2D of the float4
           accumulation buffer, render, resolve, D2H of accum + RGBA8), pinned host memory, wall clock.
  roofline: binding ceiling for this cache-resident scene is the FP32 pipe; `achieved` = algorithmic
           flops of the REFERENCE traversal (counted by the kernel's counting build on the same workload) / kernel time.
  cpu_baseline: the oracle (C++ restatement of the reference, "port") on the box's host cores, bounded sample.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WIDTH, SPP, DEPTH = 1200, 500, 50  # BASELINE.json configs[1]
SEED = 1234


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--integrator", default="auto", choices=["auto", "megakernel", "wavefront"])
    ap.add_argument("--traversal", default="sah", choices=["reference", "ordered", "sah"],
                    help="reference = the reference's left-then-right order over the host's tree (bit-exact hit index); "
                         "ordered = same tree, near child first; sah = the library's SAH re-partition of the same objects")
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--spp", type=int, default=SPP)
    ap.add_argument("--depth", type=int, default=DEPTH)
    ap.add_argument("--partition", default="samples", choices=["samples", "tiles"])
    ap.add_argument("--no-variants", action="store_true",
                    help="skip the six integrator/traversal variant timings (keeps an ncu launch list to the headline path)")
    ap.add_argument("--exchange", choices=["auto", "p2p", "nccl"], default="auto",
                    help="N>1 exchange step: p2p = fused reduce+resolve kernel over peer memory (rtb_exchange_resolve), "
                         "nccl = NCCL sum-reduce to rank 0 then rtb_resolve_device, auto = time both on this box and "
                         "take the faster (measured: p2p at N=2, NCCL's in-switch reduction at N>=4)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank renders --spp samples (sample partition only); strong: the ranks share "
                         "--spp samples (samples) or the frame's tiles (tiles) — e.g. BASELINE config 5: "
                         "--width 7680 --spp 1024 --scaling strong")
    ap.add_argument("--cpu-spp", type=int, default=0, help="samples per pixel of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- roofline accounting
def algorithmic_work(st, width, height):
    """flops and bytes of the REFERENCE traversal, from the kernel's own work counters."""
    n_moving_frac = 0.8  # book1: ~80 % of the leaves are moving spheres (32 B instead of 16 B per test)
    flops = 15 * st["n_box_tests"] + 23 * st["n_object_tests"] + (27 + 40) * st["n_hits"] + 30 * st["n_paths"] \
        + 6 * st["n_rays"]
    bytes_ = 32 * st["n_box_tests"] + (16 + 16 * n_moving_frac) * st["n_object_tests"] + 16 * st["n_hits"] \
        + 20 * width * height
    return flops, bytes_


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback"


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_render_rate(p, orc, world, cam, spp, threads, seed=4321):
    o = p.render_options(seed=seed, sample_count=spp)
    _, _, st = orc.render(world.desc, cam, o, n_threads=threads, want_rgba=True)
    return st["n_paths"] / (st["device_ms"] * 1e-3) / 1e6, st


def run_reference(a):
    """--impl reference: the reference's CPU algorithm (oracle port; the Zig binary cannot be built here), all host
    threads, bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    p = importlib.import_module("zig-raytracing-weekend_b200")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    import oracle_ffi as orc
    world = p.World.book1()
    cam = p.book1_camera(a.width, a.spp, a.depth).init()
    threads = os.cpu_count() or 1
    probe, _ = cpu_render_rate(p, orc, world, cam, 1, threads)
    npx = cam.image_width * cam.image_height
    # bounded sample: ~8 s of CPU work per step
    spp = a.cpu_spp or max(1, min(a.spp, int(probe * 1e6 * 8.0 / npx)))
    for _ in range(a.warmup):
        cpu_render_rate(p, orc, world, cam, 1, threads)
    t0 = time.perf_counter()
    paths = 0
    for k in range(a.steps):
        _, st = cpu_render_rate(p, orc, world, cam, spp, threads, seed=4321 + k)
        paths += st["n_paths"]
    dt = time.perf_counter() - t0
    value = paths / dt / 1e6
    sample = f"{cam.image_width}x{cam.image_height} x {spp} spp per step ({paths // a.steps} paths), depth {a.depth}"
    print(json.dumps({
        "impl": "reference", "metric": "Mpaths/s (Book-1 final scene)", "value": value, "unit": "Mpaths/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, cam, "cpu"),
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


def workload_config(a, cam, integrator):
    return {"workload": f"book1 (generateWorld, ~485 spheres + BVH) {cam.image_width}x{cam.image_height}, "
                        f"{a.spp} spp {'per GPU' if a.scaling == 'weak' and a.partition == 'samples' else 'in total'}, "
                        f"depth {a.depth}" + (" [BASELINE configs[1]]" if (a.width, a.spp) == (WIDTH, SPP) else ""),
            "scene_seed": 1, "bvh_seed": 2, "render_seed": SEED, "integrator": integrator, "traversal": a.traversal,
            "partition": a.partition if a.gpus > 1 else "none", "background": "sky gradient (camera.zig:204-206)",
            "l2": "flushed between steps (256 MiB memset); scene is 47 KB and cache/smem resident by nature"}


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    p = importlib.import_module("zig-raytracing-weekend_b200")
    mg = importlib.import_module("zig-raytracing-weekend_b200.multigpu")

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size != a.gpus:
        if world_size == 1 and a.gpus > 1:
            raise SystemExit(f"--gpus {a.gpus} needs torchrun (python -m torch.distributed.run --nproc-per-node {a.gpus} ...)")
        raise SystemExit(f"WORLD_SIZE={world_size} but --gpus {a.gpus}")
    if p.device_count() <= local_rank:
        raise SystemExit("no CUDA device for this rank; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)

    world = p.World.book1()
    camo = p.book1_camera(a.width, a.spp, a.depth)
    cam = camo.init()
    W, H = cam.image_width, cam.image_height
    npx = W * H
    scene = p.Scene(world, device=local_rank)
    stream = torch.cuda.current_stream(dev)
    sptr = stream.cuda_stream

    px = None
    if world_size > 1 and a.exchange != "nccl":
        # per-rank buffers owned by the library and mapped into every rank (CUDA IPC over NVLink/NVSwitch)
        try:
            px = mg.PeerExchange(npx, rank, world_size, local_rank)   # raises on EVERY rank if any rank cannot map
        except RuntimeError as e:
            if a.exchange == "p2p":
                raise
            print(f"[bench] peer-memory exchange unavailable, using NCCL: {e}", file=sys.stderr)
            px = None
    if px is not None:
        d_acc, d_rgba = px.accum, px.rgba
    else:
        d_acc = torch.zeros(npx, 4, device=dev, dtype=torch.float32)
        d_rgba = torch.zeros(npx, 4, device=dev, dtype=torch.uint8)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    exchange_kind = "p2p" if px is not None else "nccl"

    def exchange_and_resolve(part, kind=None):
        """The exchange step + resolve: afterwards rank 0 holds the combined sums and the RGBA8 frame."""
        if (kind or exchange_kind) == "p2p":
            px.exchange(float(part.total_samples), sptr)
            return
        mg.combine(d_acc, part, fix_w=False)
        if rank == 0:
            p._check(p._ffi.rtb().rtb_resolve_device(d_acc.data_ptr(), d_rgba.data_ptr(), npx,
                                                     float(part.total_samples), local_rank, sptr), "rtb_resolve_device")

    exchange_probe = None
    if world_size > 1 and a.exchange == "auto" and px is not None:   # measured, not assumed: both exchanges on this frame size, this box
        part0 = mg.plan(a.partition, rank, world_size, a.spp, weak=(a.partition == "samples" and a.scaling == "weak"))
        times = {}
        for kind in ("p2p", "nccl"):
            for _ in range(2):
                exchange_and_resolve(part0, kind)
            torch.cuda.synchronize(dev)
            dist.barrier()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record(stream)
            for _ in range(5):
                exchange_and_resolve(part0, kind)
            x1.record(stream)
            torch.cuda.synchronize(dev)
            tk = torch.tensor([x0.elapsed_time(x1) / 5], device=dev, dtype=torch.float64)
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
            times[kind] = tk.item()
        exchange_kind = "p2p" if times["p2p"] <= times["nccl"] else "nccl"
        exchange_probe = {"p2p_ms": times["p2p"], "nccl_ms": times["nccl"]}

    traversal = {"reference": p.RTB_TRAVERSAL_REFERENCE, "ordered": p.RTB_TRAVERSAL_ORDERED,
                 "sah": p.RTB_TRAVERSAL_SAH}[a.traversal]

    def make_options(step, integrator, flags=0, spp=None, trav=None):
        part = mg.plan(a.partition, rank, world_size, spp or a.spp, sample_base=0,
                       weak=(a.partition == "samples" and a.scaling == "weak"))
        o = p.render_options(seed=SEED + step, integrator=integrator, flags=flags,
                             traversal=traversal if trav is None else trav)
        return mg.apply(part, o), part

    def step_device(step, integrator, count=False, spp=None, trav=None):
        """One device-resident step: clear, render, (reduce), resolve.  Returns this rank's render stats."""
        o, part = make_options(step, integrator, p.RTB_FLAG_COUNT_WORK if count else 0, spp, trav)
        d_acc.zero_()
        st = scene.render_device(cam, o, d_acc.data_ptr(), sptr, want_stats=count)
        exchange_and_resolve(part)
        return st

    # -- integrator choice: measured, not assumed (north_star: pick from evidence) -------------------------------
    def time_once(integrator, spp):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        step_device(0, integrator, spp=spp)
        torch.cuda.synchronize(dev)
        e0.record(stream)
        step_device(0, integrator, spp=spp)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1)

    if a.integrator == "auto":
        probe_spp = max(1, min(a.spp, 32))
        t_mega = time_once(p.RTB_INTEGRATOR_MEGAKERNEL, probe_spp)
        t_wave = time_once(p.RTB_INTEGRATOR_WAVEFRONT, probe_spp)
        choice = torch.tensor([1.0 if t_wave < t_mega else 0.0], device=dev)
        if world_size > 1:
            dist.broadcast(choice, 0)
        integrator = p.RTB_INTEGRATOR_WAVEFRONT if choice.item() > 0.5 else p.RTB_INTEGRATOR_MEGAKERNEL
        probe = {"megakernel_ms": t_mega, "wavefront_ms": t_wave, "probe_spp": probe_spp}
    else:
        integrator = p.RTB_INTEGRATOR_WAVEFRONT if a.integrator == "wavefront" else p.RTB_INTEGRATOR_MEGAKERNEL
        probe = None
    integ_name = "wavefront" if integrator == p.RTB_INTEGRATOR_WAVEFRONT else "megakernel"

    # -- algorithmic work per path (counting build, untimed, reduced spp: the ratios are spp-independent) -------
    # `cst` = counters of the REFERENCE traversal (the algorithmic work of the task, SURVEY §8d: "a smarter traversal
    # that visits fewer nodes still gets credit for the reference's counts ... report both"); `cst_act` = counters of
    # the traversal actually timed.
    count_spp = max(1, min(a.spp, 16))
    cst = step_device(0, integrator, count=True, spp=count_spp, trav=p.RTB_TRAVERSAL_REFERENCE)
    cst_act = step_device(0, integrator, count=True, spp=count_spp)
    flops, bytes_ = algorithmic_work(cst, W, H)
    flops_per_path = flops / cst["n_paths"]
    bytes_per_path = (bytes_ - 20 * npx) / cst["n_paths"]
    flops_act, _ = algorithmic_work(cst_act, W, H)
    flops_act_per_path = flops_act / cst_act["n_paths"]

    # -- warm-up ----------------------------------------------------------------------------------------------
    for w in range(a.warmup):
        step_device(w, integrator)
    torch.cuda.synchronize(dev)

    # -- timed region: EXACTLY K steps, barrier + synchronize on both sides, CUDA events on the launch stream --
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(a.steps)]
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e_begin.record(stream)
    for k in range(a.steps):
        flush.zero_()  # L2 flush between timed iterations
        o, part = make_options(k, integrator)
        d_acc.zero_()
        ev[k][0].record(stream)
        scene.render_device(cam, o, d_acc.data_ptr(), sptr, want_stats=False)
        ev[k][1].record(stream)
        exchange_and_resolve(part)
        ev[k][2].record(stream)
    e_end.record(stream)
    torch.cuda.synchronize(dev)
    if world_size > 1:
        dist.barrier()
    total_ms = e_begin.elapsed_time(e_end)
    kernel_ms = sum(b.elapsed_time(e) for b, e, _ in ev) / a.steps
    exchange_ms = sum(e.elapsed_time(x) for _, e, x in ev) / a.steps   # includes waiting for the slowest rank
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([total_ms, kernel_ms, -exchange_ms], device=dev, dtype=torch.float64)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms, exchange_ms = t.tolist()
    exchange_ms = -exchange_ms   # MIN over ranks: the rank that arrives last sees the exchange without the wait

    weak = a.partition == "samples" and a.scaling == "weak"
    paths_per_step = npx * a.spp * (world_size if weak else 1)
    value = paths_per_step * a.steps / (total_ms * 1e-3) / 1e6
    # launches per step: one counting-free render reports them
    st1 = scene.render_device(cam, make_options(0, integrator, spp=a.spp)[0], d_acc.data_ptr(), sptr, want_stats=True)
    launches = (st1["n_launches"] + (1 if rank == 0 else 0)) * a.steps

    # -- the other variants, one untimed-for-the-headline step each (evidence for the choice; device-resident) ----
    variants = {}
    for vname, vint, vtrav in () if a.no_variants else (("megakernel/reference", p.RTB_INTEGRATOR_MEGAKERNEL, p.RTB_TRAVERSAL_REFERENCE),
                               ("megakernel/ordered", p.RTB_INTEGRATOR_MEGAKERNEL, p.RTB_TRAVERSAL_ORDERED),
                               ("megakernel/sah", p.RTB_INTEGRATOR_MEGAKERNEL, p.RTB_TRAVERSAL_SAH),
                               ("wavefront/reference", p.RTB_INTEGRATOR_WAVEFRONT, p.RTB_TRAVERSAL_REFERENCE),
                               ("wavefront/ordered", p.RTB_INTEGRATOR_WAVEFRONT, p.RTB_TRAVERSAL_ORDERED),
                               ("wavefront/sah", p.RTB_INTEGRATOR_WAVEFRONT, p.RTB_TRAVERSAL_SAH)):
        vspp = max(1, min(a.spp, 100))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_device(0, vint, spp=vspp, trav=vtrav)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        variants[vname] = npx * vspp * (world_size if weak else 1) / (e0.elapsed_time(e1) * 1e-3) / 1e6

    # -- e2e: reference-facing call with HOST buffers (pinned), copies inside the timed region ------------------
    e2e = None
    if not a.no_e2e:
        h_acc = torch.zeros(npx, 4, dtype=torch.float32).pin_memory()
        h_rgba = torch.zeros(npx, 4, dtype=torch.uint8).pin_memory()
        acc_np, rgba_np = h_acc.numpy(), h_rgba.numpy()
        st = p.RtbRenderStats()
        import ctypes as C

        def step_e2e(step):
            o, part = make_options(step, integrator)
            if world_size == 1:
                acc_np[:] = 0.0
                p._check(p._ffi.rtb().rtb_render(scene._h, C.byref(cam), C.byref(o), acc_np.ctypes.data,
                                                 rgba_np.ctypes.data, C.byref(st)), "rtb_render")
            else:
                h_acc.zero_()
                d_acc.copy_(h_acc, non_blocking=True)
                scene.render_device(cam, o, d_acc.data_ptr(), sptr, want_stats=False)
                exchange_and_resolve(part)
                if rank == 0:
                    h_acc.copy_(d_acc, non_blocking=True)
                    h_rgba.copy_(d_rgba, non_blocking=True)
                torch.cuda.synchronize(dev)

        step_e2e(0)
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for k in range(a.steps):
            step_e2e(k)
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world_size > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": paths_per_step * a.steps / dt.item() / 1e6, "unit": "Mpaths/s",
               "h2d_bytes_per_step": 16 * npx, "d2h_bytes_per_step": 20 * npx,
               "api": "rtb_render (host buffers)" if world_size == 1 else
                      ("H2D + rtb_render_device + rtb_exchange_resolve (peer memory) + D2H" if exchange_kind == "p2p" else
                       "H2D + rtb_render_device + NCCL reduce + rtb_resolve_device + D2H")}

    if px is not None:
        d_acc = d_rgba = None
        px.close()
    if rank != 0:
        if world_size > 1:
            dist.destroy_process_group()
        return 0

    # -- roofline (dominant kernel = the render kernel(s) of the step) -------------------------------------------
    peaks, peak_src = load_peaks()
    fp32_peak = p.measure_fp32_peak(local_rank)
    paths_per_launch_group = npx * a.spp // (1 if weak else world_size)  # per rank
    achieved_tflops = flops_per_path * paths_per_launch_group / (kernel_ms * 1e-3) / 1e12
    achieved_gbs = (bytes_per_path * paths_per_launch_group + 20 * npx) / (kernel_ms * 1e-3) / 1e9
    roofline = {
        "bound": "fp32", "kernel": "render_megakernel" if integ_name == "megakernel" else "wf_raygen+wf_extend+wf_shade+wf_tail+wf_accumulate",
        "achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved_tflops / fp32_peak,
        "peak_source": "FFMA microbenchmark measured in this run (FMA = 2 flops); the kernels run unfused, ceiling = peak/2",
        # DRAM bytes of one step, from the ncu capture in profiles/r1g_dram_summary.csv (dram__bytes_read+write summed
        # over every wavefront kernel of a render = 692.4 B per path on this workload; the megakernel moves ~16 B per
        # path): queue/state traffic, not the "algorithmic bytes" (node/sphere fetches), which shared memory serves.
        "traffic": (692.4 if integ_name == "wavefront" else 16.0) * paths_per_launch_group,
        "traffic_source": "profiles/r1g_dram_summary.csv (ncu, per path, scaled to the step)",
        "algorithmic_flops_per_path": flops_per_path, "kernel_ms_per_step": kernel_ms,
        "counting": "achieved = flops of the REFERENCE traversal (left-then-right over the host's tree, SURVEY 8d) "
                    "/ kernel time; achieved_actual = flops of the traversal that ran",
        "achieved_actual": flops_act_per_path * paths_per_launch_group / (kernel_ms * 1e-3) / 1e12,
        "frac_actual": flops_act_per_path * paths_per_launch_group / (kernel_ms * 1e-3) / 1e12 / fp32_peak,
        "actual_flops_per_path": flops_act_per_path,
        "actual_work_per_path": {"rays": cst_act["n_rays"] / cst_act["n_paths"],
                                 "box_tests": cst_act["n_box_tests"] / cst_act["n_paths"],
                                 "object_tests": cst_act["n_object_tests"] / cst_act["n_paths"],
                                 "hits": cst_act["n_hits"] / cst_act["n_paths"]},
        "hbm": {"achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved_gbs / peaks["hbm_gbs"], "peak_source": f"MEASURED_PEAKS.json ({peak_src})",
                "algorithmic_bytes_per_path": bytes_per_path},
        "work_per_path": {"rays": cst["n_rays"] / cst["n_paths"], "box_tests": cst["n_box_tests"] / cst["n_paths"],
                          "object_tests": cst["n_object_tests"] / cst["n_paths"], "hits": cst["n_hits"] / cst["n_paths"]},
    }

    # -- CPU baseline (oracle port on the box's host cores, bounded sample) -------------------------------------
    cpu = None
    if not a.no_cpu_baseline:
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
        import oracle_ffi as orc
        threads = os.cpu_count() or 1
        probe_rate, _ = cpu_render_rate(p, orc, world, cam, 1, threads)
        cpu_spp = a.cpu_spp or max(1, min(a.spp, int(probe_rate * 1e6 * 15.0 / npx)))
        rate, cst_cpu = cpu_render_rate(p, orc, world, cam, cpu_spp, threads)
        rate8, _ = cpu_render_rate(p, orc, world, cam, max(1, cpu_spp // 2), 8)
        cpu = {"value": rate, "unit": "Mpaths/s", "cores": threads, "kind": "port",
               "sample": f"{W}x{H} x {cpu_spp} spp ({cst_cpu['n_paths']} paths), depth {a.depth}, {threads} threads x static strips",
               "reference_faithful_8_threads": rate8,
               "note": "C++ restatement of the Zig renderer; expected to be faster than the original "
                       "(no 5.7 KB HitRecord copies, no CSPRNG)"}

    out = {
        "metric": "Mpaths/s (Book-1 final scene)", "value": value, "unit": "Mpaths/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
        "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, cam, integ_name),
        "mrays_per_s": value * cst["n_rays"] / cst["n_paths"],
        "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "integrator_probe": probe, "variants_mpaths_per_s": variants,
        "exchange": None if world_size == 1 else {
            "kind": "rtb_exchange_resolve: fused reduce-scatter + resolve + gather over peer memory, one kernel per rank"
                    if exchange_kind == "p2p" else "NCCL reduce(sum) to rank 0 + rtb_resolve_device",
            "ms_per_step": exchange_ms, "bytes_per_rank": 16 * npx, "probe": exchange_probe},
    }
    print(json.dumps(out))
    if world_size > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    args = parse_args()
    # Contract: rank 0 prints ONE JSON line on stdout.  Libraries (NCCL's version banner, make) also write to fd 1,
    # so everything but the final line is routed to stderr at the file-descriptor level.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import io
    captured = io.StringIO()
    py_stdout, sys.stdout = sys.stdout, captured
    try:
        rc = run_reference(args) if args.impl == "reference" else run_ours(args)
    finally:
        sys.stdout = py_stdout
    lines = [ln for ln in captured.getvalue().splitlines() if ln.strip()]
    for ln in lines[:-1]:
        print(ln, file=sys.stderr)
    if lines:
        os.write(real_stdout, (lines[-1] + "\n").encode())
    sys.exit(rc)
