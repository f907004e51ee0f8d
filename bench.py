#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 path-tracing hot path (contract: see the task brief / DESIGN.md §6).

  python bench.py [--gpus N] [--steps K] [--warmup W]             our arm (CUDA through the C ABI)
  python bench.py --impl reference [--gpus N] [--steps K] ...      the reference's CPU algorithm (oracle port)

Metric (BASELINE.json): Mpaths/s on the Book-1 final scene.  A *step* is one pass of the hot path over one
batch: `spp` samples per pixel of the 1200x675 Book-1 frame (BASELINE config 2: 500 spp, depth 50 = 405 M paths),
followed by the exchange step (NCCL sum-reduce to rank 0 when N > 1) and the resolve kernel (gamma + RGBA8).
N > 1 is WEAK scaling: the combined frame holds N*spp samples per pixel (sample-partitioned, Philox keyed by the
global sample index); the N*spp samples are split over the ranks in proportion to the rate each rank showed in the
warm-up steps, so a slower GPU gets fewer samples instead of making everybody wait.

  value  : device-resident — accumulators live in HBM, timed with CUDA events on the launch stream, max over ranks.
  value_reference_order : the same K steps timed with RTB_TRAVERSAL_REFERENCE (the mode whose nearest-hit index is
           bit-exact against the oracle on every ray, f32-undecidable ones included).
  e2e    : the same work through the reference-facing call with HOST buffers (rtb_render: H2D of the float4
           accumulation buffer, render, resolve, D2H of accum + RGBA8), pinned host memory, wall clock.
  roofline: binding ceiling for this cache-resident scene is the FP32 pipe; `achieved` = algorithmic
           flops of the REFERENCE traversal (counted by the kernel's counting build on the same workload) / kernel time;
           `traffic` = DRAM bytes per step from the ncu capture recorded in profiles/dram_per_path.json.
  parity : for the integrator/traversal that was TIMED: nearest-hit index mismatch rate against the oracle on >= 1e5
           camera + bounce rays, and RMSE / per-channel bias of a 1200x675 render against the oracle's at equal spp
           with BASELINE.md §4's thresholds evaluated.
  other_configs : BASELINE configs[2] (textured) and configs[3] (1 M spheres) at reduced spp, with their binding
           ceiling (FP32 for the cache-resident textured scene, HBM for the 96 MB sphere scene); c5_8k: configs[4]'s
           frame (strong scaling when N > 1).
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
TRAVERSALS = ("reference", "ordered", "sah", "sah16")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--integrator", default="auto", choices=["auto", "megakernel", "wavefront"])
    ap.add_argument("--traversal", default="sah16", choices=TRAVERSALS,
                    help="reference = the reference's left-then-right order over the host's tree (bit-exact hit index); "
                         "ordered = same tree, near child first; sah = the library's SAH re-partition of the same objects; "
                         "sah16 = that tree with 16-byte packed box nodes and a conservative half2 slab test")
    ap.add_argument("--scene", default="book1", choices=["book1", "textured", "million"],
                    help="book1 = BASELINE configs[1] (the metric's config); textured = configs[2]; million = configs[3]")
    ap.add_argument("--width", type=int, default=0, help="0 = the scene's BASELINE width")
    ap.add_argument("--spp", type=int, default=0, help="0 = the scene's BASELINE samples per pixel")
    ap.add_argument("--depth", type=int, default=DEPTH)
    ap.add_argument("--partition", default="samples", choices=["samples", "tiles"])
    ap.add_argument("--no-variants", action="store_true",
                    help="skip the integrator/traversal variant timings (keeps an ncu launch list to the headline path)")
    ap.add_argument("--exchange", choices=["auto", "p2p", "nccl"], default="auto",
                    help="N>1 exchange step: p2p = fused reduce+resolve kernel over peer memory (rtb_exchange_resolve), "
                         "nccl = NCCL sum-reduce to rank 0 then rtb_resolve_device, auto = time both on this box and "
                         "take the faster (measured: p2p at N=2, NCCL's in-switch reduction at N>=4)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: N*spp samples in total (sample partition only); strong: the ranks share --spp samples "
                         "(samples) or the frame's tiles (tiles)")
    ap.add_argument("--balance", default="rate", choices=["rate", "equal"],
                    help="N>1 sample split: rate = proportional to each rank's warm-up rate, equal = spp per rank")
    ap.add_argument("--cpu-spp", type=int, default=0, help="samples per pixel of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--parity-spp", type=int, default=64)
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-reference-order", action="store_true")
    ap.add_argument("--lean", action="store_true",
                    help="only the timed headline path: no variants / parity / other configs / reference order / CPU legs "
                         "(what an ncu launch list of this command should see)")
    a = ap.parse_args()
    if a.lean:
        a.no_variants = a.no_parity = a.no_other_configs = a.no_reference_order = a.no_cpu_baseline = a.no_e2e = True
        if a.integrator == "auto":  # no megakernel probe either (it would be 58 % of a serialised launch list)
            a.integrator = "wavefront"
    base = {"book1": (1200, 500), "textured": (800, 256), "million": (3840, 64)}[a.scene]
    a.width = a.width or base[0]
    a.spp = a.spp or base[1]
    return a


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


# ---------------------------------------------------------------------------------------------- scenes
def make_scene(p, name, width, spp, depth):
    """(world, Camera dataclass, workload label, moving fraction of the leaves) of a BASELINE.json config."""
    import numpy as np
    if name == "book1":
        return p.World.book1(), p.book1_camera(width, spp, depth), \
            "book1 (generateWorld, ~485 spheres + BVH)", 0.8
    if name == "textured":
        rgb = np.load(os.path.join(ROOT, "tests", "golden", "earthmap_rgb.npz"))["rgb"]
        img = np.ascontiguousarray(np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], axis=2))
        return p.World.create(p.RTW_SCENE_TEXTURED, image=img), p.textured_camera(width, spp, depth), \
            "textured (checker + earthmap image + perlin spheres)", 0.0
    if name == "million":
        return p.World.create(p.RTW_SCENE_RANDOM_SPHERES, n_spheres=1000000), p.million_camera(width, spp, depth), \
            "million (1 000 000 random spheres, 1 999 999-node BVH)", 0.0
    raise ValueError(name)


def scene_bytes(world):
    """Node + primitive bytes of the device layout the walk reads (32 B per node, 32 B per sphere)."""
    return 32 * world.n_nodes + 32 * world.n_objects


# ---------------------------------------------------------------------------------------------- roofline accounting
def algorithmic_work(st, width, height, moving_frac=0.8):
    """flops and bytes of a traversal, from the kernel's own work counters."""
    flops = 15 * st["n_box_tests"] + 23 * st["n_object_tests"] + (27 + 40) * st["n_hits"] + 30 * st["n_paths"] \
        + 6 * st["n_rays"]
    bytes_ = 32 * st["n_box_tests"] + (16 + 16 * moving_frac) * st["n_object_tests"] + 16 * st["n_hits"] \
        + 20 * width * height
    return flops, bytes_


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback"


def measured_dram_per_path(key):
    """DRAM bytes per path of one render, from the ncu capture tools/ncu_summary.py recorded (not typed in here)."""
    try:
        with open(os.path.join(ROOT, "profiles", "dram_per_path.json")) as f:
            e = json.load(f).get(key)
        return (e["bytes_per_path"], e["source"]) if e else (None, None)
    except OSError:
        return None, None


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_render_rate(p, orc, world, cam, spp, threads, seed=4321):
    o = p.render_options(seed=seed, sample_count=spp)
    _, _, st = orc.render(world.desc, cam, o, n_threads=threads, want_rgba=True)
    return st["n_paths"] / (st["device_ms"] * 1e-3) / 1e6, st


def workload_string(a, cam, label, per_gpu):
    return (f"{label} {cam.image_width}x{cam.image_height}, {a.spp} spp {'per GPU' if per_gpu else 'in total'}, "
            f"depth {a.depth}" + (" [BASELINE configs[1]]" if (a.scene, a.width, a.spp) == ("book1", WIDTH, SPP) else ""))


def run_reference(a):
    """--impl reference: the reference's CPU algorithm (oracle port; the Zig binary cannot be built here), all host
    threads, bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    p = importlib.import_module("zig-raytracing-weekend_b200")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    import oracle_ffi as orc
    world, camo, label, _ = make_scene(p, a.scene, a.width, a.spp, a.depth)
    cam = camo.init()
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
    weak = a.partition == "samples" and a.scaling == "weak"
    print(json.dumps({
        "impl": "reference", "metric": "Mpaths/s (Book-1 final scene)", "value": value, "unit": "Mpaths/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same workload string as our arm; everything else describes THIS arm truthfully (CPU, the reference's own
        # left-then-right walk over its random-axis median-split tree, static strips; no GPU, no L2 to flush)
        "config": {"workload": workload_string(a, cam, label, weak and a.gpus > 1), "scene_seed": 1, "bvh_seed": 2, "render_seed": 4321,
                   "implementation": "oracle/oracle.cpp (C++ restatement of camera.zig:93-208) on the host cores",
                   "threads": threads, "background": "sky gradient (camera.zig:204-206)" if a.scene == "book1"
                   else "solid (0.70, 0.80, 1.00)",
                   "sample_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


def workload_config(a, cam, label, integrator, world_size, per_gpu):
    return {"workload": workload_string(a, cam, label, per_gpu),
            "scene_seed": 1, "bvh_seed": 2, "render_seed": SEED, "integrator": integrator, "traversal": a.traversal,
            "partition": a.partition if world_size > 1 else "none",
            "background": "sky gradient (camera.zig:204-206)" if a.scene == "book1" else "solid (0.70, 0.80, 1.00)",
            "l2": "flushed between steps (256 MiB memset)" + ("; scene is 47 KB and cache/smem resident by nature"
                                                              if a.scene == "book1" else "")}


# ---------------------------------------------------------------------------------------------- parity block
def parity_block(p, orc, np, scene, world, camo, integrator, traversal, parity_spp):
    """Parity evidence for the integrator/traversal that was timed (VERDICT r1 'Next round' 1a)."""
    cam = camo.init()
    W, H = cam.image_width, cam.image_height
    npx = W * H
    threads = os.cpu_count() or 1
    out = {"integrator": "wavefront" if integrator == p.RTB_INTEGRATOR_WAVEFRONT else "megakernel",
           "traversal": TRAVERSALS[traversal], "oracle": "oracle/oracle.cpp — pinned on the reference's committed renders (sky, "
           "sphere silhouettes, metal reflection, refracted sky in the glass sphere, sky-lit lambertian sphere: tests/test_reference_renders.py) and its "
           "AABB / UV / stb_image fixtures; BVH visiting order, textures, quads and media unpinned"}
    # (a) nearest-hit index on >= 1e5 camera + bounce rays
    rng = np.random.default_rng(99)
    n_cam = 70000
    prim = orc.get_rays(cam, 77, rng.integers(0, npx, n_cam), rng.integers(0, 64, n_cam))
    ph = orc.trace_rays(world.desc, prim)
    sec = []
    for i in np.nonzero(ph["object"] >= 0)[0]:
        ok, _, sc = orc.scatter(world.desc, prim[i:i + 1], ph[i:i + 1], 5, int(i), 0, 1)
        if ok:
            sec.append(sc)
    rays = np.concatenate([prim] + sec)
    cpu = orc.trace_rays(world.desc, rays)
    gpu = scene.trace_rays(rays, traversal=traversal)
    same = gpu["object"] == cpu["object"]
    both = same & (cpu["object"] >= 0)
    out["hit_index"] = {"n_rays": int(rays.shape[0]), "n_camera": n_cam, "n_bounce": int(rays.shape[0] - n_cam),
                        "mismatches": int((~same).sum()), "mismatch_rate": float((~same).mean()),
                        "front_face_equal_where_same": bool(np.array_equal(gpu["front_face"][both], cpu["front_face"][both])),
                        "t_bit_exact_where_same": bool(np.array_equal(gpu["t"][both], cpu["t"][both])),
                        "bar": "bit-exact (0 mismatches) for traversal=reference; other traversals differ only on "
                               "f32-undecidable sphere tests (tests/test_gpu_parity.py)"}
    # (b) images at equal spp: RMSE_self from two oracle seeds, GPU with a third seed; and the same-stream compare
    o = lambda seed, **kw: p.render_options(seed=seed, sample_count=parity_spp, **kw)
    t0 = time.perf_counter()
    c1, _, _ = orc.render(world.desc, cam, o(4321), n_threads=threads, want_rgba=False)
    c2, _, _ = orc.render(world.desc, cam, o(8765), n_threads=threads, want_rgba=False)
    cpu_s = time.perf_counter() - t0
    g, _, _ = scene.render(cam, o(SEED, integrator=integrator, traversal=traversal), want_rgba=False)
    gs, _, _ = scene.render(cam, o(4321, integrator=integrator, traversal=traversal), want_rgba=False)
    m = lambda acc: acc[:, :3].astype(np.float64) / parity_spp
    rmse = lambda x, y: float(np.sqrt(np.mean((m(x) - m(y)) ** 2)))
    rmse_self, rmse_gc = rmse(c1, c2), rmse(g, c1)
    bias = (m(g) - m(c1)).mean(axis=0)
    bias_tol = max(0.002, 3.0 * rmse_self / (npx ** 0.5))
    q = lambda acc: orc.resolve(acc, float(parity_spp))[:, :3].astype(np.float64)
    diff = np.abs(gs[:, :3] - c1[:, :3]).max(axis=1)
    out["image"] = {"frame": f"{W}x{H}", "spp": parity_spp, "paths_per_render": npx * parity_spp,
                    "rmse_gpu_vs_cpu": rmse_gc, "rmse_self_cpu_seed_to_seed": rmse_self,
                    "rmse_threshold": 1.25 * rmse_self, "rmse_ok": bool(rmse_gc <= 1.25 * rmse_self),
                    "mean_bias_rgb": [float(b) for b in bias], "bias_threshold": bias_tol,
                    "bias_ok": bool(np.all(np.abs(bias) <= bias_tol)),
                    "rmse_8bit_after_gamma": float(np.sqrt(np.mean((q(g) - q(c1)) ** 2))),
                    "same_streams_pixels_off_by_more_than_1e-4": int(np.count_nonzero(diff > 1e-4)),
                    "same_streams_fraction_equal": float(np.mean(diff <= 1e-4)),
                    "oracle_seconds": cpu_s, "thresholds": "BASELINE.md §4"}
    return out


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

    world, camo, label, moving_frac = make_scene(p, a.scene, a.width, a.spp, a.depth)
    cam = camo.init()
    W, H = cam.image_width, cam.image_height
    npx = W * H
    scene = p.Scene(world, device=local_rank)
    stream = torch.cuda.current_stream(dev)
    sptr = stream.cuda_stream
    weak = a.partition == "samples" and a.scaling == "weak"

    # Buffers sized for the largest frame this run renders (the 8K frame of configs[4] when c5 is measured).
    want_c5 = a.scene == "book1" and not a.no_other_configs
    c5_W, c5_H = 7680, 4320
    npx_max = max(npx, c5_W * c5_H) if want_c5 else npx

    px = None
    if world_size > 1 and a.exchange != "nccl":
        # per-rank buffers owned by the library and mapped into every rank (CUDA IPC over NVLink/NVSwitch)
        try:
            px = mg.PeerExchange(npx_max, rank, world_size, local_rank)   # raises on EVERY rank if any rank cannot map
        except RuntimeError as e:
            if a.exchange == "p2p":
                raise
            print(f"[bench] peer-memory exchange unavailable, using NCCL: {e}", file=sys.stderr)
            px = None
    if px is not None:
        d_acc_all, d_rgba_all = px.accum, px.rgba
    else:
        d_acc_all = torch.zeros(npx_max, 4, device=dev, dtype=torch.float32)
        d_rgba_all = torch.zeros(npx_max, 4, device=dev, dtype=torch.uint8)
    d_acc, d_rgba = d_acc_all[:npx], d_rgba_all[:npx]
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    exchange_kind = "p2p" if px is not None else "nccl"

    def exchange_and_resolve(part, kind=None, acc=None, rgba=None, n=None):
        """The exchange step + resolve: afterwards rank 0 holds the combined sums and the RGBA8 frame."""
        acc = d_acc if acc is None else acc
        rgba = d_rgba if rgba is None else rgba
        n = npx if n is None else n
        if (kind or exchange_kind) == "p2p":
            px.exchange(float(part.total_samples), sptr, n_pixels=n)
            return
        mg.combine(acc, part, fix_w=False)
        if rank == 0:
            p._check(p._ffi.rtb().rtb_resolve_device(acc.data_ptr(), rgba.data_ptr(), n,
                                                     float(part.total_samples), local_rank, sptr), "rtb_resolve_device")

    exchange_probe = None
    if world_size > 1 and a.exchange == "auto" and px is not None:   # measured, not assumed: both exchanges on this frame size, this box
        part0 = mg.plan(a.partition, rank, world_size, a.spp, weak=weak)
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

    traversal = TRAVERSALS.index(a.traversal)
    weights = None   # per-rank share of the samples (rate-proportional split), None = equal

    def make_options(step, integrator, flags=0, spp=None, trav=None, scaling_weak=None):
        wk = weak if scaling_weak is None else scaling_weak
        part = mg.plan(a.partition, rank, world_size, spp or a.spp, sample_base=0, weak=wk,
                       weights=weights if (a.partition == "samples" and (spp is None or spp == a.spp)) else None)
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
        probe_spp = max(1, min(a.spp, 32 if a.scene != "million" else 2))
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
    count_spp = max(1, min(a.spp, 16 if a.scene != "million" else 1))
    cst = step_device(0, integrator, count=True, spp=count_spp, trav=p.RTB_TRAVERSAL_REFERENCE)
    cst_act = step_device(0, integrator, count=True, spp=count_spp)
    flops, bytes_ = algorithmic_work(cst, W, H, moving_frac)
    flops_per_path = flops / cst["n_paths"]
    bytes_per_path = (bytes_ - 20 * npx) / cst["n_paths"]
    flops_act, bytes_act = algorithmic_work(cst_act, W, H, moving_frac)
    flops_act_per_path = flops_act / cst_act["n_paths"]
    bytes_act_per_path = (bytes_act - 20 * npx) / cst_act["n_paths"]

    # -- warm-up (and, at N > 1, the rate every rank shows: the sample split of the timed steps follows it) -----
    warm_ms = []
    for w in range(a.warmup):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o, part = make_options(w, integrator)
        d_acc.zero_()
        e0.record(stream)
        scene.render_device(cam, o, d_acc.data_ptr(), sptr, want_stats=False)
        e1.record(stream)
        exchange_and_resolve(part)
        torch.cuda.synchronize(dev)
        warm_ms.append(e0.elapsed_time(e1))
    torch.cuda.synchronize(dev)
    balance = None
    if world_size > 1 and a.partition == "samples" and a.balance == "rate" and a.warmup >= 2:
        mine = torch.tensor([min(warm_ms[1:])], device=dev, dtype=torch.float64)   # first step includes one-time work
        allms = [torch.zeros_like(mine) for _ in range(world_size)]
        dist.all_gather(allms, mine)
        ms = [t.item() for t in allms]
        weights = [1.0 / t for t in ms]
        balance = {"warmup_render_ms_per_rank": ms, "samples_per_rank": None}

    # -- timed region: EXACTLY K steps, barrier + synchronize on both sides, CUDA events on the launch stream --
    def timed_region(trav=None, sample_clocks=False):
        clocks = ClockSampler(local_rank)
        if rank == 0 and sample_clocks:
            clocks.start()
        ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(a.steps)]
        e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e_begin.record(stream)
        for k in range(a.steps):
            flush.zero_()  # L2 flush between timed iterations
            o, part = make_options(k, integrator, trav=trav)
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
        clk = clocks.stop() if (rank == 0 and sample_clocks) else None
        t = torch.tensor([total_ms, kernel_ms, -exchange_ms, -kernel_ms], device=dev, dtype=torch.float64)
        if world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, kernel_ms, neg_x, neg_kmin = t.tolist()
        # exchange: MIN over ranks (the rank that arrives last sees the exchange without the wait)
        return total_ms, kernel_ms, -neg_x, -neg_kmin, clk, part

    total_ms, kernel_ms, exchange_ms, kernel_ms_min, clk, part_timed = timed_region(sample_clocks=True)
    if balance is not None:
        cnt = torch.tensor([part_timed.sample_count], device=dev, dtype=torch.int64)
        allc = [torch.zeros_like(cnt) for _ in range(world_size)]
        dist.all_gather(allc, cnt)
        balance["samples_per_rank"] = [int(c.item()) for c in allc]

    paths_per_step = npx * a.spp * (world_size if weak else 1)
    value = paths_per_step * a.steps / (total_ms * 1e-3) / 1e6
    # launches per step: one counting-free render reports them
    st1 = scene.render_device(cam, make_options(0, integrator)[0], d_acc.data_ptr(), sptr, want_stats=True)
    launches = (st1["n_launches"] + (1 if rank == 0 else 0)) * a.steps

    # -- the bit-exact mode, timed the same way (same K steps, same flush, same exchange) -------------------------
    ref_order = None
    if not a.no_reference_order and a.traversal != "reference":
        for w in range(min(a.warmup, 2)):
            step_device(w, integrator, trav=p.RTB_TRAVERSAL_REFERENCE)
        r_total, r_kernel, _, _, _, _ = timed_region(trav=p.RTB_TRAVERSAL_REFERENCE)
        r_value = paths_per_step * a.steps / (r_total * 1e-3) / 1e6
        ref_order = {"value": r_value, "unit": "Mpaths/s", "ms_per_step": r_total / a.steps, "kernel_ms_per_step": r_kernel,
                     "steps": a.steps, "traversal": "reference", "integrator": integ_name,
                     "parity": "nearest-hit index, front_face, t, p, normal bit-exact against the oracle on every ray "
                               "(tests/test_gpu_parity.py::test_trace_*, test_full_frame_*)"}

    # -- the other variants, one warmed-up step each (evidence for the choice; device-resident) ----
    variants = {}
    if not a.no_variants:
        for vint_name, vint in (("megakernel", p.RTB_INTEGRATOR_MEGAKERNEL), ("wavefront", p.RTB_INTEGRATOR_WAVEFRONT)):
            for vtrav, vtrav_name in enumerate(TRAVERSALS):
                vspp = max(1, min(a.spp, 100 if a.scene != "million" else 2))
                if a.scene == "million" and vtrav < 2:
                    continue    # the host's random-axis tree costs ~770 slab tests per ray there: seconds per sample
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                # one untimed render first: a mode's first render builds and uploads its layouts and learns the tail bounce
                step_device(0, vint, spp=max(1, vspp // 4), trav=vtrav)
                step_device(0, vint, spp=vspp, trav=vtrav)
                torch.cuda.synchronize(dev)
                e0.record(stream)
                step_device(0, vint, spp=vspp, trav=vtrav)
                e1.record(stream)
                torch.cuda.synchronize(dev)
                variants[f"{vint_name}/{vtrav_name}"] = npx * vspp * (world_size if weak else 1) / (e0.elapsed_time(e1) * 1e-3) / 1e6

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

    # -- BASELINE configs[4]: the 8K Book-1 frame, STRONG scaling (the ranks share the samples), same build, same run --
    c5 = None
    if want_c5:
        c5_spp = 256   # 8.5 G paths per step: enough samples per rank at N = 8 (32) for the batches to fill
        camo5 = p.book1_camera(c5_W, c5_spp, a.depth)
        cam5 = camo5.init()
        n5 = cam5.image_width * cam5.image_height
        acc5, rgba5 = d_acc_all[:n5], d_rgba_all[:n5]
        part5 = mg.plan("samples", rank, world_size, c5_spp, weak=False, weights=weights)

        def step_c5(step):
            o5 = mg.apply(part5, p.render_options(seed=SEED + step, integrator=integrator, traversal=traversal))
            acc5.zero_()
            x0, x1, x2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            x0.record(stream)
            scene.render_device(cam5, o5, acc5.data_ptr(), sptr, want_stats=False)
            x1.record(stream)
            exchange_and_resolve(part5, acc=acc5, rgba=rgba5, n=n5)
            x2.record(stream)
            return x0, x1, x2

        step_c5(0)
        torch.cuda.synchronize(dev)
        if world_size > 1:
            dist.barrier()
        c5_steps = 2
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(stream)
        evs = [step_c5(k) for k in range(c5_steps)]
        b1.record(stream)
        torch.cuda.synchronize(dev)
        t5 = torch.tensor([b0.elapsed_time(b1), sum(x0.elapsed_time(x1) for x0, x1, _ in evs) / c5_steps,
                           -sum(x1.elapsed_time(x2) for _, x1, x2 in evs) / c5_steps,
                           -sum(x0.elapsed_time(x1) for x0, x1, _ in evs) / c5_steps], device=dev, dtype=torch.float64)
        if world_size > 1:
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        tot5, ker5, negx5, negkmin5 = t5.tolist()
        # e2e: host buffers in the timed region (H2D of the cleared accumulator, D2H of sums + RGBA8 on rank 0)
        h5 = torch.zeros(n5, 4, dtype=torch.float32).pin_memory()
        h5r = torch.zeros(n5, 4, dtype=torch.uint8).pin_memory()
        torch.cuda.synchronize(dev)
        if world_size > 1:
            dist.barrier()
        t0 = time.perf_counter()
        acc5.copy_(h5, non_blocking=True)
        o5 = mg.apply(part5, p.render_options(seed=SEED, integrator=integrator, traversal=traversal))
        scene.render_device(cam5, o5, acc5.data_ptr(), sptr, want_stats=False)
        exchange_and_resolve(part5, acc=acc5, rgba=rgba5, n=n5)
        if rank == 0:
            h5.copy_(acc5, non_blocking=True)
            h5r.copy_(rgba5, non_blocking=True)
        torch.cuda.synchronize(dev)
        dt5 = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world_size > 1:
            dist.all_reduce(dt5, op=dist.ReduceOp.MAX)
        c5 = {"workload": f"book1 {cam5.image_width}x{cam5.image_height}, {c5_spp} spp in total (BASELINE configs[4] frame; "
                          f"1024 spp there), depth {a.depth}", "scaling": "strong", "n_gpus": world_size,
              "value": n5 * c5_spp * c5_steps / (tot5 * 1e-3) / 1e6, "unit": "Mpaths/s", "steps": c5_steps,
              "ms_per_step": tot5 / c5_steps, "render_ms_max_rank": ker5, "render_ms_min_rank": -negkmin5,
              "exchange_ms": -negx5, "samples_this_rank": part5.sample_count,
              "e2e": {"value": n5 * c5_spp / dt5.item() / 1e6, "unit": "Mpaths/s", "h2d_bytes_per_step": 16 * n5,
                      "d2h_bytes_per_step": 20 * n5},
              "traversal": a.traversal, "integrator": integ_name}
        del h5, h5r

    if px is not None:
        d_acc = d_rgba = d_acc_all = d_rgba_all = None
        if want_c5:
            acc5 = rgba5 = None
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
    hbm_binds = scene_bytes(world) > 0.5 * 126e6     # SURVEY §8d rule: node + primitive bytes <= L2/2 -> FP32 binds
    dram_key = f"{a.scene}/{integ_name}/{a.traversal}"
    dram_bpp, dram_src = measured_dram_per_path(dram_key)
    traffic = dram_bpp * paths_per_launch_group if dram_bpp is not None else None
    fp32_part = {
        "achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved_tflops / fp32_peak,
        "peak_source": "FFMA microbenchmark measured in this run (FMA = 2 flops); the kernels run unfused, ceiling = peak/2",
        "achieved_actual": flops_act_per_path * paths_per_launch_group / (kernel_ms * 1e-3) / 1e12,
        "frac_actual": flops_act_per_path * paths_per_launch_group / (kernel_ms * 1e-3) / 1e12 / fp32_peak}
    hbm_part = {"achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved_gbs / peaks["hbm_gbs"], "peak_source": f"MEASURED_PEAKS.json ({peak_src})",
                "algorithmic_bytes_per_path": bytes_per_path,
                "achieved_actual": (bytes_act_per_path * paths_per_launch_group + 20 * npx) / (kernel_ms * 1e-3) / 1e9,
                "frac_actual": (bytes_act_per_path * paths_per_launch_group + 20 * npx) / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "measured_dram_gbs": (traffic / (kernel_ms * 1e-3) / 1e9) if traffic else None,
                "measured_dram_frac": (traffic / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if traffic else None}
    bound = hbm_part if hbm_binds else fp32_part
    roofline = {
        "bound": "hbm" if hbm_binds else "fp32",
        "kernel": "render_megakernel" if integ_name == "megakernel" else "wf_raygen+wf_extend+wf_shade+wf_tail+wf_accumulate",
        "achieved": bound["achieved"], "peak": bound["peak"], "unit": bound["unit"], "frac": bound["frac"],
        "frac_actual": bound["frac_actual"], "peak_source": bound["peak_source"],
        "binding_rule": f"node + primitive bytes = {scene_bytes(world)} B {'>' if hbm_binds else '<='} L2/2 (SURVEY 8d)",
        # DRAM bytes of one step: ncu's dram__bytes_read.sum + dram__bytes_write.sum summed over every kernel of a render,
        # per path, as recorded by tools/ncu_summary.py dram in profiles/dram_per_path.json (queue/state traffic on
        # book1; node + sphere fetches that miss L2 on the million-sphere scene)
        "traffic": traffic, "traffic_key": dram_key, "traffic_source": dram_src,
        "algorithmic_flops_per_path": flops_per_path, "kernel_ms_per_step": kernel_ms,
        "kernel_ms_per_step_min_rank": kernel_ms_min,
        "counting": "achieved = work of the REFERENCE traversal (left-then-right over the host's tree, SURVEY 8d) "
                    "/ kernel time; *_actual = work of the traversal that ran",
        "actual_flops_per_path": flops_act_per_path,
        "actual_work_per_path": {"rays": cst_act["n_rays"] / cst_act["n_paths"],
                                 "box_tests": cst_act["n_box_tests"] / cst_act["n_paths"],
                                 "object_tests": cst_act["n_object_tests"] / cst_act["n_paths"],
                                 "hits": cst_act["n_hits"] / cst_act["n_paths"]},
        "fp32": fp32_part, "hbm": hbm_part,
        **({"note": "hbm frac > 1: SURVEY 8d's algorithmic node / primitive bytes are mostly served by L1 and L2 (the walk revisits "
                    "the top of the tree); hbm.measured_dram_frac is the DRAM-side fraction, and the walk is bound by L2 "
                    "latency, not by HBM bandwidth (DESIGN.md section 5, C4)"} if hbm_binds and bound["frac"] > 1.0 else {}),
        "work_per_path": {"rays": cst["n_rays"] / cst["n_paths"], "box_tests": cst["n_box_tests"] / cst["n_paths"],
                          "object_tests": cst["n_object_tests"] / cst["n_paths"], "hits": cst["n_hits"] / cst["n_paths"]},
    }
    if ref_order is not None:
        rt = flops_per_path * paths_per_launch_group / (ref_order["kernel_ms_per_step"] * 1e-3) / 1e12
        ref_order["roofline_frac_fp32"] = rt / fp32_peak

    orc = None
    if not (a.no_cpu_baseline and a.no_parity):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
        import oracle_ffi as orc

    # -- parity of the timed mode (N = 1 only: it is a property of the kernels, not of the partition) --------------
    parity = None
    if not a.no_parity and world_size == 1 and a.scene != "million":
        parity = parity_block(p, orc, np, scene, world, camo, integrator, traversal, a.parity_spp)

    # -- the other named configs at reduced spp (N = 1 only) --------------------------------------------------------
    other = None
    if not a.no_other_configs and world_size == 1 and a.scene == "book1":
        other = {}
        for oname, ospp, ocount in (("textured", 256, 16), ("million", 8, 1)):
            ow, ocamo, olabel, omov = make_scene(p, oname, 0 or {"textured": 800, "million": 3840}[oname], ospp, a.depth)
            ocam = ocamo.init()
            on = ocam.image_width * ocam.image_height
            oscene = p.Scene(ow, device=local_rank)
            oacc = torch.zeros(on, 4, device=dev, dtype=torch.float32)
            otrav = traversal
            o_opts = lambda **kw: p.render_options(seed=SEED, integrator=integrator, traversal=otrav, **kw)
            oscene.render_device(ocam, o_opts(sample_count=min(ospp, 4)), oacc.data_ptr(), sptr, want_stats=False)  # warm
            ost_c = oscene.render_device(ocam, o_opts(sample_count=ocount, flags=p.RTB_FLAG_COUNT_WORK), oacc.data_ptr(), sptr)
            best = None
            for _ in range(2):
                oacc.zero_()
                ost = oscene.render_device(ocam, o_opts(sample_count=ospp), oacc.data_ptr(), sptr)
                best = ost if best is None or ost["device_ms"] < best["device_ms"] else best
            oflops, obytes = algorithmic_work(ost_c, ocam.image_width, ocam.image_height, omov)
            sec = best["device_ms"] * 1e-3
            scale = best["n_paths"] / ost_c["n_paths"]
            hbm_b = scene_bytes(ow) > 0.5 * 126e6
            okey = f"{oname}/{integ_name}/{a.traversal}"
            obpp, osrc = measured_dram_per_path(okey)
            rec = {"workload": f"{olabel} {ocam.image_width}x{ocam.image_height}, {ospp} spp (BASELINE: "
                               f"{ {'textured': 256, 'million': 64}[oname] } spp), depth {a.depth}",
                   "value": best["n_paths"] / sec / 1e6, "unit": "Mpaths/s", "ms": best["device_ms"],
                   "mrays_per_s": best["n_paths"] / sec / 1e6 * ost_c["n_rays"] / ost_c["n_paths"],
                   "launches": best["n_launches"], "traversal": a.traversal, "integrator": integ_name,
                   "work_per_ray_actual": {"box_tests": ost_c["n_box_tests"] / ost_c["n_rays"],
                                           "object_tests": ost_c["n_object_tests"] / ost_c["n_rays"]},
                   "rays_per_path": ost_c["n_rays"] / ost_c["n_paths"], "scene_bytes": scene_bytes(ow),
                   "bound": "hbm" if hbm_b else "fp32",
                   "fp32_frac_actual": oflops * scale / sec / 1e12 / fp32_peak,
                   "hbm_frac_algorithmic_actual": obytes * scale / sec / 1e9 / peaks["hbm_gbs"],
                   "measured_dram_bytes_per_path": obpp, "measured_dram_source": osrc,
                   "hbm_frac_measured_dram": (obpp * best["n_paths"] / sec / 1e9 / peaks["hbm_gbs"]) if obpp else None,
                   "counting": "actual-count (the traversal that ran); the reference-order walk of this scene is not "
                               "run here" if oname == "million" else "actual-count (the traversal that ran)"}
            other[oname] = rec
            oscene.close()
            del oacc

    # -- CPU baseline (oracle port on the box's host cores, bounded sample) -------------------------------------
    cpu = None
    if not a.no_cpu_baseline:
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
        "metric": "Mpaths/s (Book-1 final scene)" if a.scene == "book1" else f"Mpaths/s ({a.scene} scene)",
        "value": value, "unit": "Mpaths/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
        "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, cam, label, integ_name, world_size, weak and world_size > 1),
        "mrays_per_s": value * cst["n_rays"] / cst["n_paths"],
        "value_reference_order": ref_order,
        "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "parity": parity, "other_configs": other, "c5_8k": c5,
        "integrator_probe": probe, "variants_mpaths_per_s": variants, "balance": balance,
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
