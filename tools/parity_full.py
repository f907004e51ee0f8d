#!/usr/bin/env python
"""One slow run per round: BASELINE configs[1] at its FULL 500 spp on the oracle (405 M paths, ~1.5-3 min of CPU per
render) against the GPU at equal spp — RMSE and per-channel mean bias of the linear images, with BASELINE.md §4's
thresholds — for the benchmarked mode (wavefront + SAH16) and for the bit-exact mode (wavefront + reference order).
Writes one JSON document (commit it under profiles/).

    python tools/parity_full.py [--spp 500] [--width 1200] > profiles/r2_parity_c2_500spp.json
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=1200)
ap.add_argument("--spp", type=int, default=500)
ap.add_argument("--depth", type=int, default=50)
a = ap.parse_args()

p = importlib.import_module("zig-raytracing-weekend_b200")
import oracle_ffi as orc  # noqa: E402  (test infrastructure: the checker)

world = p.World.book1()
cam = p.book1_camera(a.width, a.spp, a.depth).init()
W, H = cam.image_width, cam.image_height
npx = W * H
threads = os.cpu_count() or 1
scene = p.Scene(world)

t0 = time.perf_counter()
c1, _, s1 = orc.render(world.desc, cam, p.render_options(seed=4321), n_threads=threads, want_rgba=False)
t1 = time.perf_counter()
c2, _, _ = orc.render(world.desc, cam, p.render_options(seed=8765), n_threads=threads, want_rgba=False)
t2 = time.perf_counter()
mean = lambda acc: acc[:, :3].astype(np.float64) / a.spp
rmse = lambda x, y: float(np.sqrt(np.mean((mean(x) - mean(y)) ** 2)))
rmse_self = rmse(c1, c2)
bias_tol = max(0.002, 3.0 * rmse_self / npx ** 0.5)
out = {"workload": f"book1 {W}x{H}, {a.spp} spp, depth {a.depth} (BASELINE configs[1])", "paths_per_render": npx * a.spp,
       "oracle": {"threads": threads, "seconds_per_render": [t1 - t0, t2 - t1], "mpaths_per_s": npx * a.spp / (t1 - t0) / 1e6,
                  "rays_per_path": s1["n_rays"] / s1["n_paths"], "note": "parity unpinned (SURVEY §8c)"},
       "rmse_self_cpu_seed_to_seed": rmse_self, "rmse_threshold": 1.25 * rmse_self, "bias_threshold": bias_tol,
       "thresholds": "BASELINE.md §4: RMSE(GPU,CPU) <= 1.25 RMSE_self; |mean bias| <= max(0.002, 3 RMSE_self / sqrt(WH))",
       "modes": {}}
for name, trav in (("wavefront/sah16", p.RTB_TRAVERSAL_SAH16), ("wavefront/sah", p.RTB_TRAVERSAL_SAH),
                   ("wavefront/reference", p.RTB_TRAVERSAL_REFERENCE)):
    o = lambda seed: p.render_options(seed=seed, integrator=p.RTB_INTEGRATOR_WAVEFRONT, traversal=trav)
    g, _, st = scene.render(cam, o(1234), want_rgba=False)         # different seed: statistical parity
    gs, _, _ = scene.render(cam, o(4321), want_rgba=False)         # the oracle's seed: same paths, pixel by pixel
    bias = (mean(g) - mean(c1)).mean(axis=0)
    diff = np.abs(gs[:, :3] - c1[:, :3]).max(axis=1) / a.spp
    q = lambda acc: orc.resolve(acc, float(a.spp))[:, :3].astype(np.float64)
    out["modes"][name] = {
        "rmse_gpu_vs_cpu": rmse(g, c1), "rmse_ok": bool(rmse(g, c1) <= 1.25 * rmse_self),
        "mean_bias_rgb": [float(b) for b in bias], "bias_ok": bool(np.all(np.abs(bias) <= bias_tol)),
        "rmse_8bit_after_gamma": float(np.sqrt(np.mean((q(g) - q(c1)) ** 2))),
        "same_streams": {"pixels_whose_mean_differs_by_more_than_1e-5": int(np.count_nonzero(diff > 1e-5)),
                         "max_abs_diff_of_mean": float(diff.max()), "rmse": rmse(gs, c1),
                         "rgba8_pixels_differing": int(np.count_nonzero((q(gs) != q(c1)).any(axis=1)))},
        "gpu_ms": st["device_ms"], "gpu_mpaths_per_s": st["n_paths"] / st["device_ms"] / 1e3}
print(json.dumps(out, indent=1))
