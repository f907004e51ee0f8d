#!/usr/bin/env python
"""Small end-to-end run of every kernel for compute-sanitizer (development aid)."""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
p = importlib.import_module("zig-raytracing-weekend_b200")
img = np.load(os.path.join(ROOT, "tests/golden/earthmap_rgb.npz"))["rgb"]
img = np.ascontiguousarray(np.concatenate([img, np.full(img.shape[:2] + (1,), 255, np.uint8)], axis=2))
worlds = [(p.World.book1(), p.book1_camera(120, 3, 12)), (p.World.create(p.RTW_SCENE_TEXTURED, image=img), p.textured_camera(96, 2, 8)),
          (p.World.create(p.RTW_SCENE_CORNELL_SMOKE), p.cornell_camera(64, 2, 8)),
          (p.World.create(p.RTW_SCENE_RANDOM_SPHERES, n_spheres=30000), p.million_camera(96, 2, 6))]
rng = np.random.default_rng(1)
for world, camo in worlds:
    cam = camo.init()
    scene = p.Scene(world)
    rays = np.zeros(3000, dtype=np.dtype(p._ffi.RAY_DTYPE))
    rays["origin"] = rng.uniform(-10, 10, (3000, 3)); rays["direction"] = rng.normal(size=(3000, 3))
    rays["t_min"] = 0.001; rays["t_max"] = np.inf; rays["time"] = rng.random(3000)
    for trav in (0, 1, 2, 3):
        scene.trace_rays(rays, traversal=trav)
        for integ in (0, 1):
            scene.render(cam, p.render_options(seed=3, integrator=integ, traversal=trav, flags=p.RTB_FLAG_COUNT_WORK))
            scene.render(cam, p.render_options(seed=3, integrator=integ, traversal=trav))
    scene.close()
import torch
lib = p._ffi.rtb()
n = 5000
bufs = [torch.rand(n, 4, device="cuda") for _ in range(3)]
out, rgba = torch.zeros(n, 4, device="cuda"), torch.zeros(n, 4, dtype=torch.uint8, device="cuda")
peers = (C.c_void_p * 3)(*[b.data_ptr() for b in bufs])
for r in range(3):
    assert lib.rtb_exchange_resolve(peers, 3, r, 0, out.data_ptr(), rgba.data_ptr(), n, 4.0, 0, None) == 0
torch.cuda.synchronize()
print("sanitize smoke done")
