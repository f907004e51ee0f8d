#!/usr/bin/env python
"""Ad-hoc device-resident timing of both integrators (development aid; bench.py is the contract)."""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
p = importlib.import_module("zig-raytracing-weekend_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=1200)
ap.add_argument("--spp", type=int, default=32)
ap.add_argument("--scene", default="book1")
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--integrators", default="0,1")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--count", action="store_true")
ap.add_argument("--traversal", type=int, default=0)
ap.add_argument("--profile-last", action="store_true", help="cudaProfilerStart/Stop around the last repetition (ncu --replay-mode range)")
a = ap.parse_args()

import torch
if a.scene == "book1":
    world, camo = p.World.book1(), p.book1_camera(a.width, a.spp, 50)
elif a.scene == "million":
    world, camo = p.World.create(p.RTW_SCENE_RANDOM_SPHERES, n_spheres=a.n), p.million_camera(a.width, a.spp, 50)
else:
    img = np.load(os.path.join(ROOT, "tests/golden/earthmap_rgb.npz"))["rgb"]
    img = np.concatenate([img, np.full(img.shape[:2] + (1,), 255, np.uint8)], axis=2)
    world, camo = p.World.create(p.RTW_SCENE_TEXTURED, image=img), p.textured_camera(a.width, a.spp, 50)
cam = camo.init()
scene = p.Scene(world)
npx = cam.image_width * cam.image_height
acc = torch.zeros(npx, 4, device="cuda")
for integ in [int(x) for x in a.integrators.split(",")]:
    for rep in range(a.reps):
        acc.zero_()
        o = p.render_options(seed=1234, integrator=integ, traversal=a.traversal, flags=p.RTB_FLAG_COUNT_WORK if a.count else 0)
        torch.cuda.synchronize()
        prof = a.profile_last and rep == a.reps - 1
        if prof:
            torch.cuda.cudart().cudaProfilerStart()
        st = scene.render_device(cam, o, acc.data_ptr(), 0)
        if prof:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        mp = st["n_paths"] / st["device_ms"] / 1e3
        extra = ""
        if a.count:
            extra = (f" rays/path={st['n_rays']/st['n_paths']:.2f} box/ray={st['n_box_tests']/st['n_rays']:.1f}"
                     f" obj/ray={st['n_object_tests']/st['n_rays']:.1f} Mrays/s={st['n_rays']/st['device_ms']/1e3:.1f}")
        print(f"{a.scene} {cam.image_width}x{cam.image_height} spp={a.spp} integrator={integ} traversal={a.traversal} rep={rep}: "
              f"{st['device_ms']:.2f} ms, {mp:.1f} Mpaths/s, launches={st['n_launches']}{extra}", flush=True)
