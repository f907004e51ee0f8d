#!/usr/bin/env python
"""How long does the HOST spend inside rtb_render_device (enqueueing a whole render) compared with the device time?
If the two are close the launch queue is the bottleneck (development aid)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
p = importlib.import_module("zig-raytracing-weekend_b200")
import torch
world, cam = p.World.book1(), p.book1_camera(1200, 256, 50).init()
scene = p.Scene(world)
acc = torch.zeros(cam.image_width * cam.image_height, 4, device="cuda")
for rep in range(3):
    acc.zero_(); torch.cuda.synchronize()
    o = p.render_options(seed=1234, integrator=1, traversal=3)
    t = time.perf_counter()
    scene.render_device(cam, o, acc.data_ptr(), 0, want_stats=False)
    host = (time.perf_counter() - t) * 1e3
    torch.cuda.synchronize()
    total = (time.perf_counter() - t) * 1e3
    print(f"rep {rep}: host returned from the call after {host:.2f} ms, the device finished after {total:.2f} ms")
