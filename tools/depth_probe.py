import importlib, sys, torch
sys.path.insert(0, "/root/repo")
p = importlib.import_module("zig-raytracing-weekend_b200")
world = p.World.book1(); scene = p.Scene(world)
for depth in (50, 16, 12, 10, 8):
    cam = p.book1_camera(1200, 300, depth).init()
    acc = torch.zeros(cam.image_width * cam.image_height, 4, device="cuda")
    for rep in range(3):
        acc.zero_()
        o = p.render_options(seed=1234, integrator=1, traversal=2)
        st = scene.render_device(cam, o, acc.data_ptr(), 0)
    print(f"depth {depth}: {st['device_ms']:.2f} ms, {st['n_paths']/st['device_ms']/1e3:.1f} Mpaths/s, launches {st['n_launches']}", flush=True)
