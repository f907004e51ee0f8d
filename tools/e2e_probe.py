import sys, importlib, time, ctypes as C
sys.path.insert(0,'/root/repo')
import numpy as np, torch
p = importlib.import_module("zig-raytracing-weekend_b200")
world = p.World.book1(); cam = p.book1_camera(1200, 500, 50).init(); scene = p.Scene(world)
n = cam.image_width*cam.image_height
h_acc = torch.zeros(n,4).pin_memory(); h_rgba = torch.zeros(n,4,dtype=torch.uint8).pin_memory()
acc, rgba = h_acc.numpy(), h_rgba.numpy()
o = p.render_options(seed=1, integrator=1)
st = p.RtbRenderStats()
for rep in range(4):
    t0=time.perf_counter()
    acc[:] = 0.0
    t1=time.perf_counter()
    p._check(p._ffi.rtb().rtb_render(scene._h, C.byref(cam), C.byref(o), acc.ctypes.data, rgba.ctypes.data, C.byref(st)), "x")
    t2=time.perf_counter()
    print(f"memset {1e3*(t1-t0):.1f} ms; rtb_render wall {1e3*(t2-t1):.1f} ms device_ms {st.device_ms:.1f}")
pag = np.zeros((n,4),np.float32)
t0=time.perf_counter(); pag[:] = 0.0; t1=time.perf_counter(); print("pageable memset ms", 1e3*(t1-t0))
