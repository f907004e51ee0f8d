"""Development aid: classify rays on which RTB_TRAVERSAL_SAH and the reference-order oracle disagree."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("zig-raytracing-weekend_b200")
import oracle_ffi as orc
world = pkg.World.create(pkg.RTW_SCENE_RANDOM_SPHERES, n_spheres=200000)
scene = pkg.Scene(world)
cam = pkg.million_camera(256, 2, 8).init()
rng = np.random.default_rng(17)
size = cam.image_width * cam.image_height
r1 = orc.get_rays(cam, 3, rng.integers(0, size, 6000).astype(np.uint32), rng.integers(0, 64, 6000).astype(np.uint32))
r2 = np.zeros(6000, dtype=np.dtype(pkg._ffi.RAY_DTYPE))
r2["origin"] = rng.uniform(-450, 450, (6000, 3)).astype(np.float32)
r2["direction"] = rng.normal(size=(6000, 3)).astype(np.float32)
r2["time"] = rng.random(6000).astype(np.float32); r2["t_min"] = 0.001; r2["t_max"] = np.inf
r2["origin"][:, 1] = rng.uniform(0.1, 40, 6000).astype(np.float32)
rays = np.concatenate([r1, r2])
cpu = orc.trace_rays(world.desc, rays)
for mode in (1, 2):
    got = scene.trace_rays(rays, traversal=mode)
    bad = np.nonzero(got["object"] != cpu["object"])[0]
    print("mode", mode, "mismatches", len(bad))
    d = world.desc.contents
    H = np.ctypeslib.as_array(C := None) if False else None
    for k in bad[:40]:
        o = rays["origin"][k].astype(np.float64); dd = rays["direction"][k].astype(np.float64)
        def info(obj):
            if obj < 0: return "none"
            h = d.hittables[int(obj)]
            c = np.array(h.a[:], np.float64); r = float(h.radius)
            oc = o - c; a = dd @ dd; hb = oc @ dd; cc = oc @ oc - r * r
            disc = hb * hb - a * cc
            return f"obj {obj} r={r:.3f} |oc|={np.sqrt(oc@oc):.2f} disc/a={disc/a:.3e} t64={(-hb-np.sqrt(max(disc,0)))/a:.6f}"
        print(k, "ref:", cpu["object"][k], cpu["t"][k], info(cpu["object"][k]), "| got:", got["object"][k], got["t"][k], info(got["object"][k]))
