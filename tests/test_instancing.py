"""General instancing (src/objects.zig:264-443): Translate / RotateY / HittableList / ConstantMedium wrapping ANY
hittable, as one BVH leaf.  The reference evaluates them recursively (`inline else => |object| object.hit(r, ray_t)`);
so do the oracle (anyHit) and the device (hit_any).  Round 1 only lowered the one shape HEAD's scenes build,
Translate(RotateY(createBox)), as a flat record — that stays as the fast path and must agree with the general one."""
import ctypes as C

import numpy as np
import pytest


def _rays(pkg, rng, n, lo, hi):
    rays = np.zeros(n, dtype=np.dtype(pkg._ffi.RAY_DTYPE))
    rays["origin"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    target = rng.uniform(-2.5, 2.5, (n, 3)).astype(np.float32)
    rays["direction"] = target - rays["origin"]
    rays["time"] = rng.random(n).astype(np.float32)
    rays["t_min"], rays["t_max"] = 0.001, np.inf
    return rays


def _grey(pkg, c=(0.7, 0.7, 0.7)):
    return pkg.material_spec(color=c)


def _zoo(pkg):
    """One world with every wrapper kind, nested in several orders, next to plain objects."""
    w = pkg.World.new()
    metal = pkg.material_spec(material=pkg.RTB_MAT_METAL, color=(0.8, 0.6, 0.2), fuzz=0.2)
    glass = pkg.material_spec(material=pkg.RTB_MAT_DIELECTRIC, ir=1.5)
    light = pkg.material_spec(material=pkg.RTB_MAT_DIFFUSE_LIGHT, color=(4, 4, 4))
    checker = pkg.material_spec(texture=pkg.RTB_TEX_CHECKER, color=(0.2, 0.3, 0.1), color2=(0.9, 0.9, 0.9), scale=0.5)
    # 0: a translated moving sphere
    w.add_object(w.obj_translate(w.obj_sphere((0, 0, 0), 0.6, _grey(pkg), center2=(0.2, 0.1, 0)), (1.5, 0.5, -1.0)))
    # 1: a rotated quad (not axis-aligned any more)
    w.add_object(w.obj_rotate_y(w.obj_quad((-1, -1, 0), (2, 0, 0), (0, 2, 0), checker), 35.0))
    # 2: a list of mixed members as ONE leaf: sphere + quad + createBox, each with its own material
    members = [w.obj_sphere((-2, 0.3, 1.5), 0.5, metal), w.obj_quad((-2.8, -0.5, 2.4), (1.6, 0, 0), (0, 1.2, 0.3), _grey(pkg, (0.2, 0.4, 0.9))),
               w.obj_box((-2.6, -0.9, 0.4), (-1.6, -0.3, 1.0), glass)]
    w.add_object(w.obj_list(members))
    # 3: Translate(RotateY(Translate(list))) — three levels over a list of two spheres
    inner = w.obj_list([w.obj_sphere((0, 0, 0), 0.35, _grey(pkg, (0.9, 0.2, 0.2))), w.obj_sphere((0.6, 0.2, 0), 0.25, light)])
    w.add_object(w.obj_translate(w.obj_rotate_y(w.obj_translate(inner, (0.3, 0, 0)), -50.0), (0.5, -1.2, 1.0)))
    # 4: a ball of fog: ConstantMedium over a SPHERE boundary (the book's final scene does this; HEAD only over boxes)
    w.add_object(w.obj_medium(w.obj_sphere((1.8, -0.8, 1.6), 0.8, glass), 1.5, (0.9, 0.9, 1.0)))
    # 5: fog in a rotated, translated box, built from the general wrappers
    w.add_object(w.obj_medium(w.obj_translate(w.obj_rotate_y(w.obj_box((0, 0, 0), (1, 1.4, 1), _grey(pkg)), 20.0), (-0.5, 0.6, -2.2)),
                              0.9, (0.1, 0.1, 0.1)))
    # plain neighbours
    w.add_sphere((0, -101.5, 0), 100.0, _grey(pkg, (0.5, 0.5, 0.5)))
    w.add_sphere((0.2, 1.6, 0.3), 0.4, metal)
    return w.build()


def test_wrapper_boxes_and_lowering(pkg):
    """Host side: the general wrappers keep the reference's bounding-box arithmetic (Translate.init :314-319,
    RotateY.init :360-397, HittableList.add :274-277 starting from Aabb{} = the origin) — checked against the flat
    Translate(RotateY(createBox)) record, which the oracle's orc_box_bbox already pins — and lower to children that
    follow their wrapper."""
    flat = pkg.World.new()
    flat.add_box((0, 0, 0), (1, 2, 3), _grey(pkg), angle=33.0, offset=(0.5, -0.25, 1.0))
    flat.add_sphere((9, 9, 9), 0.5, _grey(pkg))
    flat.build()
    gen = pkg.World.new()
    gen.add_object(gen.obj_translate(gen.obj_rotate_y(gen.obj_box((0, 0, 0), (1, 2, 3), _grey(pkg)), 33.0), (0.5, -0.25, 1.0)))
    gen.add_sphere((9, 9, 9), 0.5, _grey(pkg))
    gen.build()
    fb = [flat.object_box(i) for i in range(2)]
    gb = [gen.object_box(i) for i in range(2)]
    assert any(np.array_equal(fb[i], gb[j]) for i in range(2) for j in range(2) if fb[i][3] - fb[i][0] > 1.01)
    d = gen.desc.contents
    assert d.n_hittables == 4                       # translate, sphere, + rotate_y, + box as children
    types = [d.hittables[i].type for i in range(4)]
    assert sorted(types[:2]) == [pkg.RTB_HITTABLE_SPHERE, pkg.RTB_HITTABLE_TRANSLATE]
    for i in range(4):
        h = d.hittables[i]
        if h.type in (pkg.RTB_HITTABLE_TRANSLATE, pkg.RTB_HITTABLE_ROTATE_Y):
            assert h.child > i and d.hittables[h.child].type in (pkg.RTB_HITTABLE_ROTATE_Y, pkg.RTB_HITTABLE_BOX)
    # validation (no device needed): a child must follow its wrapper, ranges must fit
    ffi = pkg._ffi
    bad = ffi.RtbSceneDesc.from_buffer_copy(bytes(d))
    hs = (ffi.RtbHittable * 4)(*[d.hittables[i] for i in range(4)])
    bad.hittables = C.cast(hs, C.POINTER(ffi.RtbHittable))
    top = [i for i in range(4) if hs[i].type == pkg.RTB_HITTABLE_TRANSLATE][0]
    n = C.c_uint32(0)
    hs[top].child = top
    assert ffi.rtb().rtb_debug_build_layout(C.byref(bad), 0, 0, None, C.byref(n)) == ffi.RTB_ERR_INVALID_ARGUMENT
    hs[top].child = 99
    assert ffi.rtb().rtb_debug_build_layout(C.byref(bad), 0, 0, None, C.byref(n)) == ffi.RTB_ERR_INVALID_ARGUMENT
    hs[top].type = 8
    assert ffi.rtb().rtb_debug_build_layout(C.byref(bad), 0, 0, None, C.byref(n)) == ffi.RTB_ERR_UNSUPPORTED


@pytest.mark.gpu
def test_general_wrappers_agree_with_the_flat_box_record(pkg, orc):
    rng = np.random.default_rng(31)
    flat = pkg.World.new()
    flat.add_box((-1, -1, -1), (2, 1, 1), _grey(pkg), angle=33.0, offset=(0.5, -0.25, 1.0))
    flat.build()
    gen = pkg.World.new()
    gen.add_object(gen.obj_translate(gen.obj_rotate_y(gen.obj_box((-1, -1, -1), (2, 1, 1), _grey(pkg)), 33.0), (0.5, -0.25, 1.0)))
    gen.build()
    rays = _rays(pkg, rng, 20000, -6, 6)
    a = pkg.Scene(flat).trace_rays(rays)
    sg = pkg.Scene(gen)
    c = orc.trace_rays(gen.desc, rays)
    assert (c["object"] >= 0).mean() > 0.2
    for mode in (0, 1, 2, 3):
        b = sg.trace_rays(rays, traversal=mode)
        for k in ("object", "front_face", "t", "p", "normal", "u", "v"):
            assert np.array_equal(a[k], b[k]), (mode, k)            # flat fast path == general recursion, bit for bit
            assert np.array_equal(b[k], c[k]), (mode, k)            # == the oracle's recursion


@pytest.mark.gpu
def test_instancing_zoo_trace_parity(pkg, orc):
    world = _zoo(pkg)
    scene = pkg.Scene(world)
    rng = np.random.default_rng(32)
    rays = np.concatenate([_rays(pkg, rng, 40000, -7, 7), _rays(pkg, rng, 10000, -2, 2)])
    cpu = orc.trace_rays(world.desc, rays)
    d = world.desc.contents
    tops = {d.hittables[i].type for i in range(8)}
    assert {pkg.RTB_HITTABLE_TRANSLATE, pkg.RTB_HITTABLE_ROTATE_Y, pkg.RTB_HITTABLE_LIST, pkg.RTB_HITTABLE_MEDIUM_OF} <= tops
    hit_objects = set(cpu["object"][cpu["object"] >= 0].tolist())
    assert hit_objects == set(range(8)), hit_objects                 # every top-level object is hit by some ray
    is_medium = np.array([o >= 0 and d.hittables[o].type == pkg.RTB_HITTABLE_MEDIUM_OF for o in cpu["object"]])
    assert is_medium.sum() > 500
    for mode in (0, 1, 2, 3):
        gpu = scene.trace_rays(rays, traversal=mode)
        same = gpu["object"] == cpu["object"]
        # log() differs by ulps between CUDA and glibc: a fog hit can flip only when the draw lands within an ulp of
        # the chord's end, and its t carries that ulp; everything solid is bit-exact
        assert (~same).mean() < 1e-3, (mode, int((~same).sum()))
        hit = (cpu["object"] >= 0) & same
        solid = hit & ~is_medium
        for k in ("front_face", "t", "p", "normal"):
            assert np.array_equal(gpu[k][solid], cpu[k][solid]), (mode, k)
        fog = hit & is_medium
        np.testing.assert_allclose(gpu["t"][fog], cpu["t"][fog], rtol=2e-5)
        np.testing.assert_allclose(gpu["p"][fog], cpu["p"][fog], rtol=2e-5, atol=1e-5)
        assert (gpu["normal"][fog] == [1, 0, 0]).all() and (gpu["front_face"][fog] == 1).all()
        np.testing.assert_allclose(gpu["u"][solid], cpu["u"][solid], rtol=1e-5, atol=1e-5)   # acos / atan2 of sphere members
        np.testing.assert_allclose(gpu["v"][solid], cpu["v"][solid], rtol=1e-5, atol=1e-5)
        if mode == 0:
            exact = same.all()
            if not exact:
                continue
            assert np.array_equal(gpu["n_box_tests"], cpu["n_box_tests"])
            assert np.array_equal(gpu["n_object_tests"], cpu["n_object_tests"])


@pytest.mark.gpu
@pytest.mark.parametrize("integrator", [0, 1])
def test_instancing_zoo_render_parity(pkg, orc, integrator):
    """Same Philox streams as the oracle: materials come from the primitive hit INSIDE the wrapper (metal, glass,
    checker, a light, two fogs), the records are rotated / translated back on the way out of the recursion."""
    world = _zoo(pkg)
    scene = pkg.Scene(world)
    cam = pkg.Camera(image_width=128, aspect_ratio=1.0, samples_per_pixel=6, max_depth=30, vfov=60.0, lookfrom=(0.5, 1.0, 7.0),
                     lookat=(0, 0, 0), defocus_angle=0.0, background=(0.6, 0.7, 0.9)).init()
    for mode in (0, 3):
        o = pkg.render_options(seed=5, integrator=integrator, traversal=mode, flags=pkg.RTB_FLAG_COUNT_WORK)
        g, _, gs = scene.render(cam, o)
        c, _, cs = orc.render(world.desc, cam, o, n_threads=8)
        diff = np.abs(g[:, :3] - c[:, :3]).max(axis=1)
        tol = 1e-4 * np.maximum(1.0, np.abs(c[:, :3]).max(axis=1))
        assert np.count_nonzero(diff > tol) <= 6e-3 * diff.shape[0], (mode, np.count_nonzero(diff > tol))
        assert gs["n_paths"] == cs["n_paths"] and abs(gs["n_rays"] - cs["n_rays"]) <= 2e-3 * cs["n_rays"]
    assert c[:, :3].max() > 6.0     # the light inside the nested list is visible


def _big_zoo(pkg, n_extra=4000):
    """The zoo above plus thousands of small spheres: the layouts no longer fit shared memory, so the kernels that can test
    complex leaves (QUADS) walk them from GLOBAL memory (32 registers per thread on that path, nested recursion on the
    local-memory stack)."""
    w = pkg.World.new()
    rng = np.random.default_rng(77)
    metal = pkg.material_spec(material=pkg.RTB_MAT_METAL, color=(0.8, 0.6, 0.2), fuzz=0.2)
    glass = pkg.material_spec(material=pkg.RTB_MAT_DIELECTRIC, ir=1.5)
    w.add_object(w.obj_translate(w.obj_sphere((0, 0, 0), 0.6, _grey(pkg), center2=(0.2, 0.1, 0)), (1.5, 0.5, -1.0)))
    w.add_object(w.obj_rotate_y(w.obj_quad((-1, -1, 0), (2, 0, 0), (0, 2, 0), _grey(pkg, (0.2, 0.8, 0.3))), 35.0))
    inner = w.obj_list([w.obj_sphere((0, 0, 0), 0.35, _grey(pkg, (0.9, 0.2, 0.2))), w.obj_box((0.4, -0.2, -0.2), (0.9, 0.3, 0.3), metal)])
    w.add_object(w.obj_translate(w.obj_rotate_y(w.obj_translate(inner, (0.3, 0, 0)), -50.0), (0.5, -1.2, 1.0)))
    w.add_object(w.obj_medium(w.obj_sphere((1.8, -0.8, 1.6), 0.8, glass), 1.5, (0.9, 0.9, 1.0)))
    w.add_box((-2.5, -1.0, -2.5), (-1.5, 0.2, -1.6), _grey(pkg), angle=25.0, offset=(0.1, 0.0, 0.2))
    for _ in range(n_extra):
        c = rng.uniform(-9, 9, 3)
        w.add_sphere(tuple(float(x) for x in c), float(rng.uniform(0.03, 0.12)), _grey(pkg, tuple(float(x) for x in rng.uniform(0.1, 0.9, 3))))
    w.add_sphere((0, -101.5, 0), 100.0, _grey(pkg, (0.5, 0.5, 0.5)))
    return w.build()


@pytest.mark.gpu
def test_complex_objects_on_the_global_memory_path(pkg, orc):
    world = _big_zoo(pkg)
    assert world.n_nodes * 32 > 96 * 1024          # one octant's layout does not fit the shared-memory staging limit
    scene = pkg.Scene(world)
    rng = np.random.default_rng(33)
    rays = np.concatenate([_rays(pkg, rng, 20000, -8, 8), _rays(pkg, rng, 6000, -2.5, 2.5)])
    cpu = orc.trace_rays(world.desc, rays)
    d = world.desc.contents
    is_medium = np.array([o >= 0 and d.hittables[o].type == pkg.RTB_HITTABLE_MEDIUM_OF for o in cpu["object"]])
    # the host's tree construction reorders the top-level objects (like the reference's constructTree sorts its slice); the
    # detached children of wrappers follow them
    types = np.array([d.hittables[i].type for i in range(d.n_hittables)])
    n_top = min(int(d.hittables[i].child) for i in range(d.n_hittables) if types[i] >= pkg.RTB_HITTABLE_TRANSLATE)
    complex_tops = [i for i in range(n_top) if types[i] != pkg.RTB_HITTABLE_SPHERE]
    assert len(complex_tops) == 5
    hit_objects = set(cpu["object"][cpu["object"] >= 0].tolist())
    assert set(complex_tops) <= hit_objects and is_medium.sum() > 200            # every complex object is hit by some ray
    for mode in (0, 2, 3):
        gpu = scene.trace_rays(rays, traversal=mode)
        same = gpu["object"] == cpu["object"]
        assert (~same).mean() < 2e-3, (mode, int((~same).sum()))
        solid = (cpu["object"] >= 0) & same & ~is_medium
        for k in ("front_face", "t", "p", "normal"):
            assert np.array_equal(gpu[k][solid], cpu[k][solid]), (mode, k)
    cam = pkg.Camera(image_width=160, aspect_ratio=1.0, samples_per_pixel=4, max_depth=20, vfov=60.0, lookfrom=(0.5, 1.0, 7.0),
                     lookat=(0, 0, 0), defocus_angle=0.0, background=(0.6, 0.7, 0.9)).init()
    ref = None
    for integrator, mode in ((0, 0), (1, 0), (1, 3)):
        o = pkg.render_options(seed=5, integrator=integrator, traversal=mode)
        g, _, _ = scene.render(cam, o)
        if ref is None:
            c, _, _ = orc.render(world.desc, cam, o, n_threads=8)
            ref = g
        diff = np.abs(g[:, :3] - c[:, :3]).max(axis=1)
        tol = 1e-4 * np.maximum(1.0, np.abs(c[:, :3]).max(axis=1))
        assert np.count_nonzero(diff > tol) <= 1e-2 * diff.shape[0], (integrator, mode, np.count_nonzero(diff > tol))
        if mode == 0:
            assert np.array_equal(g, ref)            # megakernel == wavefront in reference order, bit for bit
