"""SURVEY §8(f) rank 1 — Quad (src/objects.zig:195-262) and DiffuseLight (src/material.zig:108-126) with the scenes
that use them (quadsWorld / simpleLightWorld, src/main.zig:127-166): oracle known-answer tests on CPU, GPU parity
through the C ABI."""
import ctypes as C

import numpy as np
import pytest


def _quad_world(pkg, q, u, v, **spec):
    w = pkg.World.new()
    w.add_quad(q, u, v, pkg.material_spec(**spec))
    return w.build()


# ------------------------------------------------------------------------------------------------ CPU: oracle
def test_quad_hit_analytic(pkg, orc):
    # unit square in the z = 0 plane, normal = unit(cross(u, v)) = +z
    w = _quad_world(pkg, (0, 0, 0), (1, 0, 0), (0, 1, 0))
    h = orc.trace_rays(w.desc, orc.make_ray([0.25, 0.75, 2], [0, 0, -1]))[0]
    assert h["object"] == 0 and h["t"] == 2.0 and h["front_face"] == 1
    assert h["normal"].tolist() == [0, 0, 1] and h["p"].tolist() == [0.25, 0.75, 0]
    assert (h["u"], h["v"]) == (0.25, 0.75)                   # alpha, beta (isInterior, objects.zig:217-224)
    # from behind: same plane, back face, normal flipped against the ray
    h = orc.trace_rays(w.desc, orc.make_ray([0.25, 0.75, -2], [0, 0, 1]))[0]
    assert h["object"] == 0 and h["front_face"] == 0 and h["normal"].tolist() == [0, 0, -1]
    # outside the parallelogram, parallel to the plane (|denom| < 1e-8), behind the origin
    assert orc.trace_rays(w.desc, orc.make_ray([1.5, 0.5, 2], [0, 0, -1]))[0]["object"] == -1
    assert orc.trace_rays(w.desc, orc.make_ray([0.5, 0.5, 2], [1, 0, 0]))[0]["object"] == -1
    assert orc.trace_rays(w.desc, orc.make_ray([0.5, 0.5, 2], [0, 0, 1]))[0]["object"] == -1
    # edges are inside (a < 0 or 1 < a rejects; 0 and 1 pass)
    assert orc.trace_rays(w.desc, orc.make_ray([1.0, 0.0, 2], [0, 0, -1]))[0]["object"] == 0
    # Interval.contains is INCLUSIVE for quads (objects.zig:237), unlike the sphere's surrounds
    assert orc.trace_rays(w.desc, orc.make_ray([0.5, 0.5, 2], [0, 0, -1], t_max=2.0))[0]["object"] == 0
    assert orc.trace_rays(w.desc, orc.make_ray([0.5, 0.5, 2], [0, 0, -1], t_max=1.999))[0]["object"] == -1


def test_quad_bbox_is_padded(pkg, orc):
    # an axis-aligned quad has zero thickness: Aabb.pad widens it by delta = 1e-4 (aabb.zig:36-43)
    w = _quad_world(pkg, (0, 0, 0), (1, 0, 0), (0, 1, 0))
    box = w.object_box(0)
    assert box[2] == pytest.approx(-0.00005) and box[5] == pytest.approx(0.00005)
    assert box[0] == 0 and box[3] == 1 and box[1] == 0 and box[4] == 1
    mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
    q, u, v = (np.array(x, np.float32) for x in ((0, 0, 0), (1, 0, 0), (0, 1, 0)))
    orc.lib.orc_quad_bbox(q.ctypes.data, u.ctypes.data, v.ctypes.data, mn.ctypes.data, mx.ctypes.data)
    assert np.array_equal(np.concatenate([mn, mx]), box)


def test_diffuse_light_emits_and_stops(pkg, orc):
    # a light quad facing the camera on a black background: radiance = emit colour, path ends (material.zig:119-125)
    w = _quad_world(pkg, (-5, -5, 0), (10, 0, 0), (0, 10, 0), material=pkg.RTB_MAT_DIFFUSE_LIGHT, color=(4, 3, 2))
    cam = orc.camera_init(pkg.Camera(image_width=8, image_height=8, samples_per_pixel=1, max_depth=5,
                                     lookfrom=(0, 0, 3), lookat=(0, 0, 0), defocus_angle=0.0, vfov=40.0))
    acc, _, st = orc.render(w.desc, cam, pkg.render_options(seed=1), n_threads=1)
    assert np.allclose(acc[:, :3], [4, 3, 2]) and st["n_rays"] == 64 and st["n_hits"] == 64


def test_reference_quad_scenes_build(pkg):
    wq = pkg.World.create(pkg.RTW_SCENE_QUADS)
    d = wq.desc.contents
    assert d.n_hittables == 5 and all(d.hittables[i].type == pkg.RTB_HITTABLE_QUAD for i in range(5))
    wl = pkg.World.create(pkg.RTW_SCENE_SIMPLE_LIGHT)
    d = wl.desc.contents
    assert d.n_hittables == 4
    lights = [i for i in range(4) if d.materials[d.hittables[i].material].type == pkg.RTB_MAT_DIFFUSE_LIGHT]
    assert len(lights) == 2 and d.n_perlins == 1


# ------------------------------------------------------------------------------------------------ GPU: parity
def _rays(pkg, rng, n, lo, hi):
    rays = np.zeros(n, dtype=np.dtype(pkg._ffi.RAY_DTYPE))
    rays["origin"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    rays["direction"] = rng.normal(size=(n, 3)).astype(np.float32)
    rays["time"] = rng.random(n).astype(np.float32)
    rays["t_min"] = 0.001
    rays["t_max"] = np.inf
    return rays


def _five_quads(pkg):
    """The book's five-quad box (the reference's quadsWorld with upper_orange where the book puts it, y = +3;
    in src/main.zig:137-138 upper_orange and lower_teal are COINCIDENT at y = -3, see the tie test below)."""
    w = pkg.World.new()
    for q, u, v, c in [((-3, -2, 5), (0, 0, -4), (0, 4, 0), (1, 0.2, 0.2)), ((-2, -2, 0), (4, 0, 0), (0, 4, 0), (0.2, 1, 0.2)),
                       ((3, -2, 1), (0, 0, 4), (0, 4, 0), (0.2, 0.2, 1)), ((-2, 3, 1), (4, 0, 0), (0, 0, 4), (1, 0.5, 0)),
                       ((-2, -3, 5), (4, 0, 0), (0, 0, -4), (0.2, 0.8, 0.8))]:
        w.add_quad(q, u, v, pkg.material_spec(color=c))
    return w.build()


def _world(pkg, kind):
    if kind == "quads":
        return _five_quads(pkg)
    if kind == "quads_ref":
        return pkg.World.create(pkg.RTW_SCENE_QUADS)
    return pkg.World.create(pkg.RTW_SCENE_SIMPLE_LIGHT)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["quads", "light"])
def test_quad_scenes_trace_parity(pkg, orc, kind):
    world = _world(pkg, kind)
    scene = pkg.Scene(world)
    rng = np.random.default_rng(7)
    rays = _rays(pkg, rng, 30000, -6, 8)
    cpu = orc.trace_rays(world.desc, rays)
    assert (cpu["object"] >= 0).mean() > 0.05
    for mode in (0, 1, 2, 3):
        gpu = scene.trace_rays(rays, traversal=mode)
        assert np.array_equal(gpu["object"], cpu["object"]) and np.array_equal(gpu["front_face"], cpu["front_face"])
        hit = cpu["object"] >= 0
        for k in ("t", "p", "normal"):
            assert np.array_equal(gpu[k][hit], cpu[k][hit]), (mode, k)
        isq = hit & np.array([world.desc.contents.hittables[max(o, 0)].type == pkg.RTB_HITTABLE_QUAD
                              for o in cpu["object"]])
        assert np.array_equal(gpu["u"][isq], cpu["u"][isq]) and np.array_equal(gpu["v"][isq], cpu["v"][isq])
        if mode == 0:
            assert np.array_equal(gpu["n_box_tests"], cpu["n_box_tests"])
            assert np.array_equal(gpu["n_object_tests"], cpu["n_object_tests"])


@pytest.mark.gpu
def test_coincident_quads_tie_follows_visiting_order(pkg, orc):
    """The reference's quadsWorld holds two coincident quads (upper_orange / lower_teal, src/main.zig:137-138).
    Quad.hit accepts t == ray_t.max (Interval.contains, src/objects.zig:237), so on that surface the LAST quad
    visited wins.  RTB_TRAVERSAL_REFERENCE reproduces the reference's choice bit for bit; ORDERED / SAH find
    the same t everywhere and may only disagree about which of the two coincident quads they report."""
    world = _world(pkg, "quads_ref")
    scene = pkg.Scene(world)
    d = world.desc.contents
    pair = {i for i in range(5) if d.hittables[i].a[1] == -3.0}
    assert len(pair) == 2
    rng = np.random.default_rng(8)
    rays = _rays(pkg, rng, 40000, -6, 8)
    cpu = orc.trace_rays(world.desc, rays)
    gpu = scene.trace_rays(rays, traversal=0)
    assert np.array_equal(gpu["object"], cpu["object"]) and np.array_equal(gpu["t"], cpu["t"])
    on_pair = np.isin(cpu["object"], list(pair))
    assert on_pair.sum() > 200 and len(set(cpu["object"][on_pair].tolist())) == 1   # the reference always picks one
    for mode in (1, 2, 3):
        g = scene.trace_rays(rays, traversal=mode)
        assert np.array_equal(g["t"], cpu["t"])
        diff = g["object"] != cpu["object"]
        assert np.isin(g["object"][diff], list(pair)).all() and np.isin(cpu["object"][diff], list(pair)).all()
    # the full render in reference order matches the oracle on the reference's own scene
    cam = pkg.Camera(image_width=64, aspect_ratio=1.0, samples_per_pixel=4, max_depth=20, vfov=80.0, lookfrom=(0, 0, 9),
                     lookat=(0, 0, 0), defocus_angle=0.0, background=(0.70, 0.80, 1.00)).init()
    for integrator in (0, 1):
        o = pkg.render_options(seed=5, integrator=integrator)
        g, _, _ = scene.render(cam, o)
        c, _, _ = orc.render(world.desc, cam, o, n_threads=8)
        assert np.count_nonzero(np.abs(g[:, :3] - c[:, :3]).max(axis=1) > 1e-4) <= 5e-3 * g.shape[0]


@pytest.mark.gpu
@pytest.mark.parametrize("integrator", [0, 1])
@pytest.mark.parametrize("kind", ["quads", "light"])
def test_quad_scenes_render_parity(pkg, orc, kind, integrator):
    world = _world(pkg, kind)
    if kind == "quads":   # framed as in the book (the reference sets no camera fields for quadsWorld)
        camo = pkg.Camera(image_width=96, aspect_ratio=1.0, samples_per_pixel=4, max_depth=50, vfov=80.0,
                          lookfrom=(0, 0, 9), lookat=(0, 0, 0), defocus_angle=0.0, background=(0.70, 0.80, 1.00))
    else:                 # simpleLightWorld's camera (main.zig:156-161), black background: only the lights emit
        camo = pkg.simple_light_camera(128, 8, 50)
    scene = pkg.Scene(world)
    cam = camo.init()
    spp = cam.samples_per_pixel
    for mode in (0, 2, 3):
        o = pkg.render_options(seed=77, integrator=integrator, traversal=mode, flags=pkg.RTB_FLAG_COUNT_WORK)
        g, g_rgba, gs = scene.render(cam, o)
        c, c_rgba, cs = orc.render(world.desc, cam, o, n_threads=8)
        diff = np.abs(g[:, :3] - c[:, :3]).max(axis=1)
        tol = 2e-5 * spp * np.maximum(1.0, np.abs(c[:, :3]).max(axis=1) / spp)
        assert np.count_nonzero(diff > tol) <= 5e-3 * diff.shape[0], (kind, mode)
        assert gs["n_paths"] == cs["n_paths"] and abs(gs["n_rays"] - cs["n_rays"]) <= 1e-3 * cs["n_rays"]
        if mode == 0:
            assert abs(gs["n_box_tests"] - cs["n_box_tests"]) <= 1e-3 * cs["n_box_tests"]
    if kind == "light":
        assert c[:, :3].max() > 1.0 and (c[:, :3] == 0).any()   # emitters above 1, unlit pixels exactly black
