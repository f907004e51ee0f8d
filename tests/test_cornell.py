"""SURVEY §8(f) rank 1, second half — createBox / HittableList / RotateY / Translate (src/objects.zig:264-443,
:510-532) as the reference composes them for cornellBox (src/main.zig:168-205, HEAD's selected scene :422):
Translate.init(RotateY.init(createBox(a, b, mat), angle), offset) lowered to ONE world object (RTB_HITTABLE_BOX)."""
import numpy as np
import pytest


def _box_world(pkg, a, b, angle=None, offset=None, **spec):
    w = pkg.World.new()
    w.add_box(a, b, pkg.material_spec(**spec), angle=angle, offset=offset)
    return w.build()


def _rays(pkg, rng, n, lo, hi):
    rays = np.zeros(n, dtype=np.dtype(pkg._ffi.RAY_DTYPE))
    rays["origin"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    rays["direction"] = rng.normal(size=(n, 3)).astype(np.float32)
    rays["time"] = rng.random(n).astype(np.float32)
    rays["t_min"] = 0.001
    rays["t_max"] = np.inf
    return rays


# ------------------------------------------------------------------------------------------------ CPU
def test_box_faces_and_reference_quirk(pkg, orc):
    """createBox adds the z = min face twice and NO z = max face (src/objects.zig:520-529): a ray entering through
    z = max sails through the missing face and hits the z = min face from the inside."""
    w = _box_world(pkg, (0, 0, 0), (1, 2, 3))
    hit = lambda o, d: orc.trace_rays(w.desc, orc.make_ray(o, d))[0]
    h = hit([0.5, 1, -5], [0, 0, 1])                       # the z = min side, from outside
    assert h["object"] == 0 and h["t"] == 5.0 and h["p"].tolist() == [0.5, 1, 0]
    assert abs(h["normal"][2]) == 1.0 and h["front_face"] in (0, 1)
    h = hit([0.5, 1, 8], [0, 0, -1])                       # towards z = max: there is no face there
    assert h["object"] == 0 and h["t"] == 8.0 and h["p"].tolist() == [0.5, 1, 0]
    h = hit([5, 1, 1.5], [-1, 0, 0])                       # x = max side
    assert h["t"] == 4.0 and h["normal"].tolist() == [1, 0, 0] and h["front_face"] == 1
    h = hit([0.5, 7, 1.5], [0, -1, 0])                     # top
    assert h["t"] == 5.0 and h["normal"].tolist() == [0, 1, 0]
    assert hit([5, 5, 5], [1, 1, 1])["object"] == -1
    # two coincident faces at z = min: the later one in the list (entry 2, u = -dx) wins the tie, so alpha runs
    # from x = max down to x = min
    h = hit([0.25, 1, -5], [0, 0, 1])
    assert h["u"] == pytest.approx(0.75) and h["v"] == pytest.approx(0.5)


def test_translate_and_rotate_semantics(pkg, orc):
    # Translate: the ray is moved by -offset, the hit point moved back by +offset (objects.zig:331-342)
    w = _box_world(pkg, (0, 0, 0), (1, 1, 1), offset=(10, 20, 30))
    h = orc.trace_rays(w.desc, orc.make_ray([10.5, 20.5, 25], [0, 0, 1]))[0]
    assert h["object"] == 0 and h["t"] == 5.0 and h["p"].tolist() == [10.5, 20.5, 30]
    # RotateY by 90 degrees: the box's +x face ends up facing -z... (cos*x + sin*z, -sin*x + cos*z) (objects.zig:425-439)
    w = _box_world(pkg, (0, 0, 0), (1, 1, 2), angle=90.0)
    d = w.desc.contents.hittables[0]
    assert d.sin_theta == pytest.approx(1.0) and abs(d.cos_theta) < 1e-6
    box = w.object_box(0)
    assert np.allclose(box, [0, 0, -1, 2, 1, 0], atol=1e-4)     # local (x, z) -> world (z, -x); HittableList's box holds 0
    h = orc.trace_rays(w.desc, orc.make_ray([1.0, 0.5, -5], [0, 0, 1]))[0]
    assert h["object"] == 0 and h["t"] == pytest.approx(4.0, abs=1e-5)
    assert np.allclose(h["normal"], [0, 0, -1], atol=1e-6) and h["front_face"] == 1


def test_box_bbox_matches_oracle(pkg, orc):
    rng = np.random.default_rng(2)
    for _ in range(20):
        a, b = rng.uniform(-50, 50, 3).astype(np.float32), rng.uniform(-50, 50, 3).astype(np.float32)
        angle = float(np.float32(rng.uniform(-180, 180)))
        off = rng.uniform(-300, 300, 3).astype(np.float32)
        w = _box_world(pkg, a, b, angle=angle, offset=off)
        h = w.desc.contents.hittables[0]
        mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
        orc.lib.orc_box_bbox.argtypes = None
        orc.lib.orc_box_bbox(a.ctypes.data_as(orc.vp), b.ctypes.data_as(orc.vp), orc.f32(h.sin_theta), orc.f32(h.cos_theta),
                             off.ctypes.data_as(orc.vp), mn.ctypes.data_as(orc.vp), mx.ctypes.data_as(orc.vp))
        assert np.array_equal(np.concatenate([mn, mx]), w.object_box(0))
        assert list(h.a) == a.tolist() and list(h.b) == b.tolist() and list(h.c) == off.tolist()


def test_cornell_world(pkg):
    w = pkg.World.create(pkg.RTW_SCENE_CORNELL_BOX)
    d = w.desc.contents
    assert d.n_hittables == 8 and d.n_nodes == 15
    types = sorted(d.hittables[i].type for i in range(8))
    assert types == [pkg.RTB_HITTABLE_QUAD] * 6 + [pkg.RTB_HITTABLE_BOX] * 2
    boxes = [d.hittables[i] for i in range(8) if d.hittables[i].type == pkg.RTB_HITTABLE_BOX]
    assert sorted(list(b.c) for b in boxes) == [[130, 0, 65], [265, 0, 295]]
    assert sorted(round(float(np.degrees(np.arcsin(b.sin_theta)))) for b in boxes) == [-18, 15]
    lights = [i for i in range(8) if d.materials[d.hittables[i].material].type == pkg.RTB_MAT_DIFFUSE_LIGHT]
    assert len(lights) == 1
    cam = pkg.cornell_camera().init()
    assert (cam.image_width, cam.image_height, cam.samples_per_pixel, cam.max_depth) == (600, 600, 200, 200)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_box_and_cornell_trace_parity(pkg, orc):
    rng = np.random.default_rng(4)
    worlds = [(_box_world(pkg, (0, 0, 0), (1, 2, 3)), (-4, 6)),
              (_box_world(pkg, (-1, -1, -1), (2, 1, 1), angle=33.0, offset=(0.5, -0.25, 1.0)), (-5, 6)),
              (pkg.World.create(pkg.RTW_SCENE_CORNELL_BOX), (-50, 600))]
    for world, (lo, hi) in worlds:
        scene = pkg.Scene(world)
        rays = _rays(pkg, rng, 30000, lo, hi)
        cpu = orc.trace_rays(world.desc, rays)
        assert (cpu["object"] >= 0).mean() > 0.01
        for mode in (0, 1, 2, 3):
            gpu = scene.trace_rays(rays, traversal=mode)
            same = gpu["object"] == cpu["object"]
            assert np.array_equal(gpu["t"], cpu["t"]), mode            # the nearest t is the same in every mode
            if mode == 0:
                assert same.all()                                         # reference order: the reference's object
            else:
                # the two boxes stand ON the floor quad: their bottom faces are coincident with it, and on an exact
                # tie a quad hit replaces the previous one (Interval.contains), so the winner follows the visiting
                # order.  Only rays coming from under the room see it; anything else must agree.
                assert (~same).mean() < 0.02
                assert (np.abs(cpu["p"][~same][:, 1]) < 1e-3).all()   # every disagreement lies in the floor plane y = 0
            hit = (cpu["object"] >= 0) & same
            for k in ("front_face", "t", "p", "normal", "u", "v"):
                assert np.array_equal(gpu[k][hit], cpu[k][hit]), (mode, k)
            if mode == 0:
                assert np.array_equal(gpu["n_box_tests"], cpu["n_box_tests"])
                assert np.array_equal(gpu["n_object_tests"], cpu["n_object_tests"])


@pytest.mark.gpu
@pytest.mark.parametrize("integrator", [0, 1])
def test_cornell_render_parity(pkg, orc, integrator):
    world = pkg.World.create(pkg.RTW_SCENE_CORNELL_BOX)
    scene = pkg.Scene(world)
    cam = pkg.cornell_camera(width=96, spp=8, max_depth=200).init()
    for mode in (0, 2, 3):
        o = pkg.render_options(seed=3, integrator=integrator, traversal=mode, flags=pkg.RTB_FLAG_COUNT_WORK)
        g, _, gs = scene.render(cam, o)
        c, _, cs = orc.render(world.desc, cam, o, n_threads=8)
        diff = np.abs(g[:, :3] - c[:, :3]).max(axis=1)
        tol = 1e-4 * np.maximum(1.0, np.abs(c[:, :3]).max(axis=1))
        assert np.count_nonzero(diff > tol) <= 5e-3 * diff.shape[0], (mode, np.count_nonzero(diff > tol))
        assert gs["n_paths"] == cs["n_paths"] and abs(gs["n_rays"] - cs["n_rays"]) <= 2e-3 * cs["n_rays"]
    assert c[:, :3].max() > 1.0                # the light is visible
    assert (c[:, :3].sum(axis=1) > 0).mean() > 0.1   # and bounces light the room (small light, 8 spp: sparse)


# ================================================================================================ volumes
# SURVEY §8(f) rank 2 — ConstantMedium (src/objects.zig:445-508) + Isotropic (src/material.zig:128-144), as used by
# cornellBoxSmoke (src/main.zig:207-251).  hit() is stochastic: it draws ONE random number (:484).  Both sides key that
# draw by (seed; pixel, sample, segment, block 0x40000000 + object) — for ray queries (seed 0; pixel = ray index).
def _medium_world(pkg, density, color=(1, 1, 1)):
    w = pkg.World.new()
    w.add_medium((0, 0, 0), (2, 2, 2), density, color)
    return w.build()


def test_constant_medium_free_path_statistics(pkg, orc):
    """P(hit) over a chord of length L is 1 - exp(-density * L); the hit lies inside the boundary; the record is
    the reference's arbitrary one (normal (1,0,0), front_face true)."""
    n = 6000
    # (along z there is nothing to find: createBox has no z = max face, so the second boundary hit never happens)
    wz = _medium_world(pkg, 5.0)   # (keep every World alive while its desc pointer is in use)
    assert (orc.trace_rays(wz.desc, np.repeat(orc.make_ray([1.0, 1.0, -3.0], [0, 0, 2.0]), 500))["object"] == -1).all()
    rays = np.repeat(orc.make_ray([-3.0, 1.0, 1.0], [2.0, 0, 0]), n)   # |d| = 2: enters at t = 1.5, leaves at t = 2.5
    for density in (0.2, 1.0, 5.0):
        wd = _medium_world(pkg, density)
        h = orc.trace_rays(wd.desc, rays)
        hit = h["object"] == 0
        expect = 1.0 - np.exp(-density * 2.0)                            # chord length 2 world units
        assert abs(hit.mean() - expect) < 4 * np.sqrt(expect * (1 - expect) / n) + 1e-3
        assert ((h["t"][hit] >= 1.5) & (h["t"][hit] <= 2.5)).all()
        assert (h["normal"][hit] == [1, 0, 0]).all() and (h["front_face"][hit] == 1).all()
        # exponential free path: mean depth of the hits that happen = 1/d - L e^{-dL} / (1 - e^{-dL})
        depth = (h["t"][hit] - 1.5) * 2.0
        mean = 1 / density - 2.0 * np.exp(-density * 2.0) / expect
        assert abs(depth.mean() - mean) < 0.08
    # ray_t clipping (objects.zig:475-478): a ray that starts inside only sees the remaining chord
    inside = np.repeat(orc.make_ray([1.5, 1.0, 1.0], [1.0, 0, 0]), n)
    wi = _medium_world(pkg, 1.0)
    h = orc.trace_rays(wi.desc, inside)
    assert abs((h["object"] == 0).mean() - (1 - np.exp(-0.5))) < 0.03


def test_isotropic_scatter(pkg, orc):
    w = _medium_world(pkg, 50.0, color=(0.2, 0.4, 0.6))
    ray = orc.make_ray([-3.0, 1.0, 1.0], [1.0, 0, 0])
    hit = orc.trace_rays(w.desc, ray)
    assert hit[0]["object"] == 0
    dirs = []
    for s in range(400):
        ok, att, sc = orc.scatter(w.desc, ray, hit, 5, 9, s, 1)
        assert ok and np.allclose(att, [0.2, 0.4, 0.6]) and np.array_equal(sc["origin"][0], hit[0]["p"])
        dirs.append(sc["direction"][0])
    dirs = np.array(dirs, np.float64)
    assert np.allclose(np.linalg.norm(dirs, axis=1), 1.0, atol=1e-5)          # randomUnitVector (material.zig:140)
    assert np.abs(dirs.mean(axis=0)).max() < 0.15                              # uniform over the sphere


def test_cornell_smoke_world(pkg):
    w = pkg.World.create(pkg.RTW_SCENE_CORNELL_SMOKE)
    d = w.desc.contents
    media = [d.hittables[i] for i in range(d.n_hittables) if d.hittables[i].type == pkg.RTB_HITTABLE_CONSTANT_MEDIUM]
    assert d.n_hittables == 8 and len(media) == 2
    assert all(m.radius == pytest.approx(-100.0) for m in media)              # neg_inv_density = -1 / 0.01
    assert all(d.materials[m.material].type == pkg.RTB_MAT_ISOTROPIC for m in media)
    cols = sorted(tuple(d.textures[d.materials[m.material].texture].color) for m in media)
    assert cols == [(0, 0, 0), (1, 1, 1)]


@pytest.mark.gpu
def test_constant_medium_trace_parity(pkg, orc):
    rng = np.random.default_rng(12)
    for world, (lo, hi) in [(_medium_world(pkg, 0.7), (-3, 5)), (pkg.World.create(pkg.RTW_SCENE_CORNELL_SMOKE), (-50, 600))]:
        scene = pkg.Scene(world)
        rays = _rays(pkg, rng, 30000, lo, hi)
        cpu = orc.trace_rays(world.desc, rays)
        d = world.desc.contents
        is_medium = np.array([o >= 0 and d.hittables[o].type == pkg.RTB_HITTABLE_CONSTANT_MEDIUM for o in cpu["object"]])
        assert is_medium.sum() > 300
        for mode in (0, 1, 2, 3):
            gpu = scene.trace_rays(rays, traversal=mode)
            same = gpu["object"] == cpu["object"]
            # log() differs by ulps between CUDA and glibc: a medium hit can flip only when the draw lands within an
            # ulp of the chord's end; coincident floor/box faces tie as in the solid Cornell box (modes 1, 2)
            assert (~same).mean() < (1e-3 if mode == 0 else 0.02)
            ok = same & (cpu["object"] >= 0)
            np.testing.assert_allclose(gpu["t"][ok], cpu["t"][ok], rtol=2e-5)
            solid = ok & ~is_medium
            assert np.array_equal(gpu["t"][solid], cpu["t"][solid])           # everything else stays bit-exact
            m = ok & is_medium
            assert (gpu["normal"][m] == [1, 0, 0]).all() and (gpu["front_face"][m] == 1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("integrator", [0, 1])
def test_cornell_smoke_render_parity(pkg, orc, integrator):
    world = pkg.World.create(pkg.RTW_SCENE_CORNELL_SMOKE)
    scene = pkg.Scene(world)
    cam = pkg.cornell_camera(width=96, spp=8, max_depth=50).init()
    for mode in (0, 2, 3):
        o = pkg.render_options(seed=3, integrator=integrator, traversal=mode, flags=pkg.RTB_FLAG_COUNT_WORK)
        g, _, gs = scene.render(cam, o)
        c, _, cs = orc.render(world.desc, cam, o, n_threads=8)
        diff = np.abs(g[:, :3] - c[:, :3]).max(axis=1)
        tol = 2e-4 * np.maximum(1.0, np.abs(c[:, :3]).max(axis=1))
        assert np.count_nonzero(diff > tol) <= 2e-2 * diff.shape[0], (mode, np.count_nonzero(diff > tol))
        assert gs["n_paths"] == cs["n_paths"] and abs(gs["n_rays"] - cs["n_rays"]) <= 5e-3 * cs["n_rays"]
    a, _, _ = scene.render(cam, pkg.render_options(seed=3, integrator=0))
    b, _, _ = scene.render(cam, pkg.render_options(seed=3, integrator=1))
    assert np.array_equal(a, b)      # megakernel == wavefront, media included


@pytest.mark.gpu
def test_far_camera_needs_the_per_ray_margin_of_sah16(pkg, orc):
    """the FMA slab test of RTB_TRAVERSAL_SAH relies on boxes padded by 2^-21 of the scene extent, a margin
    derived for ray origins near the scene; the thin (1e-4) slabs of axis-aligned quads are the first to drop out when
    the camera is far away.  RTB_TRAVERSAL_SAH16 carries its margin per ray (it grows with the origin's distance), so
    from 20 and 200 scene extents away it must still find exactly the reference's hits."""
    world = pkg.World.create(pkg.RTW_SCENE_CORNELL_BOX)
    scene = pkg.Scene(world)
    rng = np.random.default_rng(12)
    for dist in (1.0e4, 1.0e5):
        n = 20000
        rays = np.zeros(n, dtype=np.dtype(pkg._ffi.RAY_DTYPE))
        target = rng.uniform(0, 555, (n, 3)).astype(np.float32)
        origin = np.array([278.0, 278.0, -dist], np.float32) + rng.uniform(-50, 50, (n, 3)).astype(np.float32)
        rays["origin"], rays["direction"] = origin, target - origin
        rays["time"], rays["t_min"], rays["t_max"] = 0.0, 0.001, np.inf
        cpu = orc.trace_rays(world.desc, rays)
        assert (cpu["object"] >= 0).mean() > 0.9
        got = scene.trace_rays(rays, traversal=pkg.RTB_TRAVERSAL_SAH16)
        # coincident surfaces (box bottoms in the floor quad) may swap owner, everything else is the reference's hit
        same = got["object"] == cpu["object"]
        assert np.array_equal(got["t"], cpu["t"])
        assert (~same).mean() < 0.02 and (np.abs(cpu["p"][~same][:, 1]) < 1e-2).all()
        ref = scene.trace_rays(rays, traversal=pkg.RTB_TRAVERSAL_REFERENCE)
        assert np.array_equal(ref["object"], cpu["object"]) and np.array_equal(ref["t"], cpu["t"])
