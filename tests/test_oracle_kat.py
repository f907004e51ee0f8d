"""CPU tests: pin the oracle against every known-answer vector the reference source carries for the hot path
 plus analytic cases.  The reference's own tests pin nothing else ("parity unpinned")."""
import ctypes as C

import numpy as np
import pytest


# ---- RNG (replaces std.crypto.random, src/rtweekend.zig:14-16) ---------------------------------------------
def test_philox_known_answers(orc):
    # Random123 kat_vectors for philox4x32-10
    assert [hex(x) for x in orc.philox([0, 0, 0, 0], [0, 0])] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    assert [hex(x) for x in orc.philox([0xffffffff] * 4, [0xffffffff] * 2)] == \
        ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    assert [hex(x) for x in orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']


def test_u01_is_half_open_unit_interval(orc):
    # the reference's own test: `val > 0 and val < 1` (src/rtweekend.zig:30-33); ours may return exactly 0
    assert orc.lib.orc_u01(0) == 0.0
    assert orc.lib.orc_u01(0xFFFFFFFF) == pytest.approx(1.0 - 2.0 ** -24) and orc.lib.orc_u01(0xFFFFFFFF) < 1.0
    assert orc.lib.orc_u01(1 << 31) == 0.5


# ---- Aabb.hit: the three cases of the stale test at src/aabb.zig:117-136 ---------------------------------------
def test_aabb_hit_reference_cases(orc):
    lo, hi = [-1, -1, -1], [1, 1, 1]
    assert not orc.aabb_hit(lo, hi, orc.make_ray([13, 2, 3], [0, 0, 0]))       # zero direction: miss
    assert orc.aabb_hit(lo, hi, orc.make_ray([2, 2, 2], [-1, -1, -1]))
    assert orc.aabb_hit(lo, hi, orc.make_ray([1, 1, 1], [-1, -1, -1]))


def test_aabb_hit_edge_semantics(orc):
    lo, hi = [-1, -1, -1], [1, 1, 1]
    # `ray_t_max <= ray_t_min` rejects a zero-length overlap (src/aabb.zig:111)
    assert not orc.aabb_hit(lo, hi, orc.make_ray([-3, 1, 0], [1, 0, 0], t_min=2.0, t_max=2.0))
    # axis-parallel ray inside the slab: 0 * inf never produces a NaN that flips the result
    assert orc.aabb_hit(lo, hi, orc.make_ray([-3, 0.5, 0.5], [1, 0, 0]))
    assert not orc.aabb_hit(lo, hi, orc.make_ray([-3, 1.5, 0.5], [1, 0, 0]))
    # negative zero direction component: invD = -inf, the swap is keyed on invD < 0
    assert orc.aabb_hit(lo, hi, orc.make_ray([-3, 0.5, 0.5], [1, -0.0, 0]))
    # box behind the ray, and box beyond t_max
    assert not orc.aabb_hit(lo, hi, orc.make_ray([3, 0, 0], [1, 0, 0]))
    assert not orc.aabb_hit(lo, hi, orc.make_ray([-3, 0, 0], [1, 0, 0], t_max=1.5))


# ---- getSphereUV: the table in the comment at src/objects.zig:105-107 ---------------------------------------------
@pytest.mark.parametrize("p,uv", [((1, 0, 0), (0.50, 0.50)), ((-1, 0, 0), (0.00, 0.50)), ((0, 1, 0), (0.50, 1.00)),
                                  ((0, -1, 0), (0.50, 0.00)), ((0, 0, 1), (0.25, 0.50)), ((0, 0, -1), (0.75, 0.50))])
def test_sphere_uv_reference_table(orc, p, uv):
    u, v = orc.sphere_uv(p)
    if p == (-1, 0, 0):  # atan2(-0, -1) = -pi -> u = 0; atan2(+0, -1) = pi -> u = 1: same point on the seam
        assert min(abs(u - 0.0), abs(u - 1.0)) < 1e-6
    else:
        assert abs(u - uv[0]) < 1e-6
    assert abs(v - uv[1]) < 1e-6


# ---- Sphere.hit (src/objects.zig:116-148) -----------------------------------------------------------------------
def _one_sphere_world(pkg, center=(0, 0, 0), radius=1.0, center2=None, **spec):
    w = pkg.World.new()
    w.add_sphere(center, radius, pkg.material_spec(**spec), center2=center2)
    return w.build()


def test_sphere_hit_analytic(pkg, orc):
    w = _one_sphere_world(pkg)
    h = orc.trace_rays(w.desc, orc.make_ray([0, 0, -3], [0, 0, 1]))[0]
    assert h["object"] == 0 and h["front_face"] == 1 and h["t"] == 2.0
    assert h["p"].tolist() == [0, 0, -1] and h["normal"].tolist() == [0, 0, -1]
    # direction is NOT normalised in the reference: t scales with 1/|d|
    assert orc.trace_rays(w.desc, orc.make_ray([0, 0, -3], [0, 0, 2]))[0]["t"] == 1.0
    # from inside: near root <= t_min, far root accepted, back face, normal flipped towards the ray
    h = orc.trace_rays(w.desc, orc.make_ray([0, 0, 0], [0, 0, 1]))[0]
    assert h["object"] == 0 and h["front_face"] == 0 and h["t"] == 1.0 and h["normal"].tolist() == [0, 0, -1]
    # strict interval: a root exactly at t_max is rejected (Interval.surrounds, src/interval.zig:12-14)
    assert orc.trace_rays(w.desc, orc.make_ray([0, 0, -3], [0, 0, 1], t_max=2.0))[0]["object"] == -1
    assert orc.trace_rays(w.desc, orc.make_ray([0, 0, -3], [0, 0, 1], t_max=2.0001))[0]["object"] == 0
    # miss
    assert orc.trace_rays(w.desc, orc.make_ray([0, 2, -3], [0, 0, 1]))[0]["object"] == -1


def test_moving_sphere_center_is_linear_in_time(pkg, orc):
    w = _one_sphere_world(pkg, center=(0, 0, 0), radius=0.5, center2=(0, 2, 0))
    for time, y in [(0.0, 0.0), (0.5, 1.0), (1.0, 2.0)]:
        h = orc.trace_rays(w.desc, orc.make_ray([0, y, -3], [0, 0, 1], time=time))[0]
        assert h["object"] == 0 and h["t"] == 2.5
    assert orc.trace_rays(w.desc, orc.make_ray([0, 2, -3], [0, 0, 1], time=0.0))[0]["object"] == -1


def test_ties_go_to_dfs_earlier_object(pkg, orc):
    """`hit_record_right orelse hit_record_left` with a strict interval: equal t keeps the left (earlier) hit."""
    w = pkg.World.new()
    w.add_sphere((0, 0, 0), 1.0, pkg.material_spec())
    w.add_sphere((0, 0, 0), 1.0, pkg.material_spec(material=pkg.RTB_MAT_METAL))
    w.build()
    d = w.desc.contents
    root = d.nodes[d.root]
    left_obj = d.nodes[root.left].leaf
    h = orc.trace_rays(w.desc, orc.make_ray([0, 0, -3], [0, 0, 1]))[0]
    assert h["object"] == left_obj


# ---- materials (src/material.zig) ---------------------------------------------------------------------------------
def test_scatter_laws(pkg, orc):
    ray = orc.make_ray([0, 3, -3], [0, -1, 1])
    for kw, check in [
        (dict(material=pkg.RTB_MAT_LAMBERTIAN, color=(0.2, 0.4, 0.6)), "lambert"),
        (dict(material=pkg.RTB_MAT_METAL, color=(0.7, 0.6, 0.5), fuzz=0.0), "mirror"),
        (dict(material=pkg.RTB_MAT_DIELECTRIC, ir=1.5), "glass"),
        (dict(material=pkg.RTB_MAT_DIFFUSE_LIGHT, color=(4, 4, 4)), "light"),
    ]:
        w = _one_sphere_world(pkg, center=(0, -1000, 0), radius=1000.0, **kw)
        hit = orc.trace_rays(w.desc, ray)
        assert hit[0]["object"] == 0
        n = hit[0]["normal"].astype(np.float64)
        for sample in range(8):
            ok, att, sc = orc.scatter(w.desc, ray, hit, 99, 5, sample, 1)
            d = sc["direction"][0].astype(np.float64)
            assert np.array_equal(sc["origin"][0], hit[0]["p"]) and sc["time"][0] == ray["time"][0]
            if check == "lambert":   # direction = normal + unit vector (material.zig:44)
                assert ok and abs(np.linalg.norm(d - n) - 1.0) < 1e-5 and np.allclose(att, [0.2, 0.4, 0.6])
            elif check == "mirror":  # reflect(unit(d), n), fuzz 0 (material.zig:66-69)
                u = np.array([0, -1, 1]) / np.sqrt(2)
                assert ok and np.allclose(d, u - 2 * (u @ n) * n, atol=1e-6) and np.allclose(att, [0.7, 0.6, 0.5])
            elif check == "glass":   # attenuation 1, always scatters, unit-length result (material.zig:81-97)
                assert ok and np.allclose(att, 1.0) and abs(np.linalg.norm(d) - 1.0) < 1e-5
            else:                    # DiffuseLight never scatters (material.zig:119-121)
                assert not ok


def test_dielectric_total_internal_reflection(pkg, orc):
    # from inside glass (ir 1.5) at a grazing angle: ratio * sin_theta > 1 -> must reflect (material.zig:88-92)
    w = _one_sphere_world(pkg, radius=1.0, material=pkg.RTB_MAT_DIELECTRIC, ir=1.5)
    ray = orc.make_ray([0.9, 0, 0], [0.05, 1, 0])
    hit = orc.trace_rays(w.desc, ray)
    assert hit[0]["object"] == 0 and hit[0]["front_face"] == 0
    n = hit[0]["normal"].astype(np.float64)
    for sample in range(16):
        ok, _, sc = orc.scatter(w.desc, ray, hit, 1, 2, sample, 1)
        assert ok and (sc["direction"][0].astype(np.float64) @ n) > 0   # stays on the inside of the surface


# ---- textures (src/textures.zig, src/perlin.zig, src/rtw_image.zig) ---------------------------------------------
def test_checker_texture(pkg, orc):
    w = _one_sphere_world(pkg, texture=pkg.RTB_TEX_CHECKER, color=(0.2, 0.3, 0.1), color2=(0.9, 0.9, 0.9), scale=0.5)
    tex = w.desc.contents.materials[0].texture
    even, odd = [0.2, 0.3, 0.1], [0.9, 0.9, 0.9]
    assert np.allclose(orc.texture_value(w.desc, tex, 0, 0, [0.1, 0.1, 0.1]), even)
    assert np.allclose(orc.texture_value(w.desc, tex, 0, 0, [0.6, 0.1, 0.1]), odd)
    assert np.allclose(orc.texture_value(w.desc, tex, 0, 0, [0.6, 0.6, 0.1]), even)
    # negative odd sums: @rem gives -1 -> odd
    assert np.allclose(orc.texture_value(w.desc, tex, 0, 0, [-0.1, 0.1, 0.1]), odd)
    assert np.allclose(orc.texture_value(w.desc, tex, 0, 0, [-0.1, -0.1, 0.1]), even)


def test_image_texture_lookup(pkg, orc):
    img = np.zeros((4, 8, 4), np.uint8)
    img[..., 0] = np.arange(8)[None, :] * 10      # R encodes x
    img[..., 1] = np.arange(4)[:, None] * 20      # G encodes y
    img[..., 2] = 255
    w = pkg.World.new()
    w.add_image(img)
    w.add_sphere((0, 0, 0), 1.0, pkg.material_spec(texture=pkg.RTB_TEX_IMAGE, image_index=0))
    w.build()
    tex = w.desc.contents.materials[0].texture
    f = lambda u, v: np.round(orc.texture_value(w.desc, tex, u, v, [0, 0, 0]) * 255).astype(int).tolist()
    assert f(0.0, 1.0) == [0, 0, 255]            # v is flipped: v = 1 is row 0 (textures.zig:91)
    assert f(0.0, 0.0) == [0, 60, 255]           # v = 0 -> j = H -> clamped to H-1 (rtw_image.zig:37-45)
    assert f(1.0, 1.0) == [70, 0, 255]           # u = 1 -> i = W -> clamped to W-1
    assert f(0.5, 0.5) == [40, 40, 255]          # nearest texel, no filtering
    assert f(-3.0, 7.0) == [0, 0, 255]           # inputs clamped to [0,1] first


def test_perlin_noise_properties(pkg, orc):
    w = _one_sphere_world(pkg, texture=pkg.RTB_TEX_NOISE, scale=4.0, perlin_seed=3)
    d = w.desc.contents
    assert d.n_perlins == 1
    pl = d.perlins
    ranvec = np.array([list(pl[0].ranvec[i]) for i in range(256)])
    assert np.allclose(np.linalg.norm(ranvec, axis=1), 1.0, atol=1e-6)          # unitVector (perlin.zig:88)
    for perm in (pl[0].perm_x, pl[0].perm_y, pl[0].perm_z):
        assert sorted(perm[i] for i in range(256)) == list(range(256))           # a permutation of 0..255
    noise = lambda p: orc.lib.orc_perlin_noise(C.byref(pl[0]), np.asarray(p, np.float32).ctypes.data)
    turb = lambda p: orc.lib.orc_perlin_turb(C.byref(pl[0]), np.asarray(p, np.float32).ctypes.data, 7)
    # gradient noise vanishes on the lattice: every corner's weight vector is zero or its blend weight is
    for p in [(0, 0, 0), (1, 2, 3), (-4, 5, -6), (255, 256, 257)]:
        assert abs(noise(p)) < 1e-6
    rng = np.random.default_rng(0)
    pts = rng.uniform(-20, 20, (500, 3))
    vals = np.array([noise(p) for p in pts])
    assert np.abs(vals).max() <= 1.0 + 1e-5 and vals.std() > 0.1
    assert noise((0.3, 0.4, 0.5)) == pytest.approx(noise((256.3, 256.4, 256.5)), abs=1e-4)   # & 255 period
    assert all(turb(p) >= 0 for p in pts[:50])                                                 # @abs(accum)
    # NoiseTexture.value = 0.5 * (1 + sin(s.z + 10 turb(s))) in [0, 1], grey (textures.zig:118-123)
    tex = d.materials[0].texture
    c = orc.texture_value(w.desc, tex, 0, 0, [0.3, 0.2, 0.1])
    assert c[0] == c[1] == c[2] and 0.0 <= c[0] <= 1.0
    s = 4.0 * np.array([0.3, 0.2, 0.1], np.float32)
    assert c[0] == pytest.approx(0.5 * (1 + np.sin(s[2] + 10 * turb(s))), abs=1e-5)


# ---- camera (src/camera.zig:118-180) ----------------------------------------------------------------------------
def test_camera_init_invariants(pkg, orc):
    cam = orc.camera_init(pkg.Camera(image_width=1200, samples_per_pixel=500, max_depth=50))
    assert (cam.image_width, cam.image_height) == (1200, 675)     # round(1200 / (16/9)) (camera.zig:119-120)
    assert list(cam.center) == [13, 2, 3]
    du, dv = np.array(cam.pixel_delta_u[:]), np.array(cam.pixel_delta_v[:])
    vh = 2 * np.tan(np.radians(20) / 2) * 10
    assert np.linalg.norm(du) * 1200 == pytest.approx(vh * 1200 / 675, rel=1e-5)
    assert np.linalg.norm(dv) * 675 == pytest.approx(vh, rel=1e-5)
    assert abs(du @ dv) < 1e-9 and dv[1] < 0                     # viewport_v points down
    # the centre of the viewport lies focus_dist along -w from the camera
    centre = np.array(cam.pixel00_loc[:]) + du * 599.5 + dv * 337.0
    w = np.array([13, 2, 3]) / np.linalg.norm([13, 2, 3])
    assert np.allclose(centre, np.array([13, 2, 3]) - 10 * w, atol=2e-4)
    r = 10 * np.tan(np.radians(0.3))
    assert np.linalg.norm(cam.defocus_disk_u[:]) == pytest.approx(r, rel=1e-5)
    assert orc.camera_init(pkg.Camera(image_width=400)).image_height == 225
    assert orc.camera_init(pkg.Camera(image_width=1, aspect_ratio=100.0)).image_height == 1   # clamped to >= 1


def test_get_ray_uses_one_based_pixels_and_disk(pkg, orc):
    cam = orc.camera_init(pkg.book1_camera(400, 10, 50))
    rays = orc.get_rays(cam, 42, np.array([0, 399, 400 * 224 + 399]), 0)
    du, dv, p00 = (np.array(v[:], np.float64) for v in (cam.pixel_delta_u, cam.pixel_delta_v, cam.pixel00_loc))
    ddu, ddv = np.array(cam.defocus_disk_u[:], np.float64), np.array(cam.defocus_disk_v[:], np.float64)
    for ray, (x, y) in zip(rays, [(1, 1), (400, 1), (400, 225)]):   # x = i % W + 1, y = i / W + 1 (camera.zig:100-101)
        o, d = ray["origin"].astype(np.float64), ray["direction"].astype(np.float64)
        target = o + d
        off = target - (p00 + du * x + dv * y)
        a, b = off @ du / (du @ du), off @ dv / (dv @ dv)
        assert -0.5 - 1e-4 <= a <= 0.5 + 1e-4 and -0.5 - 1e-4 <= b <= 0.5 + 1e-4       # pixelSampleSquare
        lens = o - np.array([13, 2, 3])
        pu, pv = lens @ ddu / (ddu @ ddu), lens @ ddv / (ddv @ ddv)
        assert pu * pu + pv * pv < 1.0 + 1e-5                                          # randomInUnitDisk
        assert 0.0 <= ray["time"] < 1.0
    # defocus_angle <= 0: every ray starts at the camera centre (camera.zig:174)
    c0 = pkg.book1_camera(400, 1, 50)
    c0.defocus_angle = 0.0
    rays = orc.get_rays(orc.camera_init(c0), 42, np.arange(50), 3)
    assert (rays["origin"] == np.array([13, 2, 3], np.float32)).all()


# ---- integrator + writer (src/camera.zig:54-66, :93-116, :182-208; src/color.zig:43-62) ---------------------------
def test_ray_color_depth_and_background(pkg, orc):
    w = _one_sphere_world(pkg, center=(0, -1000, 0), radius=1000.0, color=(0.5, 0.5, 0.5))
    c = pkg.Camera(image_width=16, image_height=9, samples_per_pixel=1, max_depth=0, background=(0.1, 0.2, 0.3),
                   lookfrom=(0, 1, 5), lookat=(0, 1, 0), defocus_angle=0.0)
    rgb = np.zeros(3, np.float32)
    cam = orc.camera_init(c)
    orc.lib.orc_path_radiance(C.cast(w.desc, C.c_void_p), C.byref(cam), 1, 0, 0, rgb.ctypes.data)
    assert rgb.tolist() == [0, 0, 0]                       # depth <= 0 -> zero (camera.zig:183-185)
    c.max_depth = 5
    cam = orc.camera_init(c)
    orc.lib.orc_path_radiance(C.cast(w.desc, C.c_void_p), C.byref(cam), 1, 0, 0, rgb.ctypes.data)   # top row: sky
    assert np.allclose(rgb, [0.1, 0.2, 0.3])               # miss -> background (camera.zig:207)
    orc.lib.orc_path_radiance(C.cast(w.desc, C.c_void_p), C.byref(cam), 1, 16 * 8 + 8, 0, rgb.ctypes.data)
    assert (rgb >= 0).all() and (rgb <= np.array([0.1, 0.2, 0.3]) * 0.5 + 1e-6).all()   # >= one 0.5 bounce
    # black background + no light = black image (HEAD's default, camera.zig:80)
    c.background = (0, 0, 0)
    acc, _, _ = orc.render(w.desc, orc.camera_init(c), pkg.render_options(seed=1), n_threads=2)
    assert (acc[:, :3] == 0).all() and (acc[:, 3] == 1).all()


def test_writer_accumulates_sum_and_sample_count(pkg, orc):
    w = pkg.World.book1()
    cam = orc.camera_init(pkg.book1_camera(64, 6, 10))
    full, rgba, st = orc.render(w.desc, cam, pkg.render_options(seed=7), n_threads=8)
    assert (full[:, 3] == 6).all() and st["n_paths"] == 64 * 36 * 6
    # buffer holds the SUM (camera.zig:55); resumable: 0..2 then 2..6 gives the same bits
    part, _, _ = orc.render(w.desc, cam, pkg.render_options(seed=7, sample_count=2), n_threads=8)
    assert (part[:, 3] == 2).all()
    part, rgba2, _ = orc.render(w.desc, cam, pkg.render_options(seed=7, sample_begin=2, sample_count=4), n_threads=3,
                                accum=part)
    assert np.array_equal(part, full) and np.array_equal(rgba, rgba2)
    # thread count / strips do not change the result (per-pixel streams)
    one, _, _ = orc.render(w.desc, cam, pkg.render_options(seed=7), n_threads=1)
    assert np.array_equal(one, full)
    # Task{thread_idx, chunk_size} strips
    n = 64 * 36
    strips = np.zeros((n, 4), np.float32)
    for t in range(8):
        strips, _, _ = orc.render(w.desc, cam, pkg.render_options(seed=7, pixel_begin=t * (n // 8), pixel_count=n // 8),
                                  n_threads=1, accum=strips)
    assert np.array_equal(strips, full)
    assert np.array_equal(rgba, orc.resolve(full))


def test_resolve_gamma_and_truncation(orc):
    acc = np.array([[4, 4, 4, 4], [1, 1, 1, 4], [0, 0, 0, 1], [9, 9, 9, 1], [0.999 ** 2, 0.5, 0.25, 1]], np.float32)
    out = orc.resolve(acc)
    assert out[0].tolist() == [255, 255, 255, 255]          # sqrt(1) clamped to 0.999 -> trunc(255.744)
    assert out[1].tolist() == [128, 128, 128, 255]          # sqrt(0.25) * 256
    assert out[2].tolist() == [0, 0, 0, 255]
    assert out[3].tolist() == [255, 255, 255, 255]
    assert out[4].tolist()[1:] == [int(256 * np.sqrt(np.float32(0.5))), 128, 255]
    assert orc.resolve(acc, 16.0)[0].tolist() == [128, 128, 128, 255]   # override n


# ---- host-side producers restated in the oracle (src/bvh.zig:43-103, src/rtweekend.zig:23-27) -----------------------
def test_random_int_range_quirk(orc):
    s = C.c_uint64(5)
    vals = [orc.lib.orc_host_random_int_range(C.byref(s), 0, 2) for _ in range(4000)]
    assert set(vals) == {0, 1, 2, 3}                          # round(0 + 3 r) reaches max + 1
    counts = np.bincount(vals)
    assert counts[0] < counts[1] and counts[3] < counts[2]     # 0 and 3 get half-width buckets


def test_oracle_bvh_is_a_valid_reference_tree(pkg, orc):
    ffi = pkg._ffi
    rng = np.random.default_rng(1)
    n = 37
    hs = (ffi.RtbHittable * n)()
    boxes = np.zeros((n, 6), np.float32)
    for i in range(n):
        c = rng.uniform(-5, 5, 3).astype(np.float32)
        hs[i].type, hs[i].material, hs[i].radius = ffi.RTB_HITTABLE_SPHERE, i, float(rng.uniform(0.1, 1))
        hs[i].a = (ffi.f32 * 3)(*c)
        orc.lib.orc_sphere_bbox(c.ctypes.data, None, hs[i].radius, boxes[i, :3].ctypes.data, boxes[i, 3:].ctypes.data)
        assert np.allclose(boxes[i, :3], c - hs[i].radius) and np.allclose(boxes[i, 3:], c + hs[i].radius)
    nodes = (ffi.RtbBvhNode * (2 * n - 1))()
    seed = C.c_uint64(2)
    root = orc.lib.orc_bvh_build(hs, boxes.ctypes.data, n, C.byref(seed), nodes)
    assert root == 2 * n - 2                                   # post-order: the root is created last
    assert sorted(h.material for h in hs) == list(range(n))    # a permutation of the input objects
    seen = []

    def walk(i, depth):
        nd = nodes[i]
        if nd.leaf >= 0:
            seen.append(nd.leaf)
            assert np.array_equal(np.array(nd.bmin[:] + nd.bmax[:], np.float32), boxes[nd.leaf])
            return 1, depth
        l, r = nodes[nd.left], nodes[nd.right]
        assert np.array_equal(np.minimum(l.bmin[:], r.bmin[:]).astype(np.float32), np.array(nd.bmin[:], np.float32))
        assert np.array_equal(np.maximum(l.bmax[:], r.bmax[:]).astype(np.float32), np.array(nd.bmax[:], np.float32))
        nl, dl = walk(nd.left, depth + 1)
        nr, dr = walk(nd.right, depth + 1)
        assert nl == (nl + nr) // 2                            # median split: left gets floor(span / 2)
        return nl + nr, max(dl, dr)

    total, depth = walk(root, 1)
    assert total == n and sorted(seen) == list(range(n)) and depth == 7   # ceil(log2 37) + 1 levels
