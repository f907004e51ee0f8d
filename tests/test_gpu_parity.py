"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against the CPU oracle.

Bars (BASELINE.json north_star):
  * identical ray batches -> nearest-hit object index and front_face bit-exact; t, normal, uv within 1e-5 relative;
  * the same Philox streams on both sides -> images agree pixel by pixel up to float association / libm ulps;
  * different seeds -> RMSE and mean bias within the oracle's own seed-to-seed spread.
"""
import ctypes as C
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REL = 1e-5  # stated tolerance for t / normal / uv (BASELINE.json)


def _rays_from_camera(orc, cam, seed, n, rng):
    size = cam.image_width * cam.image_height
    pixels = rng.integers(0, size, n).astype(np.uint32)
    samples = rng.integers(0, 64, n).astype(np.uint32)
    return orc.get_rays(cam, seed, pixels, samples)


def _random_rays(pkg, rng, n, lo, hi):
    rays = np.zeros(n, dtype=np.dtype(pkg._ffi.RAY_DTYPE))
    rays["origin"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    rays["direction"] = rng.normal(size=(n, 3)).astype(np.float32)
    rays["time"] = rng.random(n).astype(np.float32)
    rays["t_min"] = 0.001
    rays["t_max"] = np.inf
    return rays


def _secondary_rays(pkg, orc, desc, rays, hits, seed):
    """Scatter every hit with the oracle to get realistic incoherent bounce rays."""
    out = []
    for i in np.nonzero(hits["object"] >= 0)[0]:
        ok, _, sc = orc.scatter(desc, rays[i:i + 1], hits[i:i + 1], seed, int(i), 0, 1)
        if ok:
            out.append(sc)
    return np.concatenate(out) if out else rays[:0]


def _assert_hits_equal(gpu, cpu):
    assert np.array_equal(gpu["object"], cpu["object"]), \
        f"nearest-hit index differs on {np.count_nonzero(gpu['object'] != cpu['object'])} rays"
    assert np.array_equal(gpu["front_face"], cpu["front_face"])
    # the traversal itself must be the reference's: same number of slab and primitive tests per ray
    assert np.array_equal(gpu["n_box_tests"], cpu["n_box_tests"])
    assert np.array_equal(gpu["n_object_tests"], cpu["n_object_tests"])
    hit = cpu["object"] >= 0
    # decision arithmetic is IEEE-exact on both sides: t, p, normal are bit-identical, not merely close
    assert np.array_equal(gpu["t"][hit], cpu["t"][hit])
    assert np.array_equal(gpu["p"][hit], cpu["p"][hit])
    assert np.array_equal(gpu["normal"][hit], cpu["normal"][hit])
    for k in ("u", "v"):  # acos/atan2: CUDA libm vs glibc
        np.testing.assert_allclose(gpu[k][hit], cpu[k][hit], rtol=REL, atol=REL)


def test_philox_device_matches_oracle(pkg, orc):
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2**32, (4096, 4), dtype=np.uint64).astype(np.uint32)
    ctr[0] = 0
    ctr[1] = 0xFFFFFFFF
    key = np.array([0x12345678, 0x9ABCDEF0], dtype=np.uint32)
    got = pkg.philox_device(ctr, key)
    want = np.stack([orc.philox(c, key) for c in ctr])
    assert np.array_equal(got, want)
    assert np.array_equal(pkg.philox_device(np.zeros((1, 4), np.uint32), [0, 0])[0],
                          np.array([0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8], dtype=np.uint32))


@pytest.fixture(scope="module")
def book1(pkg):
    world = pkg.World.book1()
    return world, pkg.Scene(world)


def test_trace_book1_camera_and_bounce_rays(pkg, orc, book1):
    world, scene = book1
    cam = pkg.book1_camera(400, 10, 50).init()
    rng = np.random.default_rng(1)
    rays = _rays_from_camera(orc, cam, 77, 20000, rng)
    cpu = orc.trace_rays(world.desc, rays)
    gpu = scene.trace_rays(rays)
    _assert_hits_equal(gpu, cpu)
    assert (cpu["object"] >= 0).mean() > 0.5
    bounce = _secondary_rays(pkg, orc, world.desc, rays[:6000], cpu[:6000], 77)
    assert bounce.shape[0] > 1000
    _assert_hits_equal(scene.trace_rays(bounce), orc.trace_rays(world.desc, bounce))


def test_trace_book1_random_and_degenerate_rays(pkg, orc, book1):
    world, scene = book1
    rng = np.random.default_rng(2)
    rays = _random_rays(pkg, rng, 20000, -12, 12)
    rays["origin"][:, 1] = np.abs(rays["origin"][:, 1]) * 0.3
    # zero direction components: invD = +-inf, 0*inf = NaN must fall through the comparisons (aabb.zig:87-111)
    rays["direction"][:3000, 0] = 0.0
    rays["direction"][3000:6000, 1] = 0.0
    rays["direction"][6000:8000, 2] = -0.0
    rays["direction"][8000:8500] = 0.0            # the stale aabb test's zero-direction ray
    rays["t_max"][9000:12000] = rng.uniform(0.5, 20.0, 3000).astype(np.float32)   # finite ray_t.max
    rays["t_min"][12000:13000] = 0.0
    # rays starting inside / on spheres (far-root branch, objects.zig:131-134)
    d = world.desc.contents
    for k in range(13000, 16000):
        h = d.hittables[int(rng.integers(0, d.n_hittables))]
        rays["origin"][k] = np.array(h.a[:]) + rng.normal(size=3) * 0.3 * h.radius
        rays["time"][k] = 0.0
    _assert_hits_equal(scene.trace_rays(rays), orc.trace_rays(world.desc, rays))


def test_trace_edge_worlds(pkg, orc):
    rng = np.random.default_rng(3)
    spec = pkg.material_spec()
    # single sphere: the root is a leaf (BVHTree.constructTree span == 1)
    w1 = pkg.World.new()
    w1.add_sphere((0, 0, -1), 0.5, spec)
    w1.build()
    rays = _random_rays(pkg, rng, 4000, -2, 2)
    _assert_hits_equal(pkg.Scene(w1).trace_rays(rays), orc.trace_rays(w1.desc, rays))
    # two and three objects incl. a moving sphere and coincident spheres (exact t ties -> DFS-earlier wins)
    w3 = pkg.World.new()
    w3.add_sphere((0, 0, 0), 1.0, spec)
    w3.add_sphere((0, 0, 0), 1.0, pkg.material_spec(material=pkg.RTB_MAT_METAL))
    w3.add_sphere((0.5, 0, 0), 0.7, spec, center2=(0.5, 0.5, 0))
    w3.build()
    rays = _random_rays(pkg, rng, 8000, -3, 3)
    cpu = orc.trace_rays(w3.desc, rays)
    _assert_hits_equal(pkg.Scene(w3).trace_rays(rays), cpu)
    assert len(set(cpu["object"].tolist())) >= 3
    # the UV known-answer table of objects.zig:105-107 through the whole GPU path
    wu = pkg.World.new()
    wu.add_sphere((0, 0, 0), 1.0, spec)
    wu.build()
    table = {(1, 0, 0): (0.5, 0.5), (-1, 0, 0): (0.0, 0.5), (0, 1, 0): (0.5, 1.0), (0, -1, 0): (0.5, 0.0),
             (0, 0, 1): (0.25, 0.5), (0, 0, -1): (0.75, 0.5)}
    rays = np.concatenate([orc.make_ray(np.array(p, np.float32) * 3, -np.array(p, np.float32)) for p in table])
    hits = pkg.Scene(wu).trace_rays(rays)
    for h, (p, (u, v)) in zip(hits, table.items()):
        assert h["object"] == 0 and h["front_face"] == 1
        if p == (-1, 0, 0):  # u wraps: atan2(-0, -1) + pi is 0 or 2pi depending on the zero's sign
            assert min(abs(h["u"] - 0.0), abs(h["u"] - 1.0)) < 1e-6
        else:
            assert abs(h["u"] - u) < 1e-6
        assert abs(h["v"] - v) < 1e-6
    # zero rays
    assert pkg.Scene(wu).trace_rays(rays[:0]).shape[0] == 0


def _render_pair(pkg, orc, world, scene, cam, seed=1234, spp=None, **opt):
    o = pkg.render_options(seed=seed, sample_count=spp or 0, flags=pkg.RTB_FLAG_COUNT_WORK, **opt)
    g_acc, g_rgba, g_st = scene.render(cam, o)
    c_acc, c_rgba, c_st = orc.render(world.desc, cam, o, n_threads=8)
    return (g_acc, g_rgba, g_st), (c_acc, c_rgba, c_st)


def _assert_same_paths(g, c, spp, max_bad_frac=2e-3):
    (g_acc, g_rgba, g_st), (c_acc, c_rgba, c_st) = g, c
    # same streams, same decisions -> the work counters agree exactly unless a libm ulp flipped a path
    for k in ("n_paths", "n_rays", "n_box_tests", "n_object_tests", "n_hits"):
        assert abs(g_st[k] - c_st[k]) <= 2e-4 * max(1, c_st[k]), (k, g_st[k], c_st[k])
    assert np.array_equal(g_acc[:, 3], c_acc[:, 3])
    diff = np.abs(g_acc[:, :3] - c_acc[:, :3]).max(axis=1)
    tol = 2e-5 * spp * np.maximum(1.0, np.abs(c_acc[:, :3]).max(axis=1) / spp)
    bad = np.count_nonzero(diff > tol)
    assert bad <= max_bad_frac * diff.shape[0], f"{bad} of {diff.shape[0]} pixels differ beyond float association"
    d8 = np.abs(g_rgba.astype(np.int32) - c_rgba.astype(np.int32))
    assert np.count_nonzero(d8 > 1) <= max_bad_frac * d8.shape[0] * 4


def test_render_book1_same_streams_as_oracle(pkg, orc, book1):
    world, scene = book1
    cam = pkg.book1_camera(256, 4, 50).init()
    g, c = _render_pair(pkg, orc, world, scene, cam)
    _assert_same_paths(g, c, 4)
    assert g[2]["n_launches"] >= 1 and g[2]["device_ms"] > 0


def test_render_textured_same_streams_as_oracle(pkg, orc, earthmap):
    world = pkg.World.create(pkg.RTW_SCENE_TEXTURED, image=earthmap)
    scene = pkg.Scene(world)
    cam = pkg.textured_camera(192, 4, 50).init()
    g, c = _render_pair(pkg, orc, world, scene, cam)
    # image texels flip on acos/atan2 ulps and perlin's sin differs by ulps: allow more (still tiny) drift
    _assert_same_paths(g, c, 4, max_bad_frac=1e-2)
    # HEAD's Book-1 variant: checker ground + earth sphere (main.zig:257-303)
    world2 = pkg.World.book1(checker_ground=True, earth=True, image=earthmap)
    g, c = _render_pair(pkg, orc, world2, pkg.Scene(world2), pkg.book1_camera(160, 2, 50).init())
    _assert_same_paths(g, c, 2, max_bad_frac=1e-2)


def test_wavefront_equals_megakernel_bit_exact(pkg, book1, earthmap):
    world, scene = book1
    cam = pkg.book1_camera(200, 6, 50).init()
    a, ra, sa = scene.render(cam, pkg.render_options(seed=5, flags=pkg.RTB_FLAG_COUNT_WORK))
    b, rb, sb = scene.render(cam, pkg.render_options(seed=5, flags=pkg.RTB_FLAG_COUNT_WORK,
                                                     integrator=pkg.RTB_INTEGRATOR_WAVEFRONT))
    assert np.array_equal(a, b) and np.array_equal(ra, rb)
    for k in ("n_paths", "n_rays", "n_box_tests", "n_object_tests", "n_hits"):
        assert sa[k] == sb[k], k
    wt = pkg.World.create(pkg.RTW_SCENE_TEXTURED, image=earthmap)
    st = pkg.Scene(wt)
    camt = pkg.textured_camera(160, 3, 12).init()
    a, _, _ = st.render(camt, pkg.render_options(seed=6))
    b, _, _ = st.render(camt, pkg.render_options(seed=6, integrator=pkg.RTB_INTEGRATOR_WAVEFRONT))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("integrator", [0, 1])
def test_sample_ranges_and_partitions_compose_bit_exact(pkg, book1, integrator):
    world, scene = book1
    cam = pkg.book1_camera(176, 8, 50).init()   # 176 x 99: ragged against the 32 x 8 tiles
    full, _, _ = scene.render(cam, pkg.render_options(seed=9, integrator=integrator))
    # progressive: 0..3 then 3..8 into the same buffer (resumable accumulation, camera.zig:54-56)
    acc, _, _ = scene.render(cam, pkg.render_options(seed=9, sample_begin=0, sample_count=3, integrator=integrator))
    acc, _, _ = scene.render(cam, pkg.render_options(seed=9, sample_begin=3, sample_count=5, integrator=integrator),
                             accum=acc)
    assert np.array_equal(acc, full)
    # samples_per_launch batching does not change the result
    b, _, _ = scene.render(cam, pkg.render_options(seed=9, samples_per_launch=3, integrator=integrator))
    assert np.array_equal(b, full)
    # interleaved tile partition (multi-GPU mode A): 3 ranks write disjoint pixels; sum of buffers == full frame
    n = cam.image_width * cam.image_height
    parts = []
    for r in range(3):
        z = np.zeros((n, 4), np.float32)
        z, _, _ = scene.render(cam, pkg.render_options(seed=9, tile_rank=r, tile_world=3, integrator=integrator),
                               accum=z)
        parts.append(z)
    owners = sum((p[:, 3] > 0).astype(int) for p in parts)
    assert (owners == 1).all()
    assert np.array_equal(sum(parts), full)
    # Task{thread_idx, chunk_size} strips (camera.zig:94-95): 8 strips of size/8 pixels (+ remainder)
    acc2 = np.zeros((n, 4), np.float32)
    chunk = n // 8
    for t in range(8):
        cnt = chunk if t < 7 else n - 7 * chunk
        acc2, _, _ = scene.render(cam, pkg.render_options(seed=9, pixel_begin=t * chunk, pixel_count=cnt,
                                                          integrator=integrator), accum=acc2)
    assert np.array_equal(acc2, full)


def test_depth_limits_and_backgrounds(pkg, orc, book1):
    world, scene = book1
    for depth in (0, 1, 2):
        c = pkg.book1_camera(96, 2, depth)
        c.background_mode = pkg.RTB_BACKGROUND_SOLID
        c.background = (0.3, 0.6, 0.9)
        cam = c.init()
        g, cc = _render_pair(pkg, orc, world, scene, cam)
        _assert_same_paths(g, cc, 2)
        if depth == 0:
            assert (g[0][:, :3] == 0).all() and (g[0][:, 3] == 2).all()


def test_resolve_matches_oracle(pkg, orc):
    rng = np.random.default_rng(4)
    acc = np.zeros((4096, 4), np.float32)
    acc[:, :3] = rng.random((4096, 3)).astype(np.float32) * 40
    acc[:, 3] = rng.integers(1, 40, 4096)
    acc[0, :3] = 0
    acc[1, :3] = 1e9          # clamp to 0.999 -> 255
    acc[2, :3] = -1.0         # sqrt(<0) = NaN -> 0 (reference: undefined)
    acc[3, :3] = np.nan
    assert np.array_equal(pkg.resolve(acc), orc.resolve(acc))
    assert np.array_equal(pkg.resolve(acc, 7.0), orc.resolve(acc, 7.0))
    assert pkg.resolve(acc)[1].tolist() == [255, 255, 255, 255]


def test_camera_render_drop_in(pkg, book1):
    """rtw_camera_render = the call that replaces startRender's 8 threads (main.zig:314-326)."""
    world, scene = book1
    c = pkg.book1_camera(128, 3, 50)
    cam = c.init()
    buf, tex = pkg.new_writer(cam)
    buf[:] = 123.0  # scrub must reset it
    st = pkg.RtbRenderStats()
    o = c.options()
    ro = pkg.render_options(seed=11)
    rc = pkg._ffi.rtw().rtw_camera_render(scene._h, C.byref(o), C.byref(ro), 1, buf.ctypes.data, tex.ctypes.data,
                                          C.byref(st))
    assert rc == 0
    ref, ref_rgba, _ = scene.render(cam, pkg.render_options(seed=11))
    assert np.array_equal(buf, ref) and np.array_equal(tex, ref_rgba)
    assert (tex[:, 3] == 255).all() and (buf[:, 3] == 3).all()


def test_async_progress_and_cancel(pkg, book1):
    world, scene = book1
    cam = pkg.book1_camera(320, 64, 50).init()
    acc, rgba = pkg.new_writer(cam)
    job = scene.render_async(cam, pkg.render_options(seed=3, samples_per_launch=2), acc, rgba)
    seen = 0
    t0 = time.time()
    while time.time() - t0 < 60:
        done, total, running = job.progress()
        assert total == 64
        seen = max(seen, done)
        if done >= 6 or not running:
            break
        time.sleep(0.002)
    job.cancel()
    rc, st = job.wait()
    done, total, running = job.progress()
    assert not running
    if rc == pkg.RTB_ERR_CANCELLED:
        assert 0 < done < 64
    else:
        assert rc == 0 and done == 64
    # the buffers hold exactly `done` samples, like writer.buffer after STOP (main.zig:328-336)
    assert (acc[:, 3] == done).all()
    ref, _, _ = scene.render(cam, pkg.render_options(seed=3, sample_count=done))
    assert np.array_equal(acc, ref)
    job.destroy()


@pytest.mark.parametrize("integrator,traversal", [(0, 0), (1, 0), (1, 3)])
def test_statistical_parity_different_seeds(pkg, orc, book1, integrator, traversal):
    """North-star image bar: RMSE(GPU,CPU) <= 1.25 RMSE_self, |mean bias| <= max(0.002, 3 RMSE_self/sqrt(WH)) — for the
    megakernel and the wavefront integrator in reference order and for the benchmarked mode (wavefront + SAH16)."""
    world, scene = book1
    spp = 32
    cam = pkg.book1_camera(240, spp, 50).init()
    c1 = orc.render(world.desc, cam, pkg.render_options(seed=4321), want_rgba=False)[0]
    c2 = orc.render(world.desc, cam, pkg.render_options(seed=8765), want_rgba=False)[0]
    g = scene.render(cam, pkg.render_options(seed=1234, integrator=integrator, traversal=traversal))[0]
    m = lambda a: a[:, :3] / a[:, 3:4]
    rmse_self = np.sqrt(((m(c1) - m(c2)) ** 2).mean(axis=0))
    rmse = np.sqrt(((m(g) - m(c1)) ** 2).mean(axis=0))
    bias = np.abs((m(g) - m(c1)).mean(axis=0))
    assert (rmse <= 1.25 * rmse_self).all(), (rmse, rmse_self)
    assert (bias <= np.maximum(0.002, 3 * rmse_self / np.sqrt(g.shape[0]))).all(), (bias, rmse_self)


def test_error_paths(pkg, book1):
    world, scene = book1
    cam = pkg.book1_camera(64, 1, 4).init()
    with pytest.raises(pkg.RtbError) as e:
        scene.render(cam, pkg.render_options(tile_rank=2, tile_world=2))
    assert e.value.code == pkg.RTB_ERR_INVALID_ARGUMENT
    with pytest.raises(pkg.RtbError):
        scene.render(cam, pkg.render_options(pixel_begin=10, pixel_count=10**7))
    with pytest.raises(pkg.RtbError) as e:
        scene.render(cam, pkg.render_options(traversal=7))
    assert e.value.code == pkg.RTB_ERR_UNSUPPORTED
    with pytest.raises(pkg.RtbError):
        pkg.Scene(world, device=99)


def test_full_size_properties_config2(pkg, book1):
    """BASELINE config 2 frame size (1200x675), reduced spp: size-independent properties."""
    world, scene = book1
    cam = pkg.book1_camera(1200, 8, 50).init()
    assert cam.image_height == 675
    a, rgba, st = scene.render(cam, pkg.render_options(seed=1234, flags=pkg.RTB_FLAG_COUNT_WORK))
    assert st["n_paths"] == 1200 * 675 * 8
    assert np.isfinite(a).all() and (a[:, :3] >= 0).all() and (a[:, 3] == 8).all()
    # energy: no emitters and sky radiance <= 1 -> every sample <= 1
    assert (a[:, :3] <= 8.0 + 1e-3).all()
    # idempotence / determinism: the same call gives the same bits
    b, _, _ = scene.render(cam, pkg.render_options(seed=1234))
    assert np.array_equal(a, b)
    # linearity in samples: 8 spp == 5 spp + 3 spp
    c, _, _ = scene.render(cam, pkg.render_options(seed=1234, sample_count=5))
    c, _, _ = scene.render(cam, pkg.render_options(seed=1234, sample_begin=5, sample_count=3), accum=c)
    assert np.array_equal(a, c)
    assert 2.0 < st["n_rays"] / st["n_paths"] < 6.0


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_ordered_traversal_agrees(pkg, orc, book1, mode):
    """RTB_TRAVERSAL_ORDERED (1: near-child-first per octant on the host's tree) and RTB_TRAVERSAL_SAH (2: the same
    objects re-partitioned by the library): same slab/sphere arithmetic, different set/order of visited nodes.  The
    nearest hit may differ from the reference's only where float rounding puts a sphere root outside its own box;
    measure it on incoherent rays and require (near-)total agreement WITH THE ORACLE."""
    world, scene = book1
    cam = pkg.book1_camera(400, 10, 50).init()
    rng = np.random.default_rng(11)
    rays = np.concatenate([_rays_from_camera(orc, cam, 5, 30000, rng), _random_rays(pkg, rng, 30000, -12, 12)])
    ref = orc.trace_rays(world.desc, rays)
    ordr = scene.trace_rays(rays, traversal=mode)
    same = ref["object"] == ordr["object"]
    assert same.mean() >= 0.9999, f"{(~same).sum()} of {same.size} rays disagree"
    both = same & (ref["object"] >= 0)
    assert np.array_equal(ref["t"][both], ordr["t"][both])          # same arithmetic -> same bits
    assert np.array_equal(ref["front_face"][both], ordr["front_face"][both])
    bad = ~same & (ref["object"] >= 0) & (ordr["object"] >= 0)
    if bad.any():  # where they disagree the two candidate roots are within rounding of each other
        np.testing.assert_allclose(ref["t"][bad], ordr["t"][bad], rtol=1e-4)
    assert ordr["n_box_tests"].sum() <= ref["n_box_tests"].sum()
    if mode >= 2:  # the SAH partition must be much cheaper than the random-axis median-split tree
        assert ordr["n_box_tests"].sum() < 0.6 * ref["n_box_tests"].sum()


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("integrator", [0, 1])
def test_ordered_render_matches_reference_order_render(pkg, orc, book1, integrator, mode):
    """Same Philox streams, ORDERED / SAH vs the ORACLE's reference-order render: identical paths except where a hit
    index differs (then the path diverges; must stay a vanishing fraction of the pixels)."""
    world, scene = book1
    cam = pkg.book1_camera(200, 4, 50).init()
    o = pkg.render_options(seed=21, integrator=integrator, flags=pkg.RTB_FLAG_COUNT_WORK, traversal=mode)
    b, _, sb = scene.render(cam, o)
    a, _, sa = orc.render(world.desc, cam, o, n_threads=8)
    diff = np.abs(a[:, :3] - b[:, :3]).max(axis=1)
    assert np.count_nonzero(diff > 1e-4) <= 2e-3 * diff.shape[0]
    assert abs(sa["n_rays"] - sb["n_rays"]) <= 1e-3 * sa["n_rays"] and sa["n_paths"] == sb["n_paths"]
    assert sb["n_box_tests"] <= sa["n_box_tests"]


def test_sah16_is_a_conservative_superset_of_sah(pkg, orc, book1, earthmap):
    """RTB_TRAVERSAL_SAH16 walks the SAH tree with 16-byte packed box nodes and a half-precision slab test whose
    constants carry an error margin: it may enter boxes the f32 test culls, never the reverse, and leaves keep the
    reference's arithmetic.  So per ray: the same nearest hit as SAH (same bits in t), at least as many leaves tested,
    a few per cent more box nodes visited; and the three kernels that walk the packed layout (ray queries, megakernel,
    wavefront from shared memory) produce the same image bit for bit."""
    world, scene = book1
    cam = pkg.book1_camera(400, 10, 50).init()
    rng = np.random.default_rng(23)
    prim = _rays_from_camera(orc, cam, 5, 40000, rng)
    sec = _secondary_rays(pkg, orc, world.desc, prim[:6000], orc.trace_rays(world.desc, prim[:6000]), 3)
    stress = _random_rays(pkg, rng, 20000, -12, 12)
    stress["direction"][:3000, 0] = 0.0          # axis-parallel: the axis is dropped from the packed test
    stress["direction"][3000:6000, 1] *= 1e-7    # |1/d| beyond binary16
    stress["origin"][6000:9000] *= 100.0         # far origins: the margin grows with |o|
    rays = np.concatenate([prim, sec, stress])
    s32 = scene.trace_rays(rays, traversal=pkg.RTB_TRAVERSAL_SAH)
    s16 = scene.trace_rays(rays, traversal=pkg.RTB_TRAVERSAL_SAH16)
    ref = orc.trace_rays(world.desc, rays)
    assert (s16["n_object_tests"] >= s32["n_object_tests"]).all()
    same = s16["object"] == s32["object"]
    assert same.mean() >= 0.99995, int((~same).sum())
    # where they differ SAH16 (the superset) must be the one that agrees with the reference, or all three are within
    # rounding of each other
    for k in np.nonzero(~same)[0]:
        assert s16["object"][k] == ref["object"][k] or np.isclose(s16["t"][k], s32["t"][k], rtol=1e-4), int(k)
    hit = same & (s32["object"] >= 0)
    for f in ("t", "p", "normal", "front_face"):
        assert np.array_equal(s16[f][hit], s32[f][hit]), f
    assert (s16["object"] == ref["object"]).mean() >= 0.9999
    n_ord = len(prim) + len(sec)
    ratio = s16["n_box_tests"][:n_ord].sum() / s32["n_box_tests"][:n_ord].sum()
    assert 1.0 <= ratio <= 1.08, ratio
    cam2 = pkg.book1_camera(256, 6, 50).init()
    a, _, sa = scene.render(cam2, pkg.render_options(seed=4, integrator=1, traversal=3, flags=pkg.RTB_FLAG_COUNT_WORK))
    b, _, sb = scene.render(cam2, pkg.render_options(seed=4, integrator=0, traversal=3, flags=pkg.RTB_FLAG_COUNT_WORK))
    assert np.array_equal(a, b)
    for k in ("n_rays", "n_hits"):
        assert sa[k] == sb[k], k
    # wf_extend_stream parks leaves and keeps walking with the old t_max: a superset of the megakernel's visits
    assert sb["n_box_tests"] <= sa["n_box_tests"] <= 1.06 * sb["n_box_tests"]
    assert sb["n_object_tests"] <= sa["n_object_tests"] <= 1.06 * sb["n_object_tests"]
    c, _, sc = scene.render(cam2, pkg.render_options(seed=4, integrator=1, traversal=2, flags=pkg.RTB_FLAG_COUNT_WORK))
    assert np.count_nonzero(np.abs(a[:, :3] - c[:, :3]).max(axis=1) > 1e-4) <= 1e-3 * a.shape[0]
    assert sc["n_box_tests"] <= sa["n_box_tests"] <= 1.08 * sc["n_box_tests"]
    # a textured world (checker + image + noise): same check on the hit queries
    tw = pkg.World.create(pkg.RTW_SCENE_TEXTURED, image=earthmap)
    ts = pkg.Scene(tw)
    trays = _random_rays(pkg, rng, 20000, -10, 10)
    t32, t16 = ts.trace_rays(trays, traversal=2), ts.trace_rays(trays, traversal=3)
    assert np.array_equal(t32["object"], t16["object"]) and np.array_equal(t32["t"], t16["t"])


def test_wavefront_multi_batch_pipeline_bit_exact(pkg, book1):
    """Enough samples for several ~8 M-path batches on several streams: per-pixel sums must still be in sample
    order, i.e. bit-identical to the megakernel."""
    world, scene = book1
    cam = pkg.book1_camera(1200, 56, 50).init()   # 810 000 px x 56 spp = 45 M paths -> 6 batches over 4 lanes (reuse)
    a, _, _ = scene.render(cam, pkg.render_options(seed=31, integrator=pkg.RTB_INTEGRATOR_WAVEFRONT))
    b, _, _ = scene.render(cam, pkg.render_options(seed=31, integrator=pkg.RTB_INTEGRATOR_MEGAKERNEL))
    assert np.array_equal(a, b)


def test_large_scene_global_memory_path(pkg, orc):
    """BASELINE config 4's family at a size the oracle finishes in seconds: 200 000 random spheres = 399 999 host
    nodes (12.8 MB; the SAH layout 19 MB) — far beyond shared memory, so every kernel walks the layouts in global
    memory / L2 instead of the staged copy the Book-1 scene uses."""
    world = pkg.World.create(pkg.RTW_SCENE_RANDOM_SPHERES, n_spheres=200000)
    scene = pkg.Scene(world)
    cam = pkg.million_camera(256, 2, 8).init()
    rng = np.random.default_rng(17)
    inside = _random_rays(pkg, rng, 6000, -450, 450)
    inside["origin"][:, 1] = rng.uniform(0.1, 40, 6000).astype(np.float32)       # the slab the spheres live in
    rays = np.concatenate([_rays_from_camera(orc, cam, 3, 6000, rng), inside])
    cpu = orc.trace_rays(world.desc, rays)
    assert (cpu["object"] >= 0).mean() > 0.4 and len(set(cpu["object"].tolist())) > 1000
    assert cpu["n_box_tests"].mean() > 300          # the random-axis median-split tree is expensive at this size
    _assert_hits_equal(scene.trace_rays(rays), cpu)                       # reference order: bit-exact incl. visit counts
    d = world.desc.contents

    def undecidable_in_f32(k, obj):
        """True if the discriminant of Sphere.hit (src/objects.zig:122-127) for ray k and sphere obj is, in float64,
        smaller than the rounding error its two f32 terms carry — the hit/miss decision is then rounding noise."""
        if obj < 0:
            return False
        h = d.hittables[int(obj)]
        o, dd = rays["origin"][k].astype(np.float64), rays["direction"][k].astype(np.float64)
        oc = o - np.array(h.a[:], np.float64)
        hb2, ac = (oc @ dd) ** 2, (dd @ dd) * (oc @ oc - float(h.radius) ** 2)
        return abs(hb2 - ac) <= 16.0 * 2.0 ** -24 * (hb2 + abs(ac))

    for mode in (pkg.RTB_TRAVERSAL_ORDERED, pkg.RTB_TRAVERSAL_SAH):
        got = scene.trace_rays(rays, traversal=mode)
        same = got["object"] == cpu["object"]
        both = same & (cpu["object"] >= 0)
        assert np.array_equal(got["t"][both], cpu["t"][both])
        if mode == pkg.RTB_TRAVERSAL_ORDERED:      # same leaves tested as the reference (no leaf boxes): same hits
            assert same.mean() >= 0.9999, int((~same).sum())
            continue
        # SAH box-tests its leaves (with its own padded boxes).  From hundreds of units away the f32 discriminant
        # b^2 - a*c of a 0.1-0.5 radius sphere is dominated by cancellation error, so the reference — which tests
        # leaves WITHOUT a box test, after exact tests on unpadded parent boxes — reports hits on spheres the ray
        # passes a few tenths of a unit beside (and misses a few it should see); a box test decides those differently.
        # Pin exactly that: wherever SAH disagrees, one of the two spheres involved is undecidable in f32.
        assert same.mean() >= 0.995, int((~same).sum())
        for k in np.nonzero(~same)[0]:
            assert undecidable_in_f32(k, cpu["object"][k]) or undecidable_in_f32(k, got["object"][k]), int(k)
    sah = scene.trace_rays(rays, traversal=pkg.RTB_TRAVERSAL_SAH)
    assert sah["n_box_tests"].sum() < 0.5 * cpu["n_box_tests"].sum()
    # renders: same streams as the oracle, both integrators, reference order; SAH against reference order on the GPU
    o = pkg.render_options(seed=5, flags=pkg.RTB_FLAG_COUNT_WORK)
    c_acc, _, c_st = orc.render(world.desc, cam, o, n_threads=8)
    for integrator in (pkg.RTB_INTEGRATOR_MEGAKERNEL, pkg.RTB_INTEGRATOR_WAVEFRONT):
        o = pkg.render_options(seed=5, integrator=integrator, flags=pkg.RTB_FLAG_COUNT_WORK)
        g_acc, _, g_st = scene.render(cam, o)
        assert g_st["n_paths"] == c_st["n_paths"] and abs(g_st["n_rays"] - c_st["n_rays"]) <= 1e-3 * c_st["n_rays"]
        assert abs(g_st["n_box_tests"] - c_st["n_box_tests"]) <= 2e-3 * c_st["n_box_tests"]
        diff = np.abs(g_acc[:, :3] - c_acc[:, :3]).max(axis=1)
        assert np.count_nonzero(diff > 1e-4) <= 5e-3 * diff.shape[0]
        o2 = pkg.render_options(seed=5, integrator=integrator, traversal=pkg.RTB_TRAVERSAL_SAH)
        s_acc, _, _ = scene.render(cam, o2)
        diff = np.abs(s_acc[:, :3] - g_acc[:, :3]).max(axis=1)
        # ~0.25 % of the rays are undecidable in f32 from this camera (see above); a pixel holds ~4 of them
        assert np.count_nonzero(diff > 1e-4) <= 2e-2 * diff.shape[0]


def test_wavefront_multi_pass_frames_and_tile_stats(pkg, book1, tmp_path):
    """Frames with more pixels than one wavefront pass may hold are rendered as several passes over interleaved tile
    subsets (RTB_WF_MAX_SLOTS forces that on a small frame; it is read once per process, hence the subprocess):
    bit-identical to the megakernel.  Also: n_paths is exact under the tile partition."""
    import subprocess
    import sys
    script = tmp_path / "multipass.py"
    script.write_text(f"""
import importlib, sys
import numpy as np
sys.path.insert(0, {ROOT!r})
pkg = importlib.import_module("zig-raytracing-weekend_b200")
world = pkg.World.book1()
scene = pkg.Scene(world)
cam = pkg.book1_camera(200, 6, 50).init()
a, _, sa = scene.render(cam, pkg.render_options(seed=9, integrator=pkg.RTB_INTEGRATOR_WAVEFRONT))
b, _, sb = scene.render(cam, pkg.render_options(seed=9, integrator=pkg.RTB_INTEGRATOR_MEGAKERNEL))
assert np.array_equal(a, b), "multi-pass wavefront differs from the megakernel"
assert sa["n_launches"] > 5 * (2 * 50 + 2), sa["n_launches"]          # really several passes
print("ok", sa["n_launches"])
""")
    env = dict(os.environ, RTB_WF_MAX_SLOTS="4096")
    out = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr
    world, scene = book1
    cam = pkg.book1_camera(203, 2, 8).init()      # 203 x 114: ragged right and bottom tiles
    total = 0
    for rank in range(3):
        o = pkg.render_options(seed=2, integrator=pkg.RTB_INTEGRATOR_MEGAKERNEL, tile_rank=rank, tile_world=3)
        acc, _, st = scene.render(cam, o)
        assert st["n_paths"] == 2 * np.count_nonzero(acc[:, 3] == 2.0)
        total += st["n_paths"]
    assert total == 2 * cam.image_width * cam.image_height


@pytest.mark.parametrize("forced", ["0", "1", "3", "12"])
def test_wavefront_tail_switch_does_not_change_the_result(pkg, tmp_path, forced):
    """wf_tail finishes the paths still alive after bounce k in one launch (the megakernel's loop, entered mid-path).
    Where the switch happens is a performance decision — learned from the first batch, or forced here through
    RTB_WF_TAIL_BOUNCE (0 = never, 1 = the whole path after the camera ray, ...) — and must not change a single bit."""
    import subprocess
    import sys
    script = tmp_path / "tail.py"
    script.write_text(f"""
import importlib, sys
import numpy as np
sys.path.insert(0, {ROOT!r})
pkg = importlib.import_module("zig-raytracing-weekend_b200")
for world, camo in ((pkg.World.book1(), pkg.book1_camera(320, 5, 50)),
                    (pkg.World.create(pkg.RTW_SCENE_CORNELL_SMOKE), pkg.cornell_camera(96, 4, 30))):
    scene = pkg.Scene(world)
    cam = camo.init()
    for trav in (0, 2, 3):
        a, _, sa = scene.render(cam, pkg.render_options(seed=9, integrator=1, traversal=trav, flags=pkg.RTB_FLAG_COUNT_WORK))
        b, _, sb = scene.render(cam, pkg.render_options(seed=9, integrator=0, traversal=trav, flags=pkg.RTB_FLAG_COUNT_WORK))
        assert np.array_equal(a, b), "wavefront with tail switch differs from the megakernel"
        for k in ("n_paths", "n_rays", "n_box_tests", "n_object_tests", "n_hits"):
            if trav == 3 and k in ("n_box_tests", "n_object_tests"):     # wf_extend_stream visits a superset
                assert sb[k] <= sa[k] <= 1.06 * sb[k], (k, sa[k], sb[k])
            else:
                assert sa[k] == sb[k], (k, sa[k], sb[k])
print("ok")
""")
    env = dict(os.environ, RTB_WF_TAIL_BOUNCE=forced)
    out = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr


@pytest.mark.parametrize("mode", ["1", "2"])
def test_streaming_extend_kernel_gives_the_same_image(pkg, tmp_path, mode):
    """wf_extend_stream (RTB_EXTEND_STREAM=1: bounces >= 1, 2: every bounce; off by default because it measured slower,
    DESIGN.md section 5) parks leaves and refills finished lanes.  Leaves are still tested in walk order and accepted
    with the current t_max, so the image is the same bits as the plain kernel's and the megakernel's; only the visit
    counters grow (parked leaves keep the old t_max while the lane walks on)."""
    import subprocess
    import sys
    script = tmp_path / "stream.py"
    script.write_text(f"""
import importlib, sys
import numpy as np
sys.path.insert(0, {ROOT!r})
pkg = importlib.import_module("zig-raytracing-weekend_b200")
scene = pkg.Scene(pkg.World.book1())
cam = pkg.book1_camera(400, 6, 50).init()
a, _, sa = scene.render(cam, pkg.render_options(seed=9, integrator=1, traversal=3, flags=pkg.RTB_FLAG_COUNT_WORK))
b, _, sb = scene.render(cam, pkg.render_options(seed=9, integrator=0, traversal=3, flags=pkg.RTB_FLAG_COUNT_WORK))
assert np.array_equal(a, b), np.count_nonzero((a != b).any(axis=1))
assert sa["n_rays"] == sb["n_rays"] and sa["n_hits"] == sb["n_hits"]
assert sb["n_box_tests"] < sa["n_box_tests"] <= 1.25 * sb["n_box_tests"], (sa["n_box_tests"], sb["n_box_tests"])
print("ok")
""")
    env = dict(os.environ, RTB_EXTEND_STREAM=mode)
    out = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr


def test_straggler_eviction_kernel_traces_the_same_walks(pkg, tmp_path):
    """wf_extend_evict (RTB_EXTEND_EVICT=1; off by default because it measured slower) moves the last
    few walking lanes of a warp to a shared-memory buffer and finishes them later as a dense group.  A straggler resumes
    at the node it stopped at with the nearest hit it had, so the image AND every work counter equal the megakernel's."""
    import subprocess
    import sys
    script = tmp_path / "evict.py"
    script.write_text(f"""
import importlib, sys
import numpy as np
sys.path.insert(0, {ROOT!r})
pkg = importlib.import_module("zig-raytracing-weekend_b200")
scene = pkg.Scene(pkg.World.book1())
cam = pkg.book1_camera(400, 6, 50).init()
a, _, sa = scene.render(cam, pkg.render_options(seed=9, integrator=1, traversal=3, flags=pkg.RTB_FLAG_COUNT_WORK))
b, _, sb = scene.render(cam, pkg.render_options(seed=9, integrator=0, traversal=3, flags=pkg.RTB_FLAG_COUNT_WORK))
assert np.array_equal(a, b), np.count_nonzero((a != b).any(axis=1))
for k in ("n_rays", "n_hits", "n_box_tests", "n_object_tests"):
    assert sa[k] == sb[k], (k, sa[k], sb[k])
print("ok")
""")
    env = dict(os.environ, RTB_EXTEND_EVICT="1")
    out = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
