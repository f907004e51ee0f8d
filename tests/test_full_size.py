"""BASELINE.json configs 3, 4 and 5 at their FULL frame / scene sizes, through size-independent properties (the oracle
cannot finish these in seconds): determinism, megakernel == wavefront bit for bit, linearity in samples, exact
composition of sample and tile partitions, sample counters, energy bounds.  Config 2 is in test_gpu_parity.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MEGA, WAVE = 0, 1


@pytest.fixture(scope="module")
def book1(pkg):
    world = pkg.World.book1()
    return world, pkg.Scene(world)


def test_config3_textured_full_size(pkg, earthmap):
    """800 x 450, 256 spp, checker + earthmap + perlin spheres (all four texture kinds, every material class)."""
    world = pkg.World.create(pkg.RTW_SCENE_TEXTURED, image=earthmap)
    scene = pkg.Scene(world)
    cam = pkg.textured_camera(800, 256, 50).init()
    assert (cam.image_width, cam.image_height) == (800, 450)
    w, _, st = scene.render(cam, pkg.render_options(seed=7, integrator=WAVE, traversal=pkg.RTB_TRAVERSAL_SAH,
                                                    flags=pkg.RTB_FLAG_COUNT_WORK))
    assert st["n_paths"] == 800 * 450 * 256 and (w[:, 3] == 256).all() and np.isfinite(w).all()
    assert (w[:, :3] >= 0).all() and (w[:, :3] <= 256.0 * 1.001).all()      # no emitters, background <= 1
    m, _, _ = scene.render(cam, pkg.render_options(seed=7, integrator=MEGA, traversal=pkg.RTB_TRAVERSAL_SAH))
    assert np.array_equal(w, m)                                              # 92 M paths, 12 batches over 4 lanes
    # linearity in samples: 256 = 100 + 156, continued in the same buffer
    c, _, _ = scene.render(cam, pkg.render_options(seed=7, integrator=WAVE, traversal=pkg.RTB_TRAVERSAL_SAH,
                                                   sample_count=100))
    c, _, _ = scene.render(cam, pkg.render_options(seed=7, integrator=WAVE, traversal=pkg.RTB_TRAVERSAL_SAH,
                                                   sample_begin=100, sample_count=156), accum=c)
    assert np.array_equal(w, c)
    # the reference-order render of the same streams: same image up to the few f32-undecidable hits
    r, _, _ = scene.render(cam, pkg.render_options(seed=7, integrator=WAVE, traversal=pkg.RTB_TRAVERSAL_REFERENCE))
    diff = np.abs(r[:, :3] - w[:, :3]).max(axis=1) / 256.0
    assert np.count_nonzero(diff > 1e-4) <= 0.02 * diff.shape[0] and abs(float((r[:, :3] - w[:, :3]).mean())) / 256 < 1e-4


def test_config4_million_spheres_full_scene(pkg):
    """1 000 000 random spheres (1 999 999 host nodes, 64 MB; the SAH layout 96 MB per octant): every kernel walks the
    tree in global memory / L2.  3840 x 2160 at 1 spp for the SAH layouts, a quarter-size frame for reference order."""
    world = pkg.World.create(pkg.RTW_SCENE_RANDOM_SPHERES, n_spheres=1000000)
    assert world.desc.contents.n_hittables == 1000001 and world.desc.contents.n_nodes == 2000001
    scene = pkg.Scene(world)
    cam = pkg.million_camera(3840, 1, 50).init()
    assert (cam.image_width, cam.image_height) == (3840, 2160)
    o = dict(seed=11, traversal=pkg.RTB_TRAVERSAL_SAH)
    w, _, st = scene.render(cam, pkg.render_options(integrator=WAVE, flags=pkg.RTB_FLAG_COUNT_WORK, **o))
    m, _, _ = scene.render(cam, pkg.render_options(integrator=MEGA, **o))
    assert np.array_equal(w, m) and (w[:, 3] == 1).all() and np.isfinite(w).all()
    assert st["n_paths"] == 3840 * 2160 and st["n_box_tests"] / st["n_rays"] < 120     # SAH: ~55 slab tests per ray
    small = pkg.million_camera(960, 1, 50).init()
    r1, _, sr = scene.render(small, pkg.render_options(seed=11, integrator=WAVE, flags=pkg.RTB_FLAG_COUNT_WORK))
    r2, _, _ = scene.render(small, pkg.render_options(seed=11, integrator=MEGA))
    assert np.array_equal(r1, r2)
    assert sr["n_box_tests"] / sr["n_rays"] > 300           # the host's random-axis median-split tree: ~770 per ray
    s1, _, _ = scene.render(small, pkg.render_options(seed=11, integrator=WAVE, traversal=pkg.RTB_TRAVERSAL_SAH))
    diff = np.abs(s1[:, :3] - r1[:, :3]).max(axis=1)
    assert np.count_nonzero(diff > 1e-4) <= 0.03 * diff.shape[0]       # f32-undecidable hits, see DESIGN §2


def test_config5_8k_frame_partitions_compose(pkg, book1):
    """7680 x 4320 (33 177 600 pixels, 531 MB of accumulators), 2 samples: the sample partition and the tile partition
    of a two-rank run both reproduce the single-device frame exactly, pixel for pixel."""
    world, scene = book1
    cam = pkg.book1_camera(7680, 2, 50).init()
    assert (cam.image_width, cam.image_height) == (7680, 4320)
    base = dict(seed=3, integrator=WAVE, traversal=pkg.RTB_TRAVERSAL_SAH)
    full, rgba, st = scene.render(cam, pkg.render_options(**base))
    assert st["n_paths"] == 2 * 7680 * 4320 and (full[:, 3] == 2).all() and (rgba[:, 3] == 255).all()
    # "samples": rank r renders global sample r into its own zeroed buffer; the exchange step adds the buffers
    s0, _, _ = scene.render(cam, pkg.render_options(sample_begin=0, sample_count=1, **base))
    s1, _, _ = scene.render(cam, pkg.render_options(sample_begin=1, sample_count=1, **base))
    assert np.array_equal(s0[:, :3] + s1[:, :3], full[:, :3])
    del s0, s1
    # "tiles": rank r renders the 32x8 tiles t with t % 2 == r; disjoint pixels, zero elsewhere
    t0, _, a0 = scene.render(cam, pkg.render_options(tile_rank=0, tile_world=2, **base))
    t1, _, a1 = scene.render(cam, pkg.render_options(tile_rank=1, tile_world=2, **base))
    assert a0["n_paths"] + a1["n_paths"] == st["n_paths"]
    own0, own1 = t0[:, 3] == 2, t1[:, 3] == 2
    assert not (own0 & own1).any() and (own0 | own1).all()
    assert np.array_equal(np.where(own0[:, None], t0[:, :3], t1[:, :3]), full[:, :3])
    assert (t0[~own0, :3] == 0).all() and (t1[~own1, :3] == 0).all()
