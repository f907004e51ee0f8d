"""Multi-GPU behind ONE call of the C ABI (include/rtb.h: rtb_group_create / rtb_group_render) — the counterpart of
the reference's startRender, which starts all strips of one frame with one call and lets them share one buffer
(src/main.zig:314-326).  One process, one scene replica + one host thread per device, peer-memory exchange.

The tests run on the single-GPU box too: an ordinal may be listed twice (two replicas on one device), which exercises
the same partitioning, threading and exchange code; with >= 2 devices visible they also run across real devices."""
import ctypes as C

import numpy as np
import pytest


def _device_sets(pkg):
    n = pkg.device_count()
    sets = [[0], [0, 0], [0, 0, 0]]
    if n >= 2:
        sets += [[0, 1]]
    if n >= 4:
        sets += [list(range(4))]
    if n >= 8:
        sets += [list(range(8))]
    return sets


@pytest.mark.gpu
def test_group_render_reproduces_the_single_device_frame(pkg, orc):
    world = pkg.World.book1()
    scene = pkg.Scene(world)
    cam = pkg.book1_camera(320, 9, 50).init()
    for integrator, trav in ((1, 3), (0, 0)):
        o = pkg.render_options(seed=17, integrator=integrator, traversal=trav, flags=pkg.RTB_FLAG_COUNT_WORK)
        one, one_rgba, st1 = scene.render(cam, o)
        for devices in _device_sets(pkg):
            grp = pkg.SceneGroup(world, devices)
            # tiles: disjoint pixels -> the very same bits as one device, sums and RGBA8
            t, t_rgba, stt = grp.render(cam, o, partition=pkg.RTB_PARTITION_TILES)
            assert np.array_equal(t, one) and np.array_equal(t_rgba, one_rgba), (devices, integrator)
            for k in ("n_paths", "n_rays", "n_hits"):
                assert stt[k] == st1[k], (k, devices)
            for k in ("n_box_tests", "n_object_tests"):    # SAH16's streaming extend kernel parks leaves: how many
                # extra nodes a ray visits depends on the rays it shares a warp with — the hits do not
                assert abs(stt[k] - st1[k]) <= (2e-3 * st1[k] if trav == 3 else 0), (k, devices)
            # samples: the same paths; only the association of the per-pixel float sums differs
            s, s_rgba, sts = grp.render(cam, o, partition=pkg.RTB_PARTITION_SAMPLES)
            assert (s[:, 3] == 9).all() and sts["n_paths"] == st1["n_paths"] and sts["n_rays"] == st1["n_rays"]
            np.testing.assert_allclose(s[:, :3], one[:, :3], rtol=3e-6, atol=2e-6)
            assert np.abs(s_rgba.astype(int) - one_rgba.astype(int)).max() <= 1
            if len(devices) == 1:
                assert np.array_equal(s, one)
            grp.close()
    # against the oracle too (same streams), through the group call
    grp = pkg.SceneGroup(world, _device_sets(pkg)[-1])
    o = pkg.render_options(seed=1234, integrator=1, traversal=0)
    g, _, _ = grp.render(cam, o, partition=pkg.RTB_PARTITION_TILES)
    c, _, _ = orc.render(world.desc, cam, o, n_threads=8)
    assert np.count_nonzero(np.abs(g[:, :3] - c[:, :3]).max(axis=1) > 1e-4) <= 2e-3 * g.shape[0]
    grp.close()


@pytest.mark.gpu
def test_group_render_adds_to_the_callers_buffer_and_checks_arguments(pkg):
    world = pkg.World.book1()
    cam = pkg.book1_camera(160, 6, 20).init()
    grp = pkg.SceneGroup(world, [0, 0])
    o = pkg.render_options(seed=3, integrator=1, traversal=2)
    full, _, _ = grp.render(cam, o, partition=pkg.RTB_PARTITION_TILES)
    # progressive use: 2 samples, then 4 more on top of what the buffer holds (like rtb_render)
    a, _, _ = grp.render(cam, pkg.render_options(seed=3, integrator=1, traversal=2, sample_count=2),
                         partition=pkg.RTB_PARTITION_TILES)
    a, rgba, _ = grp.render(cam, pkg.render_options(seed=3, integrator=1, traversal=2, sample_begin=2, sample_count=4),
                            partition=pkg.RTB_PARTITION_TILES, accum=a)
    # the earlier sums live on the root device, so a tile owned by another device adds (s0+s1) + (s2+..+s5) instead of
    # the sequential sum: equal up to float association (exact for the root's own tiles and for a cleared buffer)
    np.testing.assert_allclose(a[:, :3], full[:, :3], rtol=2e-6, atol=1e-6)
    assert (a[:, 3] == 6).all() and (rgba[:, 3] == 255).all()
    assert np.array_equal(a, full) or np.count_nonzero((a != full).any(axis=1)) < 0.6 * a.shape[0]
    n = C.c_uint32(0)
    assert pkg._ffi.rtb().rtb_group_size(grp._h, C.byref(n)) == 0 and n.value == 2
    with pytest.raises(pkg.RtbError) as e:
        grp.render(cam, pkg.render_options(tile_rank=1, tile_world=2))
    assert e.value.code == pkg.RTB_ERR_INVALID_ARGUMENT
    with pytest.raises(pkg.RtbError):
        grp.render(cam, o, partition=7)
    with pytest.raises(pkg.RtbError):
        pkg.SceneGroup(world, [0, 99])
    with pytest.raises(pkg.RtbError):
        pkg.SceneGroup(world, [])
    grp.close()
