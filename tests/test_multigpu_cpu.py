"""CPU tests of the multi-GPU plumbing (one process per GPU, torch.distributed): partition planning and the single
exchange step, exercised with world_size = 2 over gloo.  The per-rank "render" is the oracle here (allowed in
tests/): what is under test is multigpu.plan / apply / combine, i.e. that the union of the ranks' work is exactly the
single-device frame and that the reduce puts the sum and the true samples-per-pixel on rank 0."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mg():
    return importlib.import_module("zig-raytracing-weekend_b200.multigpu")


def test_plan_sample_partition_covers_every_sample_once():
    mg = _mg()
    for world in (1, 2, 3, 4, 8):
        for spp in (1, 7, 8, 500, 1024):
            parts = [mg.plan("samples", r, world, spp, sample_base=5) for r in range(world)]
            assert sum(p.sample_count for p in parts) == spp
            covered = sorted(s for p in parts for s in range(p.sample_begin, p.sample_begin + p.sample_count))
            assert covered == list(range(5, 5 + spp))
            assert max(p.sample_count for p in parts) - min(p.sample_count for p in parts) <= 1
            assert all(p.total_samples == spp and p.tile_world == 1 for p in parts)


def test_plan_weak_and_tiles():
    mg = _mg()
    parts = [mg.plan("samples", r, 4, 500, weak=True) for r in range(4)]
    assert [p.sample_begin for p in parts] == [0, 500, 1000, 1500] and all(p.sample_count == 500 for p in parts)
    assert all(p.total_samples == 2000 for p in parts)
    parts = [mg.plan("tiles", r, 4, 64) for r in range(4)]
    assert [(p.tile_rank, p.tile_world, p.sample_begin, p.sample_count) for p in parts] == [(r, 4, 0, 64) for r in range(4)]
    with pytest.raises(ValueError):
        mg.plan("tiles", 0, 2, 8, weak=True)
    with pytest.raises(ValueError):
        mg.plan("samples", 2, 2, 8)
    with pytest.raises(ValueError):
        mg.plan("rows", 0, 2, 8)


def test_plan_rate_proportional_split():
    """bench.py's N > 1 sample split follows the rate every rank showed in warm-up (a slower GPU gets fewer samples):
    every sample exactly once, shares proportional to the weights, at least one sample per rank, identical on all
    ranks (each rank computes the whole table from the all-gathered weights)."""
    mg = _mg()
    for world, weights in ((2, [1.0, 1.0]), (4, [1.0, 0.94, 1.0, 1.02]), (8, [1 / 124.1] + [1 / 131.8] * 7),
                           (3, [1e-3, 1.0, 1.0])):
        for spp, weak in ((500, True), (64, False), (1024, False), (8, False)):
            parts = [mg.plan("samples", r, world, spp, sample_base=3, weak=weak, weights=weights) for r in range(world)]
            total = spp * world if weak else spp
            assert sum(p.sample_count for p in parts) == total and all(p.total_samples == total for p in parts)
            covered = [s for p in parts for s in range(p.sample_begin, p.sample_begin + p.sample_count)]
            assert covered == list(range(3, 3 + total))                 # contiguous, in rank order, no gap, no overlap
            assert min(p.sample_count for p in parts) >= 1
            if total >= 50 * world:
                for p, w in zip(parts, weights):
                    assert abs(p.sample_count - total * w / sum(weights)) <= 1.0 + 1e-9 or w == 1e-3
    assert mg.split_samples(10, 4) == [3, 3, 2, 2]
    with pytest.raises(ValueError):
        mg.split_samples(3, 4)
    with pytest.raises(ValueError):
        mg.split_samples(30, 3, [1.0, 0.0, 1.0])


def _tile_mask(width, height, rank, world):
    """Pixels of the 32x8 tiles with tile_index % world == rank (include/rtb.h: RtbRenderOptions.tile_rank)."""
    ys, xs = np.mgrid[0:height, 0:width]
    tiles_x = (width + 31) // 32
    tile = (ys // 8) * tiles_x + xs // 32
    return (tile % world == rank).reshape(-1)


def _worker(rank, world, mode, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("zig-raytracing-weekend_b200")
    mg = _mg()
    import oracle_ffi as orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wld = pkg.World.book1()
        cam = pkg.book1_camera(96, 6, 8).init()
        n = cam.image_width * cam.image_height
        part = mg.plan(mode, rank, world, 6)
        opt = mg.apply(part, pkg.render_options(seed=1234))
        acc = np.zeros((n, 4), np.float32)
        if mode == "samples":
            acc, _, _ = orc.render(wld.desc, cam, opt, n_threads=2, accum=acc, want_rgba=False)
        else:  # the oracle has no tile option: render everything, keep this rank's tiles
            full, _, _ = orc.render(wld.desc, cam, pkg.render_options(seed=1234), n_threads=2, accum=acc, want_rgba=False)
            mask = _tile_mask(cam.image_width, cam.image_height, part.tile_rank, part.tile_world)
            acc = np.where(mask[:, None], full, 0).astype(np.float32)
        t = torch.from_numpy(acc.copy())
        mg.combine(t, part)
        if rank == 0:
            np.save(out, t.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["samples", "tiles"])
def test_two_rank_reduce_reproduces_the_single_device_frame(pkg, orc, mode, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "combined.npy")
    port = 29500 + (os.getpid() % 2000) + (0 if mode == "samples" else 1)
    mp.spawn(_worker, args=(2, mode, port, out), nprocs=2, join=True)
    combined = np.load(out)
    wld = pkg.World.book1()
    cam = pkg.book1_camera(96, 6, 8).init()
    ref, _, _ = orc.render(wld.desc, cam, pkg.render_options(seed=1234), n_threads=4, want_rgba=False)
    assert (combined[:, 3] == 6).all()                       # .w = the frame's true samples per pixel
    if mode == "tiles":
        assert np.array_equal(combined[:, :3], ref[:, :3])   # disjoint pixels: the sum is exact
    else:
        # sample ranges: same paths, but (s0+s1+s2) + (s3+s4+s5) instead of the sequential sum
        np.testing.assert_allclose(combined[:, :3], ref[:, :3], rtol=2e-6, atol=1e-6)
    assert np.abs(orc.resolve(combined).astype(int) - orc.resolve(ref).astype(int)).max() <= 1
