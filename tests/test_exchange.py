"""The multi-GPU exchange step over peer memory (include/rtb.h: rtb_exchange_resolve).

  * CPU: the slice arithmetic (every pixel combined by exactly one rank, 256-pixel aligned).
  * GPU, one device: the fused kernel against numpy sums in rank order + the oracle's resolve, with the "ranks"
    being plain buffers of one process.
  * GPU, one device, TWO processes: the real plumbing of multigpu.PeerExchange — library-owned buffers exported over
    CUDA IPC, handles carried by torch.distributed (gloo here, because NCCL refuses two ranks on one GPU), barrier,
    one kernel per rank writing straight into rank 0's buffers.
  * GPU, >= 2 devices (skipped on the single-GPU box): the same over NCCL, one rank per GPU, against the NCCL reduce.
"""
import ctypes as C
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_exchange_slices_partition_the_frame(pkg):
    lib = pkg._ffi.rtb()
    for n in (0, 1, 255, 256, 257, 1200 * 675, 7680 * 4320, 12345677):
        for world in (1, 2, 3, 4, 8, 16):
            for root in {0, world - 1, world // 2}:
                edges = []
                for rank in range(world):
                    b, e = C.c_uint64(), C.c_uint64()
                    assert lib.rtb_exchange_slice(n, world, rank, root, C.byref(b), C.byref(e)) == 0
                    assert b.value <= e.value <= n and (b.value % 256 == 0 or b.value == n)
                    edges.append((b.value, e.value))
                assert edges[0][0] == 0 and edges[-1][1] == n
                assert all(edges[k][1] == edges[k + 1][0] for k in range(world - 1))  # contiguous, disjoint, complete
                sizes = [e - b for b, e in edges]
                # the root's NVLink ingress is the bottleneck: it combines nothing (world >= 3) or everything (world 2)
                if world >= 3:
                    assert sizes[root] == 0 and max(sizes) <= (n + world - 2) // (world - 1) + 255
                elif world == 2:
                    assert sizes[root] == n
    b, e = C.c_uint64(), C.c_uint64()
    assert lib.rtb_exchange_slice(100, 2, 2, 0, C.byref(b), C.byref(e)) == pkg.RTB_ERR_INVALID_ARGUMENT
    assert lib.rtb_exchange_slice(100, 2, 0, 2, C.byref(b), C.byref(e)) == pkg.RTB_ERR_INVALID_ARGUMENT
    assert lib.rtb_exchange_slice(100, 0, 0, 0, C.byref(b), C.byref(e)) == pkg.RTB_ERR_INVALID_ARGUMENT


def _expected(parts, spp, orc):
    total = parts[0].copy()
    for q in parts[1:]:
        total[:, :3] += q[:, :3]          # float32 adds in rank order, like the kernel
    total[:, 3] = spp
    return total, orc.resolve(total)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8, 16])
def test_exchange_resolve_kernel_matches_numpy_and_oracle(pkg, orc, world):
    import torch
    lib = pkg._ffi.rtb()
    n = 40 * 1000 + 123
    rng = np.random.default_rng(world)
    parts = [np.concatenate([rng.random((n, 3), np.float32) * 40, np.full((n, 1), 7 + r, np.float32)], axis=1)
             for r in range(world)]
    bufs = [torch.from_numpy(q).cuda() for q in parts]
    out_acc = torch.full((n, 4), -1.0, device="cuda")
    out_rgba = torch.zeros(n, 4, dtype=torch.uint8, device="cuda")
    peers = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
    spp = 64.0
    for rank in range(world):             # every "rank" combines its slice into the root's buffers
        assert lib.rtb_exchange_resolve(peers, world, rank, 0, out_acc.data_ptr(), out_rgba.data_ptr(), n, spp, 0, None) == 0
    torch.cuda.synchronize()
    want_acc, want_rgba = _expected(parts, spp, orc)
    assert np.array_equal(out_acc.cpu().numpy(), want_acc)
    assert np.array_equal(out_rgba.cpu().numpy(), want_rgba)
    # another root (different slices, same result), in place on the root's own buffer (root_accum_out == peer_accum[root])
    root = world - 1
    for rank in range(world):
        assert lib.rtb_exchange_resolve(peers, world, rank, root, bufs[root].data_ptr(), out_rgba.data_ptr(), n, spp, 0,
                                        None) == 0
    torch.cuda.synchronize()
    assert np.array_equal(bufs[root].cpu().numpy(), want_acc)
    # argument checking
    assert lib.rtb_exchange_resolve(peers, world, world, 0, out_acc.data_ptr(), out_rgba.data_ptr(), n, spp, 0, None) \
        == pkg.RTB_ERR_INVALID_ARGUMENT
    assert lib.rtb_exchange_resolve(peers, world, 0, 0, out_acc.data_ptr(), out_rgba.data_ptr(), n, 0.0, 0, None) \
        == pkg.RTB_ERR_INVALID_ARGUMENT


def _ipc_worker(rank, world, port, backend, n, spp, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    mg = importlib.import_module("zig-raytracing-weekend_b200.multigpu")
    device = rank if backend == "nccl" else 0
    torch.cuda.set_device(device)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    px = mg.PeerExchange(n, rank, world, device)
    stream = torch.cuda.current_stream().cuda_stream
    for step in range(3):                 # several rounds: the barriers must keep the rounds apart
        rng = np.random.default_rng(100 * step + rank)
        mine = np.concatenate([rng.random((n, 3), np.float32) * 9, np.full((n, 1), 5.0, np.float32)], axis=1)
        px.accum.copy_(torch.from_numpy(mine).cuda())
        px.exchange(spp, stream)
        if rank == 0:
            torch.cuda.synchronize()
            np.save(os.path.join(out_dir, f"acc{step}.npy"), px.accum.cpu().numpy())
            np.save(os.path.join(out_dir, f"rgba{step}.npy"), px.rgba.cpu().numpy())
    px.close()
    dist.destroy_process_group()


def _run_ipc(tmp_path, orc, world, backend):
    import torch.multiprocessing as mp
    n, spp = 300 * 200 + 17, 32.0
    port = _free_port()
    mp.start_processes(_ipc_worker, args=(world, port, backend, n, spp, str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    for step in range(3):
        parts = []
        for rank in range(world):
            rng = np.random.default_rng(100 * step + rank)
            parts.append(np.concatenate([rng.random((n, 3), np.float32) * 9, np.full((n, 1), 5.0, np.float32)], axis=1))
        want_acc, want_rgba = _expected(parts, spp, orc)
        assert np.array_equal(np.load(tmp_path / f"acc{step}.npy"), want_acc)
        assert np.array_equal(np.load(tmp_path / f"rgba{step}.npy"), want_rgba)


@pytest.mark.gpu
def test_peer_exchange_two_processes_one_gpu(pkg, orc, tmp_path):
    _run_ipc(tmp_path, orc, 2, "gloo")


@pytest.mark.gpu
def test_peer_exchange_nccl_one_rank_per_gpu(pkg, orc, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    _run_ipc(tmp_path, orc, min(torch.cuda.device_count(), 8), "nccl")
