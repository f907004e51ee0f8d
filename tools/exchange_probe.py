#!/usr/bin/env python
"""Development aid (run under torchrun): times multigpu.PeerExchange.exchange and the NCCL reduce + resolve on an
8K frame with CUDA events; the kernel's tuning knobs come from the environment (RTB_XCHG_PPT, RTB_XCHG_BLOCKS_PER_SM)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
p = importlib.import_module("zig-raytracing-weekend_b200")
mg = importlib.import_module("zig-raytracing-weekend_b200.multigpu")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
npx = 7680 * 4320
px = mg.PeerExchange(npx, rank, world, lr)
px.accum.fill_(1.0)
stream = torch.cuda.current_stream().cuda_stream
acc2 = torch.ones(npx, 4, device="cuda")
rgba2 = torch.zeros(npx, 4, dtype=torch.uint8, device="cuda")
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
def nccl():
    dist.reduce(acc2, dst=0)
    if rank == 0:
        p._ffi.rtb().rtb_resolve_device(acc2.data_ptr(), rgba2.data_ptr(), npx, 64.0, lr, stream)
a = timed(lambda: px.exchange(64.0, stream))
b = timed(nccl)
if rank == 0:
    print(f"world {world} PPT={os.environ.get('RTB_XCHG_PPT','1')} BPS={os.environ.get('RTB_XCHG_BLOCKS_PER_SM','8')}: "
          f"p2p exchange {a:.3f} ms, nccl reduce+resolve {b:.3f} ms", flush=True)
px.close()
dist.destroy_process_group()
