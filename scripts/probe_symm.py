"""Probe (torchrun, N GPUs): symmetric-memory / multicast availability and the cost of a 17.3 MB fp32 allreduce through
NCCL vs torch's symmetric-memory kernels vs a bare barrier.  usage: torchrun --nproc-per-node N scripts/probe_symm.py"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n = 4_340_000
t = symm.empty(n, dtype=torch.float32, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD)
if rank == 0:
    print("multicast_ptr", hex(hdl.multicast_ptr), "signal pad", hdl.signal_pad_size, "world", hdl.world_size, flush=True)
gname = dist.group.WORLD.group_name
plain = torch.zeros(n, device=dev)

def timeit(name, fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    if rank == 0:
        print(f"{name:32s} {us:8.1f} us", flush=True)

timeit("nccl all_reduce (plain buffer)", lambda: dist.all_reduce(plain))
timeit("nccl all_reduce (symm buffer)", lambda: dist.all_reduce(t))
for name in ("multimem_all_reduce_", "two_shot_all_reduce_"):
    try:
        op = getattr(torch.ops.symm_mem, name)
        timeit(name, lambda: op(t, "sum", gname))
    except Exception as e:
        if rank == 0:
            print(name, "failed:", repr(e)[:200], flush=True)
try:
    timeit("one_shot_all_reduce", lambda: torch.ops.symm_mem.one_shot_all_reduce(t, "sum", gname))
except Exception as e:
    if rank == 0:
        print("one_shot failed", repr(e)[:200])
timeit("hdl.barrier()", lambda: hdl.barrier(channel=0))
# correctness of multimem on a known pattern
t.fill_(float(rank + 1)); torch.cuda.synchronize(); dist.barrier()
try:
    torch.ops.symm_mem.multimem_all_reduce_(t, "sum", gname); torch.cuda.synchronize()
    if rank == 0:
        print("multimem result", t[:2].tolist(), "expect", world * (world + 1) / 2)
except Exception as e:
    if rank == 0:
        print("multimem check failed", repr(e)[:200])
# CUDA-graph capture of barrier + allreduce
try:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        torch.ops.symm_mem.multimem_all_reduce_(t, "sum", gname)
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        hdl.barrier(channel=0)
        torch.ops.symm_mem.multimem_all_reduce_(t, "sum", gname)
    timeit("graph(barrier + multimem ar)", g.replay)
except Exception as e:
    if rank == 0:
        print("graph capture failed", repr(e)[:300])
torch.cuda.synchronize()
os._exit(0)
