"""torchrun probe: cost of the fused exchange + noise kernel (barrier, kernel, barrier) over a 17.3 MB gradient for the
four transfer routes, against NCCL allreduce + the single-GPU noise kernel.  usage: torchrun --nproc-per-node N ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from csl_gan_b200 import _lib as L
from csl_gan_b200.dist import SymmetricFlat

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
sizes = [4800, 64, 204800, 128, 819200, 256, 3276800, 512, 8192, 1]
n = sum(sizes)
os.environ["CSLGAN_XFER"] = "mix"
sy = SymmetricFlat(n + 1, dev)
st = L.stream_ptr(dev)

def segs_of(flat):
    out, off = [], 0
    for s in sizes:
        v = flat[off:off + s]
        out.append((v, v, 0.5, None))
        off += s
    return out

def run(mode, i=0):
    flat = sy.bufs[i]
    if mode == "p2p":
        mc, peers = 0, sy.peer_ptrs[i]
    elif mode == "multimem":
        mc, peers = sy.mc_ptrs[i], None
    else:
        mc, peers = sy.mc_ptrs[i], sy.peer_ptrs[i]
        os.environ["CSLGAN_XFER_MIX"] = "1" if mode == "mc-load" else "2"
    sy.barrier(i)
    L.noise_multi_allreduce(segs_of(flat), True, 1234, 0, None, flat, mc, peers, n, rank, world, st)
    sy.barrier(i)

def timeit(name, fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{name:34s} {e0.elapsed_time(e1) * 1e3 / iters:8.1f} us", flush=True)

plain = torch.zeros(n + 1, device=dev)
def nccl():
    dist.all_reduce(plain)
    L.noise_multi(segs_of(plain), 512.0 * world, None, 512.0 * world, None, 1234, 0, None, st)
timeit("nccl allreduce + noise kernel", nccl)
ref = None
for mode in ("p2p", "multimem", "mc-load", "mc-store"):
    # correctness: every rank contributes rank+1, count 512 -> (sum / (512*world)) + noise; all routes must agree
    sy.bufs[0].fill_(float(rank + 1)); sy.bufs[0][n:].fill_(512.0)
    torch.cuda.synchronize(); dist.barrier()
    run(mode)
    torch.cuda.synchronize()
    got = sy.bufs[0][:n].clone()
    if ref is None:
        ref = got
    ok = torch.equal(got, ref)
    timeit(f"fused {mode} (2 barriers incl.)", lambda m=mode: run(m, 1))
    if rank == 0:
        print(f"   {mode}: identical to p2p result: {ok}", flush=True)
torch.cuda.synchronize()
os._exit(0)
