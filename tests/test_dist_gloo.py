"""world_size-2 gloo checks of the data-parallel host logic on CPU: sharding the batch and exchanging
ONE allreduce reproduces the single-process result, for gc (clipped sums) and for the two-exchange
immediate-sensitivity path.  The per-rank arithmetic is the CPU oracle here; the collective / sharding
helpers are the product's (csl_gan_b200/dist.py)."""
import copy
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from csl_gan_b200 import discriminators as DD
from csl_gan_b200.dist import (allreduce_flat, allreduce_sum_and_count, allreduce_weighted_mean, global_batch_size,
                               global_norm_proxy, shard_range)
from oracle import dp_oracle as O


def _data(B):
    g = torch.Generator().manual_seed(3)
    real = torch.rand(B, 1, 28, 28, generator=g)
    fake = torch.rand(B, 1, 28, 28, generator=g) * 0.5
    torch.manual_seed(42)
    D = DD.MNISTVanillaD(n_classes=0)
    return D, real, fake


def _gc_sums(D, real, fake, B, C):
    eng = O.OracleGCEngine(D, batch_size=B, noise_multiplier=0.0, max_grad_norm=C)
    (D.real_loss(D(real)[0]) + D.fake_loss(D(fake)[0])).backward()
    eng.clip(); eng.accum_grads_across_passes(); eng.accumulate_batch()
    out = [p.summed_grad.clone() for p in eng.params()]
    eng.remove()
    return out


def _worker(rank, world, port, B, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    D, real, fake = _data(B)
    lo, hi = shard_range(B, rank, world)
    # ---- gc: per-rank clipped sums + ONE allreduce == full batch; the live sample count rides in the same flat
    # buffer, so unequal shards (B % world != 0) divide by the true global batch on every rank
    local = _gc_sums(copy.deepcopy(D), real[lo:hi], fake[lo:hi], hi - lo, 1.5)
    flat = torch.cat([t.reshape(-1) for t in local] + [torch.tensor([float(hi - lo)])])
    allreduce_sum_and_count(flat)
    assert flat[-1].item() == B and global_batch_size(hi - lo) == B
    red, off = [], 0
    for t in local:
        red.append(flat[off:off + t.numel()].view(t.shape))
        off += t.numel()
    full = _gc_sums(copy.deepcopy(D), real, fake, B, 1.5)
    gc_err = max(((a - b).norm() / b.norm()).item() for a, b in zip(red, full))
    gc_err = max(gc_err, max(((a / flat[-1] - b / B).norm() / (b / B).norm()).item() for a, b in zip(red, full)))
    # ---- is: allreduce(mean) of g, local second-order pass through the proxy, allreduce(MAX)
    Dl = copy.deepcopy(D)
    x = real[lo:hi].clone().requires_grad_(True)
    loss = Dl.real_loss(Dl(x)[0]) + Dl.fake_loss(Dl(fake[lo:hi])[0])
    params = list(Dl.parameters())
    g = torch.autograd.grad(loss, params, create_graph=True)
    # per-rank losses are means over shards of unequal size: the global mean gradient is sum_r B_r g_r / sum_r B_r
    g_glob, weight = allreduce_weighted_mean([t.detach() for t in g], hi - lo)
    proxy = global_norm_proxy(torch.cat([t.reshape(-1) for t in g]), torch.cat([t.reshape(-1) for t in g_glob]), weight)
    sx = torch.autograd.grad(proxy, x)[0]
    s = O.row_l2_norm(sx).max().reshape(1)
    dist.all_reduce(s, op=dist.ReduceOp.MAX)
    Df = copy.deepcopy(D)
    xf = real.clone().requires_grad_(True)
    lf = Df.real_loss(Df(xf)[0]) + Df.fake_loss(Df(fake)[0])
    g_ref, s_ref, _ = O.immediate_sensitivity(list(Df.parameters()), lf, xf)
    g_err = max(((a - b).norm() / b.norm()).item() for a, b in zip(g_glob, g_ref))
    if rank == 0:
        ret.put((gc_err, g_err, abs(s.item() - s_ref) / s_ref))
    dist.destroy_process_group()


def test_shard_range_partitions_the_batch():
    for n, w in [(600, 8), (128, 8), (7, 2), (5, 8)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


@pytest.mark.timeout(300)
def test_two_rank_sharded_step_equals_single_process():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, ret)) for r in range(2)]     # 4 + 3: unequal shards
    for p in procs:
        p.start()
    gc_err, g_err, s_err = ret.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert gc_err < 1e-5, gc_err
    assert g_err < 1e-5, g_err
    assert s_err < 1e-4, s_err
