"""Data-parallel plumbing for the DP D-step.  The reference has no distributed code;
this is the one exchange step the sharded path needs.

The batch is sharded over ranks (one process per GPU, NCCL over NVLink); every sample's gradient,
norm and clip factor are computed on the rank that holds it, so the only data-path collective is ONE
allreduce(SUM) of the flattened clipped-sum gradient per step (|theta| fp32: 17.3 MB for the CelebA
critic).  Noise is then drawn identically on every rank from the shared Philox (seed, offset) and
added after the reduce, so all replicas apply the same update.  Immediate sensitivity needs two
exchanges: allreduce(mean) of g before ||g|| is differentiated, and allreduce(MAX) of the sensitivities.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n samples for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_flat(tensors: Sequence[torch.Tensor], op=None, group=None) -> List[torch.Tensor]:
    """One collective for a list of tensors: flatten-concatenate, all_reduce, return views shaped like
    the inputs (latency, not bandwidth, is what matters at these sizes: one launch instead of nine)."""
    op = dist.ReduceOp.SUM if op is None else op
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=op, group=group)
    out, off = [], 0
    for t in tensors:
        n = t.numel()
        out.append(flat[off:off + n].view(t.shape))
        off += n
    return out


def global_norm_proxy(local_flat: torch.Tensor, global_flat: torch.Tensor, world: int) -> torch.Tensor:
    """A scalar whose gradient w.r.t. this rank's inputs equals d||g_global||/dx_i for the samples the
    rank holds: with v = g_global / ||g_global|| constant, d||g_global||/dx_i = <v, d g_local/dx_i> / world."""
    n = global_flat.norm(2)
    v = torch.where(n > 0, global_flat / n, torch.zeros_like(global_flat))
    return (local_flat * v.detach()).sum() / world
