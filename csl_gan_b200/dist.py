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


def global_norm_proxy(local_flat: torch.Tensor, global_flat: torch.Tensor, weight) -> torch.Tensor:
    """A scalar whose gradient w.r.t. this rank's inputs equals d||g_global||/dx_i for the samples the
    rank holds: with v = g_global / ||g_global|| constant, d||g_global||/dx_i = weight * <v, d g_local/dx_i>,
    where g_global = sum_r weight_r g_local_r.  `weight` is this rank's share of the global batch B_r / sum B
    (a float or a 0-dim tensor; 1/world for equal shards -- an int `world` is still accepted and means that)."""
    if isinstance(weight, int) and not isinstance(weight, bool) and weight >= 1:
        weight = 1.0 / weight
    n = global_flat.norm(2)
    v = torch.where(n > 0, global_flat / n, torch.zeros_like(global_flat))
    return (local_flat * v.detach()).sum() * weight


def allreduce_sum_and_count(flat: torch.Tensor, group=None) -> torch.Tensor:
    """allreduce(SUM) of a flat buffer whose LAST element carries this rank's live sample count: the clipped sums
    and the global batch size arrive in one collective, so unequal shards (B % world != 0, a short last batch on
    one rank) divide by the true number of samples on every rank.  In place; returns `flat`."""
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def allreduce_weighted_mean(tensors: Sequence[torch.Tensor], local_count: int, group=None):
    """Global mean gradient of per-rank MEAN gradients over shards of unequal size:
    g = sum_r B_r g_r / sum_r B_r.  One collective (the count rides in the flat buffer).
    Returns (list of tensors shaped like the inputs, this rank's weight B_r / sum B as a 0-dim tensor)."""
    flat = torch.cat([t.reshape(-1) * float(local_count) for t in tensors]
                     + [torch.full((1,), float(local_count), dtype=tensors[0].dtype, device=tensors[0].device)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    total = flat[-1]
    out, off = [], 0
    for t in tensors:
        n = t.numel()
        out.append((flat[off:off + n] / total).view(t.shape))
        off += n
    return out, float(local_count) / total


def global_batch_size(local_batch: int, device=None, group=None) -> int:
    """Sum of the per-rank batch sizes (one tiny collective at engine construction): the sampling rate the RDP
    accountant needs is global_batch / sample_size, not this rank's shard."""
    t = torch.tensor([float(local_batch)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(round(t.item()))


class SymmetricFlat:
    """Two flat fp32 buffers of `n` elements in symmetric (peer-mapped, NVSwitch-multicast) memory, one process per
    GPU: the clipped sums of a step are written straight into one of them, and `cg_noise_finalize_allreduce` then
    does the cross-rank sum, the noise and the broadcast of the finished gradient in ONE kernel over the multicast
    mapping (no NCCL call on the data path).  Two buffers so that gradient accumulation can hold a running sum in one
    while clip() fills the other.  Construction is collective (every rank of `group` must call it).

    `available()` is False when the group has one rank, the allocator or the multicast mapping is missing, or
    CSLGAN_FUSED_ALLREDUCE=0 -- the engine then keeps the NCCL allreduce."""

    def __init__(self, n: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.n = int(n)
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.bufs, self.hdls = [], []
        for _ in range(2):
            t = symm.empty(self.n, dtype=torch.float32, device=device)
            self.hdls.append(symm.rendezvous(t, self.group))
            self.bufs.append(t)
        self.mc_ptrs = [int(h.multicast_ptr) for h in self.hdls]
        self.peer_ptrs = [[int(x) for x in h.buffer_ptrs] for h in self.hdls]
        # exchange route: plain 16-byte loads / stores through the NVLink peer mappings (default up to 8 ranks) or
        # NVSwitch multicast (multimem.ld_reduce / multimem.st).  Measured on B200 for the 17.3 MB CelebA gradient
        # (kernel only; NCCL allreduce + separate noise kernel for comparison): 2 GPUs 40 us p2p / 100 us multimem /
        # 84 us NCCL+noise; 8 GPUs 66 / 147 / ~150 us.  CSLGAN_XFER=p2p|multimem overrides.
        import os
        want = os.environ.get("CSLGAN_XFER", "")
        self.use_multicast = all(self.mc_ptrs) and (want in ("multimem", "mix") or (want != "p2p" and self.world > 8))
        self.mix = want == "mix"                             # both mappings to the kernel (CSLGAN_XFER_MIX picks the halves)
        if not self.use_multicast and self.world > 8:
            raise RuntimeError("peer-to-peer exchange covers up to 8 ranks and there is no multicast mapping")

    @staticmethod
    def available(group=None) -> bool:
        import os
        if os.environ.get("CSLGAN_FUSED_ALLREDUCE", "1") == "0":
            return False
        if not (dist.is_available() and dist.is_initialized()):
            return False
        g = group if group is not None else dist.group.WORLD
        if dist.get_world_size(g) < 2 or dist.get_backend(g) != "nccl":
            return False
        try:
            import torch.distributed._symmetric_memory  # noqa: F401
        except Exception:
            return False
        return True

    def index_of(self, flat: Optional[torch.Tensor]) -> int:
        """Which of the two buffers `flat` is (by storage address), or -1."""
        if flat is None:
            return -1
        for i, b in enumerate(self.bufs):
            if flat.data_ptr() == b.data_ptr() and flat.numel() == b.numel():
                return i
        return -1

    def barrier(self, i: int):
        """Cross-rank barrier on the current stream (a small kernel over the signal pads; capturable)."""
        self.hdls[i].barrier(channel=0)
