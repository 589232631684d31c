"""ISPrivacyEngine: immediate-sensitivity engine behind `--dp_mode is`.

Drop-in for the twosixlabs/opacus fork's `ISPrivacyEngine` as reference train.py drives it
(constructor :103-107, attach/_set_seed :135-136, backward(loss, inputs) :457/:469,
batch_sensitivity :332-338, scaling_vec / set_scaling_vec :247-249, patched step :484).

Semantics are INFERRED from the call sites and the published definition of immediate
sensitivity:
    g   = grad_theta loss                      (create_graph; p.grad = g, read at train.py:249)
    s   = max_i || d/dx_i ||g||_2 ||_2          flat            (optionally ||g_k / v_k|| with a
                                                                 per-parameter scaling vector)
    s_k = max_i || d/dx_i ||g_k||_2 ||_2        per parameter   (per_param / -ispp)
    step: p.grad += N(0, (sigma * s_k [* v_k])^2)                (`noise_div_batch` divides by B)
The double backward stays in autograd (cuDNN); the per-sample row norm over [B, C*H*W], the
max over B and the Philox noise run in the CUDA kernels behind the C ABI, and the sensitivities
stay on the device until somebody reads `batch_sensitivity` (logging).
"""
from __future__ import annotations

import ctypes as C
import types
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import autograd, nn

from . import _lib as L
from .accountant import compute_rdp, get_privacy_spent as _rdp_to_eps
from .functional import row_l2_norm, vec_max


class ISPrivacyEngine:
    def __init__(self, module: nn.Module, *, batch_size: int, sample_size: int,
                 alphas: Sequence[float] = tuple([1 + x / 10.0 for x in range(1, 100)] + list(range(12, 64))),
                 noise_multiplier: float, per_param: bool = False,
                 scaling_vec: Optional[Sequence[float]] = None, noise_div_batch: bool = False,
                 process_group=None, data_parallel: bool = False, global_batch_size: Optional[int] = None, **misc):
        """`batch_size` is THIS rank's batch; the accountant's sampling rate uses the global batch
        (`global_batch_size`, or the sum over ranks taken once at construction under data parallelism)."""
        L.load()
        self.module = module
        self.batch_size = batch_size
        self.sample_size = sample_size
        self.global_batch_size = global_batch_size
        self.alphas = list(alphas)
        self.noise_multiplier = float(noise_multiplier)
        self.per_param = per_param
        self.noise_div_batch = noise_div_batch
        self.process_group = process_group
        self.data_parallel = data_parallel or process_group is not None
        self.misc_settings = misc
        self.steps = 0
        self.optimizer = None
        self._params = [p for p in module.parameters() if p.requires_grad]
        self.device = self._params[0].device
        if self.device.type != "cuda":
            raise L.CslGanCudaError(
                f"ISPrivacyEngine needs the module on a CUDA device (found {self.device}); there is no CPU path")
        self.scaling_vec = None
        if scaling_vec is not None:
            self.set_scaling_vec(scaling_vec)
        self._sens_dev: Optional[torch.Tensor] = None
        self._per_sample_sens: Optional[torch.Tensor] = None
        self._seed = int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)
        self._philox_offset = 0
        if self.global_batch_size is None:
            self.global_batch_size = batch_size
            if self.data_parallel:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    from .dist import global_batch_size as _gbs
                    self.global_batch_size = _gbs(batch_size, self.device, self.process_group)

    @property
    def sample_rate(self) -> float:
        return self.global_batch_size / self.sample_size

    # ------------------------------------------------------------------ attach / seed
    def attach(self, optimizer):
        self.optimizer = optimizer
        engine = self

        def dp_step(opt_self, closure=None):
            engine.step()
            return opt_self.original_step(closure)

        optimizer.privacy_engine = self
        optimizer.original_step = optimizer.step
        optimizer.step = types.MethodType(dp_step, optimizer)

    def detach(self):
        opt = self.optimizer
        if opt is not None:
            opt.step = opt.original_step
            del opt.privacy_engine, opt.original_step
            self.optimizer = None

    def _set_seed(self, seed: int):
        self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._philox_offset = 0
        if getattr(self, "_offset_dev", None) is not None:
            self._offset_dev.zero_()

    def enable_graph_safe_rng(self):
        """Philox offset in device memory: a CUDA-graph-captured step draws fresh noise at every replay."""
        if getattr(self, "_offset_dev", None) is None:
            self._offset_dev = torch.tensor([self._philox_offset], dtype=torch.int64, device=self.device)

    @property
    def philox_offset(self) -> int:
        od = getattr(self, "_offset_dev", None)
        return int(od.item()) if od is not None else self._philox_offset

    def state_dict(self) -> Dict:
        return {"steps": self.steps, "seed": self._seed, "philox_offset": self.philox_offset,
                "scaling_vec": self.scaling_vec}

    def load_state_dict(self, sd: Dict):
        self.steps, self._seed, self._philox_offset = sd["steps"], sd["seed"], sd["philox_offset"]
        if getattr(self, "_offset_dev", None) is not None:
            self._offset_dev.fill_(self._philox_offset)      # graph-safe RNG: resume where the stream left off
        if sd.get("scaling_vec") is not None:
            self.set_scaling_vec(sd["scaling_vec"])

    def set_scaling_vec(self, vec: Sequence[float]):
        vec = [float(v) for v in vec]
        if len(vec) != len(self._params):
            raise ValueError(f"scaling_vec needs {len(self._params)} entries, got {len(vec)}")
        self.scaling_vec = vec

    # ------------------------------------------------------------------ backward
    def backward(self, loss: torch.Tensor, inputs: torch.Tensor):
        """Compute p.grad and the batch's immediate sensitivity (reference train.py:457, 469)."""
        if not inputs.requires_grad:
            raise RuntimeError("inputs must require grad (reference train.py:375 sets img.requires_grad = True)")
        params = self._params
        g = autograd.grad(loss, params, create_graph=True, allow_unused=True)
        g = [gi if gi is not None else torch.zeros_like(p) for gi, p in zip(g, params)]
        weight = None                                # this rank's share of the global batch (data parallel)
        if self.data_parallel:
            # the sensitivity is that of the GLOBAL mean gradient: every rank needs the global g before
            # ||g|| is differentiated (two exchange steps).  Shards may be unequal: the global mean
            # is sum_r B_r g_r / sum_r B_r, the sample count rides in the same collective.
            from .dist import allreduce_weighted_mean
            g_glob, weight = allreduce_weighted_mean([gi.detach() for gi in g], inputs.shape[0],
                                                     group=self.process_group)
        else:
            g_glob = [gi.detach() for gi in g]
        for p, gi in zip(params, g_glob):
            p.grad = gi.clone()

        def sens_of(norm_scalar: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
            if not norm_scalar.requires_grad:
                z = torch.zeros(inputs.shape[0], device=self.device)
                return z, torch.zeros(1, device=self.device)
            sx = autograd.grad(norm_scalar, inputs, retain_graph=True, allow_unused=True)[0]
            if sx is None:
                z = torch.zeros(inputs.shape[0], device=self.device)
                return z, torch.zeros(1, device=self.device)
            per_sample = row_l2_norm(sx.detach())
            return per_sample, vec_max(per_sample)

        def norm_of(local: torch.Tensor, glob: torch.Tensor) -> torch.Tensor:
            """||g_global|| as a function of this rank's inputs: with v = g_global/||g_global|| held
            constant, d||g_global||/dx_i = weight * <v, d g_local / dx_i>."""
            if weight is None:
                return row_l2_norm(local.reshape(1, -1)).sum()
            from .dist import global_norm_proxy
            return global_norm_proxy(local.reshape(-1), glob.reshape(-1), weight)

        if self.per_param:
            rows, maxes = [], []
            for gi, gg in zip(g, g_glob):
                ps, mx = sens_of(norm_of(gi, gg))
                rows.append(ps)
                maxes.append(mx)
            self._per_sample_sens = torch.stack(rows)
            sens = torch.cat(maxes)
        else:
            sv = self.scaling_vec
            if weight is None:
                # ||g|| over all parameters without concatenating them: sqrt(sum_k ||g_k||^2 / v_k^2); every
                # ||g_k|| is one multi-block CUDA row norm with its own backward / double backward
                nk = torch.stack([row_l2_norm(gi.reshape(1, -1)).sum() / (1.0 if sv is None else sv[k])
                                  for k, gi in enumerate(g)])
                total = nk.pow(2).sum().sqrt() if nk.requires_grad else nk.detach().pow(2).sum().sqrt()
                ps, sens = sens_of(total)
            else:
                if sv is None:
                    loc = torch.cat([gi.reshape(-1) for gi in g])
                    glo = torch.cat([gi.reshape(-1) for gi in g_glob])
                else:
                    loc = torch.cat([gi.reshape(-1) / v for gi, v in zip(g, sv)])
                    glo = torch.cat([gi.reshape(-1) / v for gi, v in zip(g_glob, sv)])
                ps, sens = sens_of(norm_of(loc, glo))
            self._per_sample_sens = ps
        if self.data_parallel:
            import torch.distributed as dist
            dist.all_reduce(sens, op=dist.ReduceOp.MAX, group=self.process_group)
        self._sens_dev = sens

    @property
    def batch_sensitivity(self):
        """float (flat) or np.ndarray (per_param): reference train.py:332-338.  Host sync happens here,
        not in backward()."""
        if self._sens_dev is None:
            raise RuntimeError("backward() has not been called")
        v = self._sens_dev.detach().cpu().numpy()
        return v.astype(np.float64) if self.per_param else float(v[0])

    # ------------------------------------------------------------------ step
    def step(self):
        """p.grad += N(0, (sigma * s_k)^2), drawn per parameter tensor from the Philox stream."""
        if self._sens_dev is None:
            raise RuntimeError("step() before backward()")
        self.steps += 1
        st = L.stream_ptr(self.device)
        ndiv = float(self.global_batch_size) if self.noise_div_batch else 0.0
        od = getattr(self, "_offset_dev", None)
        segs = []
        for k, p in enumerate(self._params):
            mult = self.noise_multiplier
            if self.per_param:
                sdev = self._sens_dev[k:k + 1]
            else:
                sdev = self._sens_dev[:1]
                if self.scaling_vec is not None:
                    mult *= self.scaling_vec[k]
            segs.append((p.grad, p.grad, mult, sdev))
        # ONE launch for all parameter tensors (bit-identical to one torch.normal per tensor, in order)
        inc = L.noise_multi(segs, 0.0, None, ndiv, None, self._seed, 0 if od is not None else self._philox_offset, od, st)
        if od is not None:
            if inc:
                L.call("cg_philox_advance", L.ptr(od), inc, st)
        else:
            self._philox_offset += inc

    # ------------------------------------------------------------------ accountant
    def get_privacy_spent(self, target_delta: Optional[float] = None):
        if target_delta is None:
            target_delta = self.misc_settings.get("target_delta", 1e-6)
        rdp = compute_rdp(self.sample_rate, self.noise_multiplier, self.steps, self.alphas)
        return _rdp_to_eps(self.alphas, rdp, target_delta)
