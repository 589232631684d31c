"""Device-resident sampler of the noisy class-mean "public" images that feed adaptive clipping and the
gradient penalty every D step.

Mirror of reference mean_sampler.py (`sample` :75-84, `make_mean_samples` :47-64, `get_privacy_cost`
:86-92; called at train.py:200, 214).  The reference builds every public batch on the CPU (randperm +
two normal_ draws) and copies it to the device on the step's critical path; here the mean images live
on the device and `sample` issues a handful of device ops with no host synchronisation.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch

from .accountant import compute_rdp, get_privacy_spent

ALPHAS = [1 + x / 10.0 for x in range(1, 100)] + list(range(12, 400))


class DeviceMeanSampler:
    def __init__(self, mean_samples: torch.Tensor, noise_std: float, mean_size: int, dataset_size: int,
                 smallest_class_size: Optional[int] = None, generator: Optional[torch.Generator] = None):
        """mean_samples: [n_classes, num_samples, C, H, W] (already noised, reference :63-64)."""
        if mean_samples.dim() != 5:
            raise ValueError("mean_samples must be [n_classes, num_samples, C, H, W]")
        self.mean_samples = mean_samples
        self.n_classes, self.num_samples = mean_samples.shape[:2]
        self.ch, self.res = mean_samples.shape[2], mean_samples.shape[3]
        self.noise_std, self.mean_size, self.dataset_size = noise_std, mean_size, dataset_size
        self.sample_rate = mean_size / (dataset_size if smallest_class_size is None else smallest_class_size)
        self.generator = generator

    @classmethod
    def from_batches(cls, batches: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor]]], n_classes: int,
                     mean_size: int, noise_std: float, dataset_size: int, device="cuda", **kw):
        """One noisy per-class mean per batch (reference make_mean_samples :47-64)."""
        per_class = [[] for _ in range(n_classes)]
        for x, y in batches:
            x = x.to(device)
            for c in range(n_classes):
                s = x[(y.to(device) == c)][:mean_size] if n_classes > 1 else x
                m = s.sum(dim=0) / mean_size
                per_class[c].append(m + torch.randn_like(m) * noise_std)
        return cls(torch.stack([torch.stack(v) for v in per_class]), noise_std, mean_size, dataset_size, **kw)

    def sample(self, size: int, noise_std: float = 0.01, noise_mean_std: float = 0.01,
               requested_labels: Optional[torch.Tensor] = None):
        dev, g = self.mean_samples.device, self.generator
        reps = (size - 1) // self.num_samples + 1
        perms = torch.cat([torch.randperm(self.num_samples, device=dev, generator=g) for _ in range(reps)])[:size]
        if requested_labels is None:
            requested_labels = torch.randint(0, self.n_classes, (size,), device=dev, generator=g)
        r = self.mean_samples[requested_labels.to(dev), perms].clone()
        if noise_mean_std is not None and noise_mean_std > 0:
            r += torch.randn(size, 1, 1, 1, device=dev, generator=g) * noise_mean_std
        if noise_std is not None and noise_std > 0:
            r += torch.randn(r.shape, device=dev, generator=g) * noise_std
        return r, (requested_labels if self.n_classes > 1 else None)

    def get_privacy_cost(self, target_delta: float = 1e-6, alphas=ALPHAS):
        pixel_sensitivity = 1 / self.mean_size / 2
        l2_sensitivity = math.sqrt(self.ch * self.res ** 2 * pixel_sensitivity ** 2)
        rdp = compute_rdp(self.sample_rate, self.noise_std / l2_sensitivity, self.num_samples * self.n_classes, alphas)
        return get_privacy_spent(alphas, rdp, target_delta)
