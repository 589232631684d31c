"""Discriminators the DP engine drops in behind.

Host-side mirror of the reference's critic definitions (only the D side: the
generators are trained non-privately and never see the engine).
Parameter names (`lin1`, `lin2`, `linOutAux`, `blocks.N`, `linOut`) and
forward/loss semantics follow the reference so checkpoints and golden vectors
line up:

  * abstract critic + ACGAN / WCGAN aux loss ... reference models.py:23-67
  * vanilla MNIST critic ....................... reference MNIST_models.py:28-52
  * strided-conv "DCResNet" critic ............. reference DCResNet_models.py:109-153
  * CelebA 64/48 and MNIST presets ............. reference CelebA_models.py:14-24,
                                                 MNIST_models.py:58-60

Only `nn.Linear` and `nn.Conv2d(k=5, s=2, p=2)` appear, with functional
ReLU / LeakyReLU(0.2) between them, which is exactly the layer set the
per-sample-gradient kernels cover.

Unlike the reference, channel lists are copied on entry, so building two
conditional critics in one process does not mutate a shared default
(SURVEY.md §4 "test-harness trap").
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn.functional as F
from torch import nn


def _onehot(y: torch.Tensor, n: int) -> torch.Tensor:
    return F.one_hot(y, num_classes=n)


class Critic(nn.Module):
    """Common conditional plumbing and the auxiliary-classifier loss."""

    def __init__(self, n_classes: int = 0, emb_mode: Optional[str] = "concat",
                 conditional_arch: str = "CGAN", aux_loss_type: str = "wasserstein",
                 aux_loss_scalar: float = 1.0):
        super().__init__()
        self.n_classes = n_classes
        self.emb_mode = emb_mode
        self.conditional_arch = conditional_arch
        self.aux_loss_type = aux_loss_type
        self.aux_loss_scalar = aux_loss_scalar
        if n_classes > 1:
            if emb_mode == "embed":
                raise NotImplementedError("label embedding is not defined for critics")
            if conditional_arch == "ACGAN":
                # ACGAN critics never see the label on the input side.
                self.emb_mode = None
                if aux_loss_type == "cross_entropy":
                    self.aux_criterion = nn.CrossEntropyLoss()

    @property
    def is_conditional(self) -> bool:
        return self.n_classes > 1

    def aux_loss(self, output, labels, device=None, fake: bool = False):
        if self.conditional_arch == "ACGAN":
            if self.aux_loss_type == "cross_entropy":
                return self.aux_loss_scalar * self.aux_criterion(output, labels)
            # "wasserstein" aux: signed sigmoid score normalised by class count in the batch
            hot = _onehot(labels, self.n_classes)
            sign = hot * (-2) + 1
            per_class = hot.sum(dim=0)[labels].unsqueeze(1).expand_as(output)
            return self.aux_loss_scalar * torch.sum(sign * torch.sigmoid(output) / per_class)
        if self.conditional_arch == "WCGAN":
            return torch.tensor([0], device=device)
        return None


class MNISTVanillaD(Critic):
    """784(+classes) -> 128 -> 1 (+ aux head), BCE-with-logits adversarial loss."""

    def __init__(self, **kw):
        super().__init__(**kw)
        if self.is_conditional and self.aux_loss_type != "cross_entropy":
            raise ValueError("the vanilla critic supports only the cross_entropy aux loss")
        self.criterion = nn.BCEWithLogitsLoss()
        self.lin1 = nn.Linear(784 + self.n_classes, 128)
        self.lin2 = nn.Linear(128, 1)
        if self.is_conditional:
            self.linOutAux = nn.Linear(128, self.n_classes) if self.conditional_arch == "ACGAN" else None

    def forward(self, x, y=None, aux: bool = True):
        h = x.reshape(x.size(0), -1)
        if y is not None:
            h = torch.cat([h, _onehot(y, self.n_classes)], dim=1)
        h = F.relu(self.lin1(h))
        want_aux = aux and self.conditional_arch == "ACGAN" and self.is_conditional
        return self.lin2(h), (self.linOutAux(h) if want_aux else None)

    def real_loss(self, output, device=None):
        return self.criterion(output, torch.ones_like(output))

    def fake_loss(self, output, device=None):
        return self.criterion(output, torch.zeros_like(output))


class DCResNetD(Critic):
    """Stack of 5x5 stride-2 convolutions + linear head(s), Wasserstein loss."""

    def __init__(self, channels: Sequence[int], last_filter_size: int, **kw):
        super().__init__(**kw)
        ch = list(channels)
        if self.emb_mode == "concat" and self.is_conditional:
            ch[0] += self.n_classes
        self.blocks = nn.ModuleList(
            nn.Conv2d(cin, cout, 5, stride=2, padding=2) for cin, cout in zip(ch[:-1], ch[1:]))
        feat = ch[-1] * last_filter_size ** 2
        if not self.is_conditional or self.conditional_arch != "WCGAN":
            self.linOut = nn.Linear(feat, 1, bias=False)
        if self.is_conditional and self.conditional_arch in ("ACGAN", "WCGAN"):
            self.linOutAux = nn.Linear(feat, self.n_classes)

    def forward(self, x, y=None, aux: bool = True):
        h = x
        if self.emb_mode == "concat" and self.is_conditional:
            planes = _onehot(y, self.n_classes).view(x.size(0), -1, 1, 1)
            h = torch.cat((x, planes.expand(-1, -1, x.size(2), x.size(3))), dim=1)
        for conv in self.blocks:
            h = F.leaky_relu(conv(h), 0.2)
        h = h.reshape(x.size(0), -1)
        out_aux = self.linOutAux(h) if aux and hasattr(self, "linOutAux") else None
        if out_aux is not None and self.conditional_arch == "WCGAN":
            out = (out_aux * _onehot(y, self.n_classes)).sum(dim=1)
        else:
            out = self.linOut(h)
        return out, out_aux

    def real_loss(self, output, device=None):
        return -torch.mean(output)

    def fake_loss(self, output, device=None):
        return torch.mean(output)


class CelebA_DCRN_D64(DCResNetD):
    def __init__(self, channels=(3, 64, 128, 256, 512), last_filter_size=4, **kw):
        super().__init__(channels, last_filter_size, **kw)


class CelebA_DCRN_D48(DCResNetD):
    def __init__(self, channels=(3, 128, 256, 512), last_filter_size=6, **kw):
        super().__init__(channels, last_filter_size, **kw)


class MNIST_DCRN_D(DCResNetD):
    def __init__(self, channels=(1, 64, 128), last_filter_size=7, n_classes=10, **kw):
        super().__init__(channels, last_filter_size, n_classes=n_classes, **kw)


def build_discriminator(dataset: str, model: str, *, n_classes: int = 0, im_size: int = 64,
                        emb_mode: str = "concat", conditional_arch: str = "ACGAN",
                        aux_loss_type: Optional[str] = None, aux_loss_scalar: float = 1.0,
                        weights_seed: Optional[int] = 42, device="cpu") -> Critic:
    """Factory mirroring reference init_util.py:44-69 (D side only): seeds torch with
    `weights_seed` before construction so the initial weights are reproducible."""
    if dataset == "MNIST":
        cls = {"Vanilla": MNISTVanillaD, "DeepConvResNet": MNIST_DCRN_D}[model]
        aux_loss_type = aux_loss_type or "cross_entropy"
    elif dataset == "CelebA":
        if model != "DeepConvResNet":
            raise ValueError("No vanilla architecture for CelebA.")
        cls = CelebA_DCRN_D48 if im_size == 48 else CelebA_DCRN_D64
        aux_loss_type = aux_loss_type or "wasserstein"
    else:
        raise ValueError(f"unknown dataset {dataset!r}")
    if weights_seed is not None:
        torch.manual_seed(weights_seed)
    d = cls(n_classes=n_classes, emb_mode=emb_mode, conditional_arch=conditional_arch,
            aux_loss_type=aux_loss_type, aux_loss_scalar=aux_loss_scalar)
    return d.to(device)
