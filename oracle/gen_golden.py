"""Generate golden vectors by importing the reference's OWN Python modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):   python oracle/gen_golden.py
Writes small .npz fixtures to tests/golden/.  TEST INFRASTRUCTURE ONLY.

What gets pinned (everything that exists in-tree in the reference for this path):
  * critics' forward + losses ........ MNIST_models.py:28-52, DCResNet_models.py:109-153,
                                       CelebA_models.py:14-24, models.py:23-67
  * micro-batch per-sample gradients of those critics' losses (ground truth for the
    grad-sample contraction; the fork's hooks are not in the tree)
  * per-sample input-gradient norm and penalties ... gradient_penalty.py:31-65
  * l2_clip ......................................... backprop_clip.py:18-22

`opacus` (util.py:3 unused import) and `torchinfo` (backprop_clip.py:4) are absent from
the image, so empty stub modules are injected in sys.modules before importing; no
reference source is modified or copied.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _stub_modules():
    op = types.ModuleType("opacus")
    op_utils = types.ModuleType("opacus.utils")
    us = types.ModuleType("opacus.utils.uniform_sampler")
    us.UniformWithReplacementSampler = object
    op.utils = op_utils
    op_utils.uniform_sampler = us
    sys.modules.update({"opacus": op, "opacus.utils": op_utils, "opacus.utils.uniform_sampler": us})
    ti = types.ModuleType("torchinfo")
    ti.summary = lambda *a, **k: None
    sys.modules["torchinfo"] = ti


def _micro(model, loss_of, B):
    params = list(model.parameters())
    gs = [np.zeros((B,) + tuple(p.shape), np.float32) for p in params]
    for i in range(B):
        g = torch.autograd.grad(loss_of(i), params, allow_unused=True)
        for k, gi in enumerate(g):
            if gi is not None:
                gs[k][i] = gi.numpy()
    return gs


def main():
    _stub_modules()
    sys.path.insert(0, REF)
    import MNIST_models as RM            # noqa: E402  (reference modules)
    import CelebA_models as RC           # noqa: E402
    import DCResNet_models as RD         # noqa: E402
    import gradient_penalty as RGP       # noqa: E402
    import backprop_clip as RBC          # noqa: E402

    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)

    cases = {
        # name: (ctor, kwargs, input shape, n_classes for labels, B)
        "mnist_vanilla_acgan": (RM.MNISTVanillaD, dict(n_classes=10, emb_mode="concat", conditional_arch="ACGAN",
                                                       aux_loss_type="cross_entropy", aux_loss_scalar=1), (1, 28, 28), 10, 6),
        "mnist_vanilla_uncond": (RM.MNISTVanillaD, dict(n_classes=0, emb_mode="concat", conditional_arch="ACGAN",
                                                        aux_loss_type="cross_entropy", aux_loss_scalar=1), (1, 28, 28), 0, 6),
        "mnist_dcrn_acgan": (RM.MNIST_DCRN_D, dict(channels=[1, 64, 128], n_classes=10, emb_mode="concat",
                                                   conditional_arch="ACGAN", aux_loss_type="wasserstein",
                                                   aux_loss_scalar=1), (1, 28, 28), 10, 3),
        "celeba_d64_uncond": (RC.CelebA_DCRN_D64, dict(channels=[3, 64, 128, 256, 512], n_classes=0, emb_mode="concat",
                                                       conditional_arch="ACGAN", aux_loss_type="wasserstein",
                                                       aux_loss_scalar=1), (3, 64, 64), 0, 3),
        "celeba_d64_cgan": (RC.CelebA_DCRN_D64, dict(channels=[3, 64, 128, 256, 512], n_classes=2, emb_mode="concat",
                                                     conditional_arch="CGAN", aux_loss_type="wasserstein",
                                                     aux_loss_scalar=1), (3, 64, 64), 2, 2),
        "celeba_d48_uncond": (RC.CelebA_DCRN_D48, dict(channels=[3, 128, 256, 512], n_classes=0, emb_mode="concat",
                                                       conditional_arch="ACGAN", aux_loss_type="wasserstein",
                                                       aux_loss_scalar=1), (3, 48, 48), 0, 2),
    }

    for name, (ctor, kw, ishape, ncls, B) in cases.items():
        torch.manual_seed(42)                       # reference init_util.py:63 (weights_seed default)
        D = ctor(**kw)
        g = torch.Generator().manual_seed(1234)
        x = torch.rand((B,) + ishape, generator=g)
        if ishape[0] == 3:
            x = x * 2 - 1
        y = torch.randint(0, ncls, (B,), generator=g) if ncls > 1 else None
        out, aux = D(x, y)
        real_loss = D.real_loss(out, "cpu")
        fake_loss = D.fake_loss(out, "cpu")
        use_aux = aux is not None and kw["aux_loss_type"] == "cross_entropy"   # separable per-sample mean only
        aux_loss = D.aux_loss(aux, y, "cpu") if aux is not None else None

        def loss_of(i, D=D, x=x, y=y, use_aux=use_aux):
            o, a = D(x[i:i + 1], None if y is None else y[i:i + 1])
            l = D.real_loss(o, "cpu")
            if use_aux:
                l = l + D.aux_loss(a, y[i:i + 1], "cpu")
            return l

        gs = _micro(D, loss_of, B)
        names = [n for n, _ in D.named_parameters()]
        rec = {
            "param_names": np.array(names),
            "x": x.numpy(), "y": (y.numpy() if y is not None else np.zeros(0, np.int64)),
            "out": out.detach().numpy(),
            "aux": (aux.detach().numpy() if aux is not None else np.zeros(0, np.float32)),
            "real_loss": real_loss.detach().numpy(), "fake_loss": fake_loss.detach().numpy(),
            "aux_loss": (aux_loss.detach().numpy() if aux_loss is not None else np.zeros(0, np.float32)),
            "micro_uses_aux": np.array(use_aux),
            "micro_norms": np.stack([np.sqrt((gk.reshape(B, -1).astype(np.float64) ** 2).sum(1)) for gk in gs]).astype(np.float32),
            "param_sums": np.array([p.detach().double().sum().item() for p in D.parameters()]),
            "param_abs_sums": np.array([p.detach().double().abs().sum().item() for p in D.parameters()]),
        }
        # keep full per-sample gradients only where they are small; otherwise a strided slice
        for k, gk in enumerate(gs):
            flat = gk.reshape(B, -1)
            if flat.shape[1] <= 20000:
                rec[f"gs_{k}"] = flat
            else:
                idx = np.linspace(0, flat.shape[1] - 1, 4096).astype(np.int64)
                rec[f"gs_idx_{k}"] = idx
                rec[f"gs_{k}"] = flat[:, idx]

        # gradient_penalty.py: per-sample penalties on the same inputs + WGAN-GP with seeded alpha
        pen = RGP.calc_lipschitz_penalty_WRT(D, x.clone(), y, device="cpu", per_sample=True, one_sided=False, aux_penalty=True)
        pen1 = RGP.calc_lipschitz_penalty_WRT(D, x.clone(), y, device="cpu", per_sample=True, one_sided=True, aux_penalty=False)
        rec["lip_pen_two_sided_aux"] = pen.detach().numpy()
        rec["lip_pen_one_sided_noaux"] = pen1.detach().numpy()
        fake = torch.rand(x.shape, generator=g) * (2 if ishape[0] == 3 else 1) - (1 if ishape[0] == 3 else 0)
        torch.manual_seed(777)                     # gradient_penalty.py:33 draws alpha = torch.rand(B,1) on the CPU
        gp = RGP.calc_WGAN_GP_penalty(D, x.clone(), y, fake, y, device="cpu", per_sample=False, aux_penalty=True)
        rec["fake"] = fake.numpy()
        rec["wgan_gp_seed777"] = gp.detach().numpy()
        gp_grad = torch.autograd.grad(gp, list(D.parameters()), allow_unused=True)
        rec["wgan_gp_grad_norms"] = np.array([0.0 if t is None else t.double().norm().item() for t in gp_grad])
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **rec)
        print(name, "params", sum(p.numel() for p in D.parameters()), "norms", rec["micro_norms"][:, 0])

    # backprop_clip.l2_clip on fixed tensors, thresholds on both sides of the norms
    g = torch.Generator().manual_seed(99)
    rec = {}
    for j, shape in enumerate([(5, 784), (4, 8, 7, 7), (3, 1, 28, 28), (2, 128)]):
        t = torch.randn(shape, generator=g) * (0.1 + j)
        for C in (0.5, 5.0, 50.0):
            rec[f"in_{j}"] = t.numpy()
            rec[f"out_{j}_C{C}"] = RBC.l2_clip(t, C).numpy()
    np.savez_compressed(os.path.join(OUT, "l2_clip.npz"), **rec)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
