"""Pin the oracle and the host-side critics against golden vectors generated from the
reference's own Python modules (oracle/gen_golden.py)."""
import os

import numpy as np
import pytest
import torch

from csl_gan_b200 import discriminators as DD
from oracle import dp_oracle as O

CASES = {
    "mnist_vanilla_acgan": lambda: DD.MNISTVanillaD(n_classes=10, conditional_arch="ACGAN", aux_loss_type="cross_entropy"),
    "mnist_vanilla_uncond": lambda: DD.MNISTVanillaD(n_classes=0, conditional_arch="ACGAN", aux_loss_type="cross_entropy"),
    "mnist_dcrn_acgan": lambda: DD.MNIST_DCRN_D(n_classes=10, conditional_arch="ACGAN", aux_loss_type="wasserstein"),
    "celeba_d64_uncond": lambda: DD.CelebA_DCRN_D64(n_classes=0, conditional_arch="ACGAN", aux_loss_type="wasserstein"),
    "celeba_d64_cgan": lambda: DD.CelebA_DCRN_D64(n_classes=2, conditional_arch="CGAN", aux_loss_type="wasserstein"),
    "celeba_d48_uncond": lambda: DD.CelebA_DCRN_D48(n_classes=0, conditional_arch="ACGAN", aux_loss_type="wasserstein"),
}


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def _build(name):
    torch.manual_seed(42)
    return CASES[name]()


@pytest.mark.parametrize("name", list(CASES))
def test_critic_matches_reference_forward_and_losses(golden_dir, name):
    g = _load(golden_dir, name)
    D = _build(name)
    assert [n for n, _ in D.named_parameters()] == list(g["param_names"])
    np.testing.assert_allclose([p.detach().double().sum().item() for p in D.parameters()], g["param_sums"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose([p.detach().double().abs().sum().item() for p in D.parameters()], g["param_abs_sums"], rtol=1e-9)
    x = torch.from_numpy(g["x"])
    y = torch.from_numpy(g["y"]) if g["y"].size else None
    out, aux = D(x, y)
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=1e-5, atol=1e-6)
    if g["aux"].size:
        np.testing.assert_allclose(aux.detach().numpy(), g["aux"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(D.aux_loss(aux, y).detach().numpy(), g["aux_loss"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(D.real_loss(out).detach().numpy(), g["real_loss"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(D.fake_loss(out).detach().numpy(), g["fake_loss"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", list(CASES))
def test_hook_grad_samples_match_reference_microbatch(golden_dir, name):
    """The oracle's hook contraction (B * backprop x activation) must reproduce the
    micro-batch per-sample gradients computed on the reference's own critic."""
    g = _load(golden_dir, name)
    D = _build(name)
    x = torch.from_numpy(g["x"])
    y = torch.from_numpy(g["y"]) if g["y"].size else None
    B = x.shape[0]
    eng = O.OracleGCEngine(D, batch_size=B, noise_multiplier=0.0, max_grad_norm=1.0)
    out, aux = D(x, y)
    loss = D.real_loss(out)
    if bool(g["micro_uses_aux"]):
        loss = loss + D.aux_loss(aux, y)
    loss.backward()
    gs = eng.grad_samples()
    norms = torch.stack(O.calc_sample_norms(gs, flat=False))[:, 0].numpy()
    live = g["micro_norms"] > 0
    np.testing.assert_allclose(norms[live], g["micro_norms"][live], rtol=2e-4, atol=1e-7)
    for k, gk in enumerate(gs):
        if not live[k].any():
            continue
        flat = gk[0].reshape(B, -1).numpy()
        if f"gs_idx_{k}" in g.files:
            flat = flat[:, g[f"gs_idx_{k}"]]
        ref = g[f"gs_{k}"]
        scale = np.abs(ref).max() + 1e-12
        assert np.abs(flat - ref).max() / scale < 2e-5, (name, k)
    eng.remove()


@pytest.mark.parametrize("name", list(CASES))
def test_gradient_penalty_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    D = _build(name)
    x = torch.from_numpy(g["x"])
    y = torch.from_numpy(g["y"]) if g["y"].size else None
    pen = O.lipschitz_penalty(D, x, y, per_sample=True, one_sided=False, aux_penalty=True)
    np.testing.assert_allclose(pen.detach().numpy(), g["lip_pen_two_sided_aux"], rtol=2e-4, atol=1e-6)
    pen1 = O.lipschitz_penalty(D, x, y, per_sample=True, one_sided=True, aux_penalty=False)
    np.testing.assert_allclose(pen1.detach().numpy(), g["lip_pen_one_sided_noaux"], rtol=2e-4, atol=1e-6)
    torch.manual_seed(777)
    alpha = torch.rand(x.shape[0], 1)
    gp = O.wgan_gp_penalty(D, x, y, torch.from_numpy(g["fake"]), alpha, aux_penalty=True)
    np.testing.assert_allclose(gp.detach().numpy(), g["wgan_gp_seed777"], rtol=2e-4, atol=1e-6)
    grads = torch.autograd.grad(gp, list(D.parameters()), allow_unused=True)
    got = np.array([0.0 if t is None else t.double().norm().item() for t in grads])
    np.testing.assert_allclose(got, g["wgan_gp_grad_norms"], rtol=5e-4, atol=1e-7)


def test_l2_clip_matches_reference(golden_dir):
    g = _load(golden_dir, "l2_clip")
    for j in range(4):
        t = torch.from_numpy(g[f"in_{j}"])
        for C in (0.5, 5.0, 50.0):
            np.testing.assert_allclose(O.l2_clip(t, C).numpy(), g[f"out_{j}_C{C}"], rtol=1e-6, atol=1e-7)
