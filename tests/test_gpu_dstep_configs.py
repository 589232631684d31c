"""The reference's only test is test_configs.sh: {MNIST, CelebA} x {gc, is} x {unconditional, conditional}
with `-nms 1`, pass = "did not crash".  Same matrix here on synthetic data through the re-authored D step
(csl_gan_b200.dstep.DiscriminatorStep), plus the README invocations (adaptive per-layer clipping, per-parameter
IS with gradient penalty) and a check that the step actually moves the weights and spends privacy budget."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from csl_gan_b200 import discriminators as DD  # noqa: E402
from csl_gan_b200 import options as OPT  # noqa: E402
from csl_gan_b200.dstep import DiscriminatorStep, setup_privacy_engine  # noqa: E402

DEV = "cuda"


def _run(argv, B=16, steps=2, collect=False):
    opt = OPT.parse(argv + ["-bs", str(B), "-tss", "1000", "--manual_seed", "3"])
    n_classes = opt.n_classes if opt.conditional else 0
    D = DD.build_discriminator(opt.dataset, opt.model, n_classes=n_classes, im_size=opt.im_size,
                               emb_mode=opt.d_label_emb_mode, conditional_arch=opt.conditional_arch,
                               aux_loss_type=opt.aux_loss_type, aux_loss_scalar=opt.aux_loss_scalar,
                               weights_seed=opt.weights_seed, device=DEV)
    if opt.dataset == "CelebA":
        D = D.to(memory_format=torch.channels_last)
    n_params = len(list(D.parameters()))
    if opt.dp_mode == "gc" and opt.clipping_param_per_layer is not None and len(opt.clipping_param_per_layer) != n_params:
        opt.clipping_param_per_layer = [100.0] * n_params      # conditional critics have more tensors than the 9 defaults
    d_opt = torch.optim.Adam(D.parameters(), lr=opt.d_lr, betas=(opt.adam_b1, opt.adam_b2))
    eng = setup_privacy_engine(opt, D, d_opt)
    shape = (1, 28, 28) if opt.dataset == "MNIST" else (3, opt.im_size, opt.im_size)
    g = torch.Generator().manual_seed(0)

    def public_batch(n, labels):
        """stand-in for MeanSampler.sample / the public partition (out of scope): a noisy 'mean' batch"""
        x = torch.rand((n,) + shape, generator=g).to(DEV) * 0.5 + 0.25
        y = labels if labels is not None else (torch.randint(0, n_classes, (n,), generator=g).to(DEV) if n_classes else None)
        return x, y

    step = DiscriminatorStep(opt, D, d_opt, eng, public_batch=public_batch, collect_stats=collect)
    before = [p.detach().clone() for p in D.parameters()]
    out = None
    for _ in range(steps):
        real = torch.rand((B,) + shape, generator=g).to(DEV)
        fake = torch.rand((B,) + shape, generator=g).to(DEV)
        y = torch.randint(0, n_classes, (B,), generator=g).to(DEV) if n_classes else None
        out = step(real, y, fake, y, use_dp=True)
    torch.cuda.synchronize()
    assert all(torch.isfinite(p).all() for p in D.parameters())
    assert any((p.detach() - b).abs().max() > 0 for p, b in zip(D.parameters(), before))
    assert eng.steps == steps
    eps, alpha = eng.get_privacy_spent(opt.delta)
    assert eps > 0 and math.isfinite(eps)
    return opt, eng, out


@pytest.mark.parametrize("dataset", ["MNIST", "CelebA"])
@pytest.mark.parametrize("mode", ["gc", "is"])
@pytest.mark.parametrize("cond", [False, True])
def test_reference_smoke_matrix(dataset, mode, cond):
    """test_configs.sh:1-11"""
    argv = [dataset, "-dpm", mode, "-nms", "1", "--mean_sample_size", "10"] + (["--conditional"] if cond else [])
    opt, eng, out = _run(argv, B=8 if dataset == "CelebA" else 16)
    host = out.to_host()
    assert math.isfinite(host["D Adv Loss"])
    if dataset == "CelebA":
        assert "D Penalty" in host and math.isfinite(host["D Penalty"])       # WGAN-GP on the public batch (CELEBA_DEFAULTS)
    if mode == "is":
        s = eng.batch_sensitivity
        assert (s >= 0).all() if hasattr(s, "__len__") else s >= 0


def test_readme_invocations():
    """README.md:30-52: MNIST gc/is with sigma 10; CelebA gc adaptive-pl with -nms 32; CelebA is -ispp True."""
    _run(["MNIST", "--conditional", "--dp_mode", "gc", "--sigma", "10"], B=32, collect=True)
    _run(["MNIST", "--conditional", "--dp_mode", "is", "--sigma", "10"], B=32, collect=True)
    opt, eng, out = _run(["CelebA", "-nms", "32", "--dp_mode", "gc", "-gcm", "adaptive-pl"], B=8, collect=True)
    assert eng.is_per_layer and eng._thresholds_host is None                   # thresholds stayed on the device
    host = out.to_host()
    assert len(host["D Layer Grad Norm Means"]) == 9 and len(host["Grads Clipped"]) == 9
    opt, eng, out = _run(["CelebA", "-nms", "32", "--dp_mode", "is", "-ispp", "True"], B=8, collect=True)
    assert len(eng.batch_sensitivity) == 9
    # adaptive (flat) clipping and flat standard clipping on the conv critic
    _run(["CelebA", "-nms", "32", "--dp_mode", "gc", "-gcm", "adaptive"], B=8)
    _run(["MNIST", "--dp_mode", "gc", "-gcm", "constant-pl", "-cpl", "1.0"], B=16)
    # scaled flat immediate sensitivity with the moving-average update
    _run(["MNIST", "--dp_mode", "is", "-issm", "moving-avg-pl", "-issv", "1", "1", "1", "1"], B=16)


def test_cuda_graph_step_matches_eager_and_draws_fresh_noise():
    """GraphedDiscriminatorStep: same weights as the eager step after several noisy steps (same seed), which
    also proves every replay consumes a fresh part of the Philox stream."""
    import copy
    from csl_gan_b200.dstep import GraphedDiscriminatorStep
    B = 32
    results = {}
    for mode in ("eager", "graph"):
        opt = OPT.parse(["MNIST", "-dpm", "gc", "--conditional", "--sigma", "2", "-bs", str(B), "-tss", "1000", "--manual_seed", "5"])
        D = DD.build_discriminator("MNIST", "Vanilla", n_classes=10, conditional_arch="ACGAN", aux_loss_type="cross_entropy",
                                   weights_seed=42, device=DEV)
        d_opt = torch.optim.Adam(D.parameters(), lr=1e-3, capturable=True)
        eng = setup_privacy_engine(opt, D, d_opt)
        stepper = DiscriminatorStep(opt, D, d_opt, eng)
        g = torch.Generator().manual_seed(1)
        batches = [(torch.rand(B, 1, 28, 28, generator=g).to(DEV), torch.randint(0, 10, (B,), generator=g).to(DEV),
                    torch.rand(B, 1, 28, 28, generator=g).to(DEV)) for _ in range(7)]
        if mode == "graph":
            eng.enable_graph_safe_rng()
        # the first 3 steps run eagerly in both modes (PyTorch wants cuBLAS/cuDNN warmed up on a side stream
        # before a capture); the graph then replays steps 4..7
        if mode == "graph":
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for x, y, f in batches[:3]:
                    out = stepper(x, y, f, y, use_dp=True)
            torch.cuda.current_stream().wait_stream(side)
            x, y, f = batches[3]
            runner = GraphedDiscriminatorStep(stepper, (x, y, f, y), warmup=0)
            for x, y, f in batches[3:]:
                out = runner(x, y, f, y)
        else:
            for x, y, f in batches:
                out = stepper(x, y, f, y, use_dp=True)
        torch.cuda.synchronize()
        results[mode] = ([p.detach().clone() for p in D.parameters()], eng.philox_offset, eng.steps,
                         out.to_host()["D Adv Loss"])
    for a, b in zip(results["eager"][0], results["graph"][0]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)
    assert results["eager"][1] == results["graph"][1] > 0          # same amount of Philox stream consumed
    assert results["eager"][2] == results["graph"][2] == 7
    assert abs(results["eager"][3] - results["graph"][3]) < 1e-4
