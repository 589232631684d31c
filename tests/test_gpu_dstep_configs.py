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


# --------------------------------------------------------------------------------------------------------------
# DiscriminatorStep against the oracle, end to end, the per-sample penalty on private data
# (train.py:434-450) and the engine state a checkpoint needs (8 f-3)
# --------------------------------------------------------------------------------------------------------------
def _rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / (b.norm() + 1e-5 * max(1.0, b.numel() ** 0.5))).item()


def _oracle_step_grads(D, opt, real, labels, fake, B, extra_summed=None):
    """The reference's train_D arithmetic (train.py:382-431, 484) on the CPU oracle, sigma = 0:
    fake pass, real pass, d_loss.backward() with hooks, clip, accum, accumulate_batch, [+ B * grad(penalty)], / B."""
    from oracle import dp_oracle as O
    per_layer = opt.grad_clip_mode[-3:] == "-pl"
    eng = O.OracleGCEngine(D, batch_size=B, noise_multiplier=0.0,
                           max_grad_norm=opt.clipping_param_per_layer if per_layer else opt.clipping_param,
                           accum_passes=not opt.grad_clip_split, num_private_passes=1 if opt.grad_clip_split else None)
    of, af = D(fake, labels)
    orr, ar = D(real, labels)
    loss = D.real_loss(orr) + D.fake_loss(of)
    if ar is not None:
        loss = loss + D.aux_loss(ar, labels) + D.aux_loss(af, labels, fake=True)
    loss.backward()
    eng.disable_hooks()
    eng.clip()
    if opt.grad_clip_split:
        eng.accum_grads_across_passes()
    eng.accumulate_batch()
    if extra_summed is not None:
        for p, g in zip(eng.params(), extra_summed(D)):
            if g is not None:
                p.summed_grad = p.summed_grad + g * B
    eng.step_grads(None)
    out = [p.grad.clone() for p in eng.params()]
    eng.remove()
    return out


@pytest.mark.parametrize("argv,B", [
    (["MNIST", "-dpm", "gc", "--conditional", "--sigma", "0"], 48),
    (["MNIST", "-dpm", "gc", "--sigma", "0", "-gcs", "False", "-c", "1.0"], 33),
    (["CelebA", "-dpm", "gc", "-gcm", "constant-pl", "--sigma", "0", "-nms", "8", "-cpl", "0.6", "0.2", "1.5", "0.2", "2.0",
      "0.1", "2.0", "0.05", "1.0"], 8),
])
def test_discriminator_step_matches_oracle_end_to_end(argv, B):
    """One whole DiscriminatorStep (critic forward/backward through cuDNN, capture hooks, norms, clip, accumulate,
    WGAN-GP on a public batch where configured, noise-free step) against the same sequence on the CPU oracle."""
    import copy
    from oracle import dp_oracle as O
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        opt = OPT.parse(argv + ["-bs", str(B), "-tss", "1000", "--manual_seed", "9"])
        ncls = opt.n_classes if opt.conditional else 0
        D = DD.build_discriminator(opt.dataset, opt.model, n_classes=ncls, im_size=opt.im_size, emb_mode=opt.d_label_emb_mode,
                                   conditional_arch=opt.conditional_arch, aux_loss_type=opt.aux_loss_type,
                                   aux_loss_scalar=opt.aux_loss_scalar, weights_seed=opt.weights_seed, device="cpu")
        shape = (1, 28, 28) if opt.dataset == "MNIST" else (3, opt.im_size, opt.im_size)
        g = torch.Generator().manual_seed(B)
        real = torch.rand((B,) + shape, generator=g)
        fake = torch.rand((B,) + shape, generator=g) * 0.5 + 0.1 * torch.randn((B,) + shape, generator=g)
        labels = torch.randint(0, ncls, (B,), generator=g) if ncls else None
        pub = torch.rand((B,) + shape, generator=g) * 0.7
        alpha = torch.rand(B, 1, generator=g)
        Dg = copy.deepcopy(D).to(DEV)
        d_opt = torch.optim.SGD(Dg.parameters(), lr=0.0)
        eng = setup_privacy_engine(opt, Dg, d_opt)
        stepper = DiscriminatorStep(opt, Dg, d_opt, eng,
                                    public_batch=lambda n, lab: (pub[:n].to(DEV), lab))
        extra = None
        if opt.penalty:
            import csl_gan_b200.dstep as DS
            orig = DS.calc_penalty
            DS.calc_penalty = lambda *a, **k: orig(*a, **k, alpha=alpha.to(DEV))      # the same interpolation factors
            extra = lambda Dm: torch.autograd.grad(
                O.wgan_gp_penalty(Dm, pub, labels, fake, alpha, weight=10.0, aux_penalty=opt.aux_penalty),
                list(Dm.parameters()), allow_unused=True)
        try:
            res = stepper(real.to(DEV), None if labels is None else labels.to(DEV), fake.to(DEV),
                          None if labels is None else labels.to(DEV), use_dp=True)
        finally:
            if opt.penalty:
                DS.calc_penalty = orig
        ref = _oracle_step_grads(D, opt, real, labels, fake, B, extra)
        for k, (p, r) in enumerate(zip(Dg.parameters(), ref)):
            assert _rel(p.grad, r) < 2e-3, (k, _rel(p.grad, r))      # 1e-3 kernels + cuDNN-vs-CPU critic backward
        assert eng.steps == 1 and torch.isfinite(res.d_real_loss)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_per_sample_penalty_on_private_data_matches_oracle():
    """-pupd False (train.py:434-450): per-sample penalty gradients are added into p.grad_sample[0, i], then the batch
    is clipped again.  The only consumer of a writable grad_sample."""
    import copy
    from oracle import dp_oracle as O
    B = 6
    opt = OPT.parse(["MNIST", "-dpm", "gc", "--sigma", "0", "--penalty", "WGAN-GP", "-pupd", "False", "-c", "2.0",
                     "-bs", str(B), "-tss", "1000", "--manual_seed", "3"])
    D = DD.build_discriminator("MNIST", "Vanilla", n_classes=0, weights_seed=42, device="cpu")
    g = torch.Generator().manual_seed(1)
    real = torch.rand(B, 1, 28, 28, generator=g)
    fake = torch.rand(B, 1, 28, 28, generator=g) * 0.5
    alpha = torch.rand(B, 1, generator=g)
    Dg = copy.deepcopy(D).to(DEV)
    d_opt = torch.optim.SGD(Dg.parameters(), lr=0.0)
    eng = setup_privacy_engine(opt, Dg, d_opt)
    import csl_gan_b200.dstep as DS
    orig = DS.calc_penalty
    DS.calc_penalty = lambda *a, **k: orig(*a, **k, alpha=alpha.to(DEV))
    try:
        DiscriminatorStep(opt, Dg, d_opt, eng)(real.to(DEV), None, fake.to(DEV), None, use_dp=True)
    finally:
        DS.calc_penalty = orig
    # oracle: grad_sample[0, i] += grad(penalty_i); norms, factors and the weighted sum from the modified tensors
    oe = O.OracleGCEngine(D, batch_size=B, noise_multiplier=0.0, max_grad_norm=2.0, num_private_passes=1)
    of, _ = D(fake, None)
    orr, _ = D(real, None)
    (D.real_loss(orr) + D.fake_loss(of)).backward()
    oe.disable_hooks()
    gs = oe.grad_samples()
    pens = O.wgan_gp_penalty(D, real, None, fake, alpha, weight=10.0, per_sample=True, aux_penalty=opt.aux_penalty)
    params = list(D.parameters())
    for i in range(B):
        pg = torch.autograd.grad(pens[i], params, retain_graph=True, allow_unused=True)
        for k, gk in enumerate(pg):
            if gk is not None:
                gs[k][0, i] += gk
    norms = O.calc_sample_norms(gs, flat=True)
    fac = O.calc_clipping_factors(norms, 2.0, len(params))
    ref = [O.weighted_sum(f, t).sum(dim=0) / B for f, t in zip(fac, gs)]
    oe.remove()
    assert (fac[0] < 0.999).any()
    for p, r in zip(Dg.parameters(), ref):
        assert _rel(p.grad, r) < 2e-3


def test_engine_state_dict_round_trip_resumes_the_noise_stream():
    """the reference forgets the engine on resume (util.py:16-42 saves only models and
    optimizers, so its accountant restarts at 0).  state_dict() carries steps, the Philox (seed, offset) -- also when
    the offset lives on the device for CUDA-graph replay -- and the thresholds: a resumed engine continues the exact
    noise stream and the privacy accounting."""
    import copy
    import io
    from csl_gan_b200.dstep import GraphedDiscriminatorStep
    B = 16
    g = torch.Generator().manual_seed(2)
    batches = [(torch.rand(B, 1, 28, 28, generator=g).to(DEV), torch.rand(B, 1, 28, 28, generator=g).to(DEV)) for _ in range(6)]

    def build(graph_rng):
        opt = OPT.parse(["MNIST", "-dpm", "gc", "--sigma", "1.5", "-bs", str(B), "-tss", "1000", "--manual_seed", "77"])
        D = DD.build_discriminator("MNIST", "Vanilla", n_classes=0, weights_seed=42, device=DEV)
        d_opt = torch.optim.Adam(D.parameters(), lr=1e-3, capturable=True)
        eng = setup_privacy_engine(opt, D, d_opt)
        if graph_rng:
            eng.enable_graph_safe_rng()
        return opt, D, d_opt, eng, DiscriminatorStep(opt, D, d_opt, eng)

    for graph_rng in (False, True):
        # uninterrupted run: 6 steps
        _, Da, _, ea, sa = build(graph_rng)
        for x, f in batches:
            sa(x, None, f, None, use_dp=True)
        # interrupted after 3 steps: checkpoint (through torch.save / torch.load), fresh objects, 3 more steps
        _, Db, ob, eb, sb = build(graph_rng)
        for x, f in batches[:3]:
            sb(x, None, f, None, use_dp=True)
        buf = io.BytesIO()
        torch.save({"D": Db.state_dict(), "opt": ob.state_dict(), "engine": eb.state_dict()}, buf)
        buf.seek(0)
        ck = torch.load(buf, weights_only=False)
        _, Dc, oc, ec, sc = build(graph_rng)
        Dc.load_state_dict(ck["D"]); oc.load_state_dict(ck["opt"]); ec.load_state_dict(ck["engine"])
        assert ec.steps == 3 and ec.philox_offset == eb.philox_offset > 0
        for x, f in batches[3:]:
            sc(x, None, f, None, use_dp=True)
        torch.cuda.synchronize()
        assert ec.steps == ea.steps == 6 and ec.philox_offset == ea.philox_offset
        assert ec.get_privacy_spent(1e-5) == ea.get_privacy_spent(1e-5)
        for a, c in zip(Da.parameters(), Dc.parameters()):
            assert torch.equal(a, c)
