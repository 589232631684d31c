"""Engine-level parity on the GPU: csl_gan_b200.PrivacyEngine / ISPrivacyEngine (CUDA kernels through the
C ABI) against the CPU oracle on identical seeded inputs, plus golden vectors from the reference."""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import csl_gan_b200 as cg  # noqa: E402
from csl_gan_b200 import discriminators as DD  # noqa: E402
from oracle import dp_oracle as O  # noqa: E402

DEV = "cuda"
REL_TOL = 1e-3      # north_star: clipped summed gradients within 1e-3 relative tolerance (sigma = 0)


@pytest.fixture(autouse=True)
def _fp32_library_math():
    """cuDNN / cuBLAS default to TF32 for the critic's own forward/backward on the GPU; the parity
    target is the fp32 CPU oracle, so the library side runs in fp32 here (our kernels' TF32 operand
    rounding is what the 1e-3 tolerance covers)."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    """normwise relative error with a small absolute floor (a reference that is exactly 0, e.g. a
    bias whose clipped contributions cancel, must not turn rounding noise into an infinite ratio)."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / (b.norm() + 1e-5 * max(1.0, b.numel() ** 0.5))).item()


def make(name):
    torch.manual_seed(42)
    if name == "mnist":
        D = DD.MNISTVanillaD(n_classes=10, conditional_arch="ACGAN", aux_loss_type="cross_entropy")
        shape, ncls, lo = (1, 28, 28), 10, 0.0
    elif name == "mnist_uncond":
        D = DD.MNISTVanillaD(n_classes=0)
        shape, ncls, lo = (1, 28, 28), 0, 0.0
    elif name == "mnist_dcrn":
        D = DD.MNIST_DCRN_D(n_classes=0)
        shape, ncls, lo = (1, 28, 28), 0, 0.0
    elif name == "d64":
        D = DD.CelebA_DCRN_D64(n_classes=0)
        shape, ncls, lo = (3, 64, 64), 0, -1.0
    elif name == "d48":
        D = DD.CelebA_DCRN_D48(n_classes=0)
        shape, ncls, lo = (3, 48, 48), 0, -1.0
    else:
        raise KeyError(name)
    return D, shape, ncls, lo


def batch(shape, ncls, lo, B, seed):
    g = torch.Generator().manual_seed(seed)
    real = torch.rand((B,) + shape, generator=g) * (1 - lo) + lo
    # a fake batch from a visibly different distribution so fake/real gradients do not cancel
    fake = (torch.rand((B,) + shape, generator=g) * (1 - lo) + lo) * 0.5 + 0.25 * torch.randn((B,) + shape, generator=g)
    y = torch.randint(0, ncls, (B,), generator=g) if ncls > 1 else None
    return real, fake, y


def d_loss(D, real, fake, y):
    """train.py:382-384: fake pass first, then real pass; loss = real + fake (+ aux terms)."""
    of, af = D(fake, y)
    orr, ar = D(real, y)
    loss = D.real_loss(orr) + D.fake_loss(of)
    if ar is not None:
        loss = loss + D.aux_loss(ar, y) + D.aux_loss(af, y, fake=True)
    return loss


def run_oracle(D, real, fake, y, B, C, sigma=0.0):
    eng = O.OracleGCEngine(D, batch_size=B, noise_multiplier=sigma, max_grad_norm=C,
                           accum_passes=False, num_private_passes=1)
    eng.disable_hooks()
    eng.enable_hooks()
    d_loss(D, real, fake, y).backward()
    eng.disable_hooks()
    captured = [dict(eng.captured[k]) for k in sorted(eng.captured)]
    norms = eng.sample_norms()
    factors = eng.clipping_factors(norms)
    per_param_norms = O.calc_sample_norms(eng.grad_samples(), flat=False)
    clipped = [c.clone() for c in eng.clip()]
    eng.accum_grads_across_passes()
    eng.accumulate_batch()
    summed = [p.summed_grad.clone() for p in eng.params()]
    eng.step_grads(None)
    grads = [p.grad.clone() for p in eng.params()]
    eng.remove()
    return dict(norms=norms, factors=factors, per_param_norms=per_param_norms, clipped=clipped, summed=summed, grads=grads,
                captured=captured)


def run_cuda(Dg, real, fake, y, B, C, sigma=0.0, seed=7, captures=None, **engine_kw):
    opt = torch.optim.Adam(Dg.parameters(), lr=0.0)
    eng = cg.PrivacyEngine(Dg, batch_size=B, sample_size=60000, noise_multiplier=sigma, max_grad_norm=C,
                           accum_passes=False, num_private_passes=1, auto_clip_and_accum_on_step=False, **engine_kw)
    eng.disable_hooks()
    eng.attach(opt)
    eng._set_seed(seed)
    for p in Dg.parameters():
        p.grad = None
    eng.enable_hooks()
    yg = None if y is None else y.to(DEV)
    if captures is None:
        d_loss(Dg, real.to(DEV), fake.to(DEV), yg).backward()
    else:
        # bit-identical inputs to the oracle's: isolates the DP kernels from cuDNN-vs-CPU differences of
        # the critic's own backward (cuDNN's 5x5 dgrad algorithms differ from the CPU by ~1e-4, which an
        # ill-conditioned quantity such as a bias-gradient norm amplifies beyond 1e-3)
        eng.ingest_captures([{n: (a.to(DEV), g.to(DEV)) for n, (a, g) in layers.items()} for layers in captures])
    eng.expose_grad_sample_attrs()
    assert next(iter(Dg.parameters())).grad_sample.size(1) == B          # train.py:388
    eng.disable_hooks()
    all_norms = cg.calc_sample_norms(named_params=eng.clipper._named_grad_samples(),
                                     flat=not eng.clipper.norm_clipper.is_per_layer)          # train.py:311-314
    per_param = eng.per_sample_norms().clone()
    it = iter(eng.clipper.norm_clipper.calc_clipping_factors(all_norms))                      # train.py:324
    factors = [next(it).clone() for _ in range(len(all_norms))]
    eng.clip()
    eng.accum_grads_across_passes()
    eng.accumulate_batch()
    summed = [p.summed_grad.clone() for p in Dg.parameters()]
    opt.step()
    grads = [p.grad.clone() for p in Dg.parameters()]
    torch.cuda.synchronize()
    return dict(norms=[n.clone() for n in all_norms], factors=factors, per_param_norms=per_param, summed=summed,
                grads=grads, engine=eng)


def run_oracle_chunked(D, real, fake, y, B, C, chunk=64):
    """The oracle over a large batch in `chunk`-sample pieces: per-sample gradients, norms and clip factors of a
    batch-mean loss do not depend on the other samples (B * d(mean loss)/d(out_i) = d(loss_i)/d(out_i)), so the
    clipped sums of the chunks add up to the full-batch clipped sum while only chunk x |theta| floats are ever
    materialised.  Captured grad_outputs are rescaled from the chunk's mean to the full batch's mean."""
    summed, clipped, norms, pp, factors = None, None, [], [], []
    caps = None
    for lo_ in range(0, B, chunk):
        hi_ = min(B, lo_ + chunk)
        n = hi_ - lo_
        r = run_oracle(copy.deepcopy(D), real[lo_:hi_], fake[lo_:hi_], None if y is None else y[lo_:hi_], n, C)
        summed = r["summed"] if summed is None else [a + b for a, b in zip(summed, r["summed"])]
        clipped = r["clipped"] if clipped is None else [a + b for a, b in zip(clipped, r["clipped"])]
        norms.append(r["norms"]); pp.append(r["per_param_norms"]); factors.append(r["factors"])
        if caps is None:
            caps = [{k: ([], []) for k in layers} for layers in r["captured"]]
        for ps, layers in enumerate(r["captured"]):
            for k, (a, g) in layers.items():
                caps[ps][k][0].append(a)
                caps[ps][k][1].append(g * (n / B))
    cat = lambda parts: [torch.cat([p[i] for p in parts], dim=1) for i in range(len(parts[0]))]
    captured = [{k: (torch.cat(a), torch.cat(g)) for k, (a, g) in layers.items()} for layers in caps]
    return dict(norms=cat(norms), per_param_norms=cat(pp), factors=cat(factors), summed=summed, clipped=clipped,
                grads=[t / B for t in summed], captured=captured)


CASES = [
    # model, B, C (float -> flat, list -> per layer), note
    ("mnist", 600, 4.0),                   # BASELINE config 1: everything clips at random init
    ("mnist", 64, "median"),               # both branches of min(1, C/n)
    ("mnist", 37, [3.0, 0.2, 0.5, 0.2, 1.0, 0.5]),   # ragged batch, per-layer
    ("mnist_uncond", 50, 2.0),
    ("mnist_dcrn", 9, "median"),           # conv, Q=196 / 49 (ragged k tails), odd batch
    ("mnist_dcrn", 32, "median"),          # batch == Bpad: both passes normed in one launch
    ("d64", 6, [1000, 200, 1000, 100, 1000, 100, 1000, 5, 2500]),   # reference CelebA per-layer defaults: nothing clips
    ("d64", 6, "median"),
    ("d64", 5, "median-pl"),
    ("d48", 3, "median"),
    ("d64", 32, "median-pl"),              # batch == Bpad on the channels-last + ghost path
]


@pytest.mark.parametrize("name,B,C", CASES)
def test_gc_step_matches_oracle(name, B, C):
    D, shape, ncls, lo = make(name)
    real, fake, y = batch(shape, ncls, lo, B, seed=B)
    Dg = copy.deepcopy(D).to(DEV)
    if isinstance(C, str):
        probe = run_oracle(copy.deepcopy(D), real, fake, y, B, 1e9)
        if C == "median":
            C = float(probe["norms"][0].median())
        else:
            C = [float(n.median()) for n in probe["per_param_norms"]]
    ref = run_oracle(D, real, fake, y, B, C)
    got = run_cuda(Dg, real, fake, y, B, C, captures=ref["captured"] if B >= 32 and name != "mnist" else None)
    # norms [n_passes, B] (flat: one item; per-layer: one per parameter)
    assert len(got["norms"]) == len(ref["norms"])
    for a, b in zip(got["norms"], ref["norms"]):
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), rtol=1e-3, atol=1e-7)
    for k, b in enumerate(ref["per_param_norms"]):
        np.testing.assert_allclose(got["per_param_norms"][k].cpu().numpy(), b.numpy(), rtol=1e-3, atol=1e-7)
    for a, b in zip(got["factors"], ref["factors"]):
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), rtol=1e-3, atol=1e-7)
    if not isinstance(C, list) and name != "d64":
        f = ref["factors"][0]
        assert (f < 0.999).any(), "test must exercise the clipped branch"
    for k, (a, b) in enumerate(zip(got["summed"], ref["summed"])):
        # the fake and real sums can cancel exactly (e.g. a bias whose every sample is clipped to +-C_k):
        # allow 1e-5 of the per-pass magnitudes on top of the 1e-3 relative tolerance
        floor = 1e-5 * sum(c.double().norm().item() for c in ref["clipped"][k])
        err = (a.cpu().double() - b.double()).norm().item()
        assert err <= REL_TOL * b.double().norm().item() + floor, (k, err, b.norm().item())
        assert rel(got["grads"][k], ref["grads"][k]) < REL_TOL or err / B <= floor / B


CELEBA_CPL = [1000, 200, 1000, 100, 1000, 100, 1000, 5, 2500]           # reference options.py:80


@pytest.mark.parametrize("B,C,n_clip", [
    (128, CELEBA_CPL, 0),          # BASELINE configs 3/4: batch 128, the reference's per-layer defaults (nothing clips)
    (128, "median-pl", 1),         # both branches of min(1, C/n) at the configured batch size
    (512, "median-pl", 1),         # what bench.py times: 512 per GPU x 2 passes = 1024 slots, split-K over thousands of
                                   # k-blocks, several waves of CTA pairs
    (512, CELEBA_CPL, 0),
])
def test_gc_celeba_at_benchmark_batch_sizes_matches_chunked_oracle(B, C, n_clip):
    """the CelebA configurations that are timed are compared with the oracle at the size they are
    timed at (the oracle runs in 64-sample chunks; per-sample work is independent, the clipped sums add)."""
    D, shape, ncls, lo = make("d64")
    real, fake, y = batch(shape, ncls, lo, B, seed=1000 + B)
    if isinstance(C, str):
        probe = run_oracle_chunked(D, real, fake, y, B, 1e9)
        C = [float(n.median()) for n in probe["per_param_norms"]]
    ref = run_oracle_chunked(D, real, fake, y, B, C)
    Dg = copy.deepcopy(D).to(DEV).to(memory_format=torch.channels_last)      # the layout bench.py runs the critic in
    got = run_cuda(Dg, real, fake, y, B, C, captures=ref["captured"])
    for k, b in enumerate(ref["per_param_norms"]):
        np.testing.assert_allclose(got["per_param_norms"][k].cpu().numpy(), b.numpy(), rtol=1e-3, atol=1e-7)
    for a, b in zip(got["factors"], ref["factors"]):
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), rtol=1e-3, atol=1e-7)
        assert bool((b < 0.999).any()) == bool(n_clip)
    for k, (a, b) in enumerate(zip(got["summed"], ref["summed"])):
        floor = 1e-5 * ref["clipped"][k].double().abs().sum(dim=0).norm().item()
        err = (a.cpu().double() - b.double()).norm().item()
        assert err <= REL_TOL * b.double().norm().item() + floor, (k, err, b.norm().item())
        assert rel(got["grads"][k], ref["grads"][k]) < REL_TOL or err <= floor


@pytest.mark.parametrize("name,B,C", [("mnist", 40, "median"), ("mnist_dcrn", 7, "median"), ("d64", 4, "median-pl"),
                                      ("mnist", 64, [3.0, 0.2, 0.5, 0.2, 1.0, 0.5])])
def test_gc_joint_clipping_accum_passes(name, B, C):
    """U2 (-gcs False): accum_passes=True sums fake_i + real_i per sample and clips the sum once."""
    D, shape, ncls, lo = make(name)
    real, fake, y = batch(shape, ncls, lo, B, seed=B + 100)

    def oracle(Cv):
        Dm = copy.deepcopy(D)
        eng = O.OracleGCEngine(Dm, batch_size=B, noise_multiplier=0.0, max_grad_norm=Cv, accum_passes=True,
                               num_private_passes=None)
        d_loss(Dm, real, fake, y).backward()
        captured = [dict(eng.captured[k]) for k in sorted(eng.captured)]
        norms = eng.sample_norms()
        pp = O.calc_sample_norms(eng.grad_samples(), flat=False)
        eng.clip(); eng.accumulate_batch(); eng.step_grads(None)
        return norms, pp, [p.grad.clone() for p in eng.params()], captured

    if isinstance(C, str):
        norms, pp, _, _ = oracle(1e9)
        C = float(norms[0].median()) if C == "median" else [float(n.median()) for n in pp]
    norms, pp, grads, captured = oracle(C)
    assert norms[0].shape[0] == 1                                   # one joint "pass"
    Dg = copy.deepcopy(D).to(DEV)
    opt = torch.optim.SGD(Dg.parameters(), lr=0.0)
    eng = cg.PrivacyEngine(Dg, batch_size=B, sample_size=1000, noise_multiplier=0.0, max_grad_norm=C,
                           accum_passes=True, num_private_passes=None, auto_clip_and_accum_on_step=False)
    eng.attach(opt)
    eng.ingest_captures([{n: (a.to(DEV), g.to(DEV)) for n, (a, g) in layers.items()} for layers in captured])
    got_pp = eng.per_sample_norms()
    assert tuple(got_pp.shape) == (len(pp), 1, B)
    for k, b in enumerate(pp):
        np.testing.assert_allclose(got_pp[k].cpu().numpy(), b.numpy(), rtol=1e-3, atol=1e-7)
    eng.clip(); eng.accumulate_batch()
    opt.step()
    for p, g in zip(Dg.parameters(), grads):
        assert rel(p.grad, g) < REL_TOL


@pytest.mark.parametrize("name,sizes", [("mnist_dcrn", (40, 13, 40, 1)), ("d48", (12, 5, 9))])
def test_smaller_last_batch_and_engine_reuse_across_steps(name, sizes):
    """A batch smaller than the configured batch_size (last batch of an epoch) and several consecutive steps on
    one engine: buffers are reused, stale slots of the previous (larger) step must not leak in."""
    D, shape, ncls, lo = make(name)
    Dg = copy.deepcopy(D).to(DEV)
    opt = torch.optim.SGD(Dg.parameters(), lr=0.0)
    eng = cg.PrivacyEngine(Dg, batch_size=sizes[0], sample_size=1000, noise_multiplier=0.0, max_grad_norm=0.4,
                           num_private_passes=1, auto_clip_and_accum_on_step=False)
    eng.attach(opt)
    for B in sizes:
        real, fake, y = batch(shape, ncls, lo, B, seed=50 + B)
        ref = run_oracle(copy.deepcopy(D), real, fake, y, B, 0.4)
        eng.enable_hooks()
        eng.ingest_captures([{n: (a.to(DEV), g.to(DEV)) for n, (a, g) in layers.items()} for layers in ref["captured"]])
        eng.disable_hooks()
        eng.clip(); eng.accum_grads_across_passes(); eng.accumulate_batch()
        opt.step()
        for p, g in zip(Dg.parameters(), ref["grads"]):
            assert rel(p.grad, g) < REL_TOL, B
    assert eng.steps == len(sizes)


@pytest.mark.parametrize("name", ["mnist", "d48"])
def test_frozen_weights_backward_matches_plain_backward(name):
    """enable_hooks(freeze_weights=True) + engine.backward(loss) (dgrad chain only, no weight gradients) stages
    exactly the same per-sample operands as the reference-style loss.backward(); requires_grad is restored."""
    D, shape, ncls, lo = make(name)
    B = 6
    real, fake, y = batch(shape, ncls, lo, B, seed=21)
    real, fake, y = real.to(DEV), fake.to(DEV), (y.to(DEV) if y is not None else None)
    outs = {}
    for mode in ("plain", "inputs", "frozen"):
        Dg = copy.deepcopy(D).to(DEV)
        eng = cg.PrivacyEngine(Dg, batch_size=B, sample_size=1000, noise_multiplier=0.0, max_grad_norm=0.3,
                               num_private_passes=1, auto_clip_and_accum_on_step=False)
        eng.enable_hooks(freeze_weights=(mode == "frozen"))
        if mode == "frozen":
            assert not any(p.requires_grad for p in Dg.parameters())
        loss = d_loss(Dg, real, fake, y)
        loss.backward() if mode == "plain" else eng.backward(loss)
        eng.disable_hooks()
        assert all(p.requires_grad for p in Dg.parameters())
        if mode != "plain":
            assert all(p.grad is None for p in Dg.parameters())          # no batch-summed weight gradients formed
        norms = eng.per_sample_norms().clone()
        outs[mode] = (norms, [c.clone() for c in eng.clip()])
        eng.accumulate_batch()
    for mode in ("inputs", "frozen"):
        assert torch.allclose(outs[mode][0], outs["plain"][0], rtol=1e-5, atol=1e-8)
        for a, b in zip(outs[mode][1], outs["plain"][1]):
            assert rel(a, b) < 1e-5


def test_lazy_grad_sample_view_and_materialize():
    D, shape, ncls, lo = make("mnist_dcrn")
    B = 5
    real, fake, y = batch(shape, ncls, lo, B, seed=3)
    ref_eng = O.OracleGCEngine(D, batch_size=B, noise_multiplier=0.0, max_grad_norm=1.0)
    d_loss(D, real, fake, y).backward()
    ref_gs = ref_eng.grad_samples()
    Dg = copy.deepcopy(D).to(DEV)
    for p in Dg.parameters():
        p.grad = None
    eng = cg.PrivacyEngine(Dg, batch_size=B, sample_size=1000, noise_multiplier=0.0, max_grad_norm=1.0,
                           auto_clip_and_accum_on_step=False)
    d_loss(Dg, real.to(DEV), fake.to(DEV), None).backward()
    eng.expose_grad_sample_attrs()
    for k, p in enumerate(Dg.parameters()):
        gn = p.grad_sample[0].view(B, -1).norm(2, dim=1)                 # train.py:233 access pattern
        np.testing.assert_allclose(gn.cpu().numpy(), ref_gs[k][0].reshape(B, -1).norm(2, dim=1).numpy(), rtol=1e-3, atol=1e-7)
        full = p.grad_sample.materialize()
        assert tuple(full.shape) == tuple(ref_gs[k].shape)
        assert rel(full, ref_gs[k]) < REL_TOL
    th = eng.adaptive_thresholds("mean", 1.5, pass_idx=0)                 # train.py:230-241 without host syncs
    ref_th = torch.stack([g[0].reshape(B, -1).norm(2, dim=1).mean() * 1.5 for g in ref_gs])
    np.testing.assert_allclose(th.cpu().numpy(), ref_th.numpy(), rtol=1e-3)
    eng.set_max_grad_norm(th)                                             # device thresholds, per layer
    assert eng.is_per_layer
    eng.clip()
    eng.accumulate_batch()
    ref_eng.remove()


def test_noise_step_bit_exact_vs_torch_generator():
    """sigma > 0: the noise the engine adds equals what the upstream op sequence draws from a CUDA
    torch.Generator seeded the same way (one torch.normal per parameter tensor, in order)."""
    D, shape, ncls, lo = make("mnist")
    B, C, sigma, seed = 64, 4.0, 10.0, 4242
    real, fake, y = batch(shape, ncls, lo, B, seed=1)
    Dg = copy.deepcopy(D).to(DEV)
    got = run_cuda(Dg, real, fake, y, B, C, sigma=sigma, seed=seed)
    gen = torch.Generator(device=DEV)
    gen.manual_seed(seed)
    for s, g in zip(got["summed"], got["grads"]):
        noise = torch.normal(0.0, sigma * C, s.shape, device=DEV, generator=gen)
        ref = s / B
        noise /= B
        ref += noise
        assert torch.equal(g, ref)
    # the stream continues where it left off on the next step (offset bookkeeping)
    eng = got["engine"]
    off = int.from_bytes(bytes(gen.get_state()[8:16].tolist()), "little")
    assert eng._philox_offset == off
    # per-layer thresholds: every parameter is noised with sigma * ||C||_2, the L2 sensitivity of the per-layer
    # clipped sum (what the accountant's noise_multiplier = sigma assumes); per_layer_noise="own" -> sigma * C_k
    Cs = [3.0, 0.2, 0.5, 0.2, 1.0, 0.5]
    c_l2 = float(np.sqrt(np.sum(np.square(np.array(Cs, dtype=np.float64)))))
    for mode, stds in (("l2norm", [sigma * c_l2] * len(Cs)), ("own", [sigma * c for c in Cs])):
        Dg2 = copy.deepcopy(D).to(DEV)
        got2 = run_cuda(Dg2, real, fake, y, B, Cs, sigma=sigma, seed=seed, per_layer_noise=mode)
        assert got2["engine"].noise_stds() == pytest.approx(stds, rel=1e-12)
        gen.manual_seed(seed)
        for s, g, sd in zip(got2["summed"], got2["grads"], stds):
            noise = torch.normal(0.0, sd, s.shape, device=DEV, generator=gen)
            ref = s / B
            noise /= B
            ref += noise
            assert torch.equal(g, ref)
    # the oracle states the same rule
    oe = O.OracleGCEngine(copy.deepcopy(D), batch_size=B, noise_multiplier=sigma, max_grad_norm=Cs)
    assert oe.noise_stds() == pytest.approx([sigma * c_l2] * len(Cs), rel=1e-12)
    oe.remove()


def test_per_layer_noise_variance_is_sigma_times_l2_norm_of_thresholds():
    """per-layer clipping at C_k has L2 sensitivity ||C||_2; the empirical noise variance of every
    parameter must be (sigma * ||C||_2 / B)^2, also when the thresholds live on the device (adaptive clipping)."""
    D, shape, ncls, lo = make("mnist")
    B, sigma = 32, 2.0
    Cs = [3.0, 0.2, 0.5, 0.2, 1.0, 0.5]
    c_l2 = float(np.sqrt(np.sum(np.square(Cs))))
    real, fake, y = batch(shape, ncls, lo, B, seed=2)
    for dev_thresholds in (False, True):
        Dg = copy.deepcopy(D).to(DEV)
        opt = torch.optim.SGD(Dg.parameters(), lr=0.0)
        eng = cg.PrivacyEngine(Dg, batch_size=B, sample_size=60000, noise_multiplier=sigma,
                               max_grad_norm=torch.tensor(Cs, device=DEV) if dev_thresholds else Cs,
                               num_private_passes=1, auto_clip_and_accum_on_step=False)
        eng.attach(opt)
        eng._set_seed(11)
        d_loss(Dg, real.to(DEV), fake.to(DEV), y.to(DEV)).backward()
        eng.disable_hooks()
        eng.clip(); eng.accum_grads_across_passes(); eng.accumulate_batch()
        clean = [p.summed_grad.clone() / B for p in Dg.parameters()]
        opt.step()
        w = next(iter(Dg.parameters()))                       # lin1.weight: 101 632 elements
        z = (w.grad - clean[0]).flatten().double()
        want = sigma * c_l2 / B
        assert abs(z.std().item() / want - 1.0) < 0.02, (z.std().item(), want)
        assert abs(z.mean().item()) < 0.02 * want


def test_split_clip_fake_switch_and_penalty_on_summed_grad():
    """U1 switch: fake pass unclipped; and train.py:431 mutates p.summed_grad between accumulate_batch and step."""
    D, shape, ncls, lo = make("mnist_uncond")
    B = 24
    real, fake, y = batch(shape, ncls, lo, B, seed=9)
    ref_eng = O.OracleGCEngine(D, batch_size=B, noise_multiplier=0.0, max_grad_norm=1.0, split_clip_fake=False)
    d_loss(D, real, fake, y).backward()
    ref_eng.clip(); ref_eng.accum_grads_across_passes(); ref_eng.accumulate_batch()
    Dg = copy.deepcopy(D).to(DEV)
    for p in Dg.parameters():
        p.grad = None
    opt = torch.optim.SGD(Dg.parameters(), lr=0.0)
    eng = cg.PrivacyEngine(Dg, batch_size=B, sample_size=1000, noise_multiplier=0.0, max_grad_norm=1.0,
                           num_private_passes=1, split_clip_fake=False, auto_clip_and_accum_on_step=False)
    eng.attach(opt)
    d_loss(Dg, real.to(DEV), fake.to(DEV), None).backward()
    eng.disable_hooks()
    eng.clip(); eng.accum_grads_across_passes(); eng.accumulate_batch()
    for p, q in zip(Dg.parameters(), ref_eng.params()):
        assert rel(p.summed_grad, q.summed_grad) < REL_TOL
        p.summed_grad += torch.ones_like(p) * B          # what train.py:431 does with the penalty gradient
    opt.step()
    for p, q in zip(Dg.parameters(), ref_eng.params()):
        assert rel(p.grad, q.summed_grad / B + 1.0) < REL_TOL
    ref_eng.remove()
    with pytest.raises(ValueError):
        eng.step()                                         # nothing accumulated


@pytest.mark.parametrize("name,B,per_param", [("mnist", 64, False), ("mnist", 32, True), ("mnist_dcrn", 6, True),
                                              ("d64", 4, False)])
def test_immediate_sensitivity_matches_oracle(name, B, per_param):
    D, shape, ncls, lo = make(name)
    real, fake, y = batch(shape, ncls, lo, B, seed=11)
    x = real.clone().requires_grad_(True)
    of, af = D(fake, y)
    orr, ar = D(x, y)
    loss = D.real_loss(orr) + D.fake_loss(of)
    if ar is not None:
        loss = loss + D.aux_loss(ar, y) + D.aux_loss(af, y, fake=True)
    g_ref, s_ref, ps_ref = O.immediate_sensitivity(list(D.parameters()), loss, x, per_param=per_param)

    Dg = copy.deepcopy(D).to(DEV)
    opt = torch.optim.SGD(Dg.parameters(), lr=0.0)
    eng = cg.ISPrivacyEngine(Dg, batch_size=B, sample_size=60000, noise_multiplier=0.0, per_param=per_param)
    eng.attach(opt)
    eng._set_seed(5)
    xg = real.to(DEV).requires_grad_(True)
    yg = None if y is None else y.to(DEV)
    of, af = Dg(fake.to(DEV), yg)
    orr, ar = Dg(xg, yg)
    lg = Dg.real_loss(orr) + Dg.fake_loss(of)
    if ar is not None:
        lg = lg + Dg.aux_loss(ar, yg) + Dg.aux_loss(af, yg, fake=True)
    eng.backward(lg, xg)
    for p, gr in zip(Dg.parameters(), g_ref):
        assert rel(p.grad, gr) < 1e-4
    s = eng.batch_sensitivity
    if per_param:
        np.testing.assert_allclose(np.asarray(s), np.asarray(s_ref), rtol=2e-3, atol=1e-7)
    else:
        assert abs(s - s_ref) / s_ref < 2e-3
    np.testing.assert_allclose(eng._per_sample_sens.cpu().numpy(), ps_ref.numpy(), rtol=5e-3, atol=1e-6)
    before = [p.grad.clone() for p in Dg.parameters()]
    opt.step()                                            # sigma = 0: grads unchanged
    for p, b in zip(Dg.parameters(), before):
        assert torch.equal(p.grad, b)


def test_immediate_sensitivity_noise_bit_exact():
    D, shape, ncls, lo = make("mnist_uncond")
    B, sigma, seed = 16, 10.0, 99
    real, fake, y = batch(shape, ncls, lo, B, seed=2)
    Dg = copy.deepcopy(D).to(DEV)
    opt = torch.optim.SGD(Dg.parameters(), lr=0.0)
    eng = cg.ISPrivacyEngine(Dg, batch_size=B, sample_size=60000, noise_multiplier=sigma, per_param=False)
    eng.attach(opt)
    eng._set_seed(seed)
    xg = real.to(DEV).requires_grad_(True)
    loss = Dg.real_loss(Dg(xg)[0]) + Dg.fake_loss(Dg(fake.to(DEV))[0])
    eng.backward(loss, xg)
    s = torch.tensor(eng.batch_sensitivity, dtype=torch.float32, device=DEV)
    before = [p.grad.clone() for p in Dg.parameters()]
    opt.step()
    gen = torch.Generator(device=DEV)
    gen.manual_seed(seed)
    std = torch.tensor(sigma, dtype=torch.float32, device=DEV) * s          # fp32 product, like the kernel
    for p, b in zip(Dg.parameters(), before):
        z = torch.normal(0.0, 1.0, p.shape, device=DEV, generator=gen)
        assert torch.equal(p.grad, b + z * std)


@pytest.mark.parametrize("name", ["mnist_vanilla_acgan", "mnist_dcrn_acgan", "celeba_d64_uncond", "celeba_d48_uncond"])
def test_gradient_penalty_matches_reference_golden(golden_dir, name):
    """calc_lipschitz_penalty_WRT / calc_WGAN_GP_penalty with the CUDA row-norm vs outputs of the
    reference's gradient_penalty.py (generated by oracle/gen_golden.py)."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    ctor = {"mnist_vanilla_acgan": lambda: DD.MNISTVanillaD(n_classes=10, conditional_arch="ACGAN", aux_loss_type="cross_entropy"),
            "mnist_dcrn_acgan": lambda: DD.MNIST_DCRN_D(n_classes=10, conditional_arch="ACGAN", aux_loss_type="wasserstein"),
            "celeba_d64_uncond": lambda: DD.CelebA_DCRN_D64(n_classes=0),
            "celeba_d48_uncond": lambda: DD.CelebA_DCRN_D48(n_classes=0)}[name]
    torch.manual_seed(42)
    D = ctor().to(DEV)
    x = torch.from_numpy(g["x"]).to(DEV)
    y = torch.from_numpy(g["y"]).to(DEV) if g["y"].size else None
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    pen = cg.calc_lipschitz_penalty_WRT(D, x, y, per_sample=True, one_sided=False, aux_penalty=True)
    np.testing.assert_allclose(pen.detach().cpu().numpy(), g["lip_pen_two_sided_aux"], rtol=2e-3, atol=1e-5)
    pen1 = cg.calc_lipschitz_penalty_WRT(D, x, y, per_sample=True, one_sided=True, aux_penalty=False)
    np.testing.assert_allclose(pen1.detach().cpu().numpy(), g["lip_pen_one_sided_noaux"], rtol=2e-3, atol=1e-5)
    torch.manual_seed(777)
    alpha = torch.rand(x.shape[0], 1)                                      # CPU draw, like gradient_penalty.py:33
    gp = cg.calc_penalty(D, ["WGAN-GP"], x, y, torch.from_numpy(g["fake"]).to(DEV), y, aux_penalty=True, alpha=alpha)
    np.testing.assert_allclose(gp.detach().cpu().numpy(), g["wgan_gp_seed777"], rtol=2e-3, atol=1e-5)
    grads = torch.autograd.grad(gp, list(D.parameters()), allow_unused=True)
    got = np.array([0.0 if t is None else t.double().norm().item() for t in grads])
    np.testing.assert_allclose(got, g["wgan_gp_grad_norms"], rtol=5e-3, atol=1e-6)


def test_engine_backward_skips_weight_grads_and_handles_channels_last():
    """engine.backward(loss) delivers the same per-sample machinery inputs as loss.backward() without
    autograd's own weight gradients; a channels_last critic gives the same result."""
    D, shape, ncls, lo = make("mnist_dcrn")
    B = 6
    real, fake, y = batch(shape, ncls, lo, B, seed=21)
    ref = run_oracle(copy.deepcopy(D), real, fake, y, B, 0.3)
    for cl in (False, True):
        Dg = copy.deepcopy(D).to(DEV)
        if cl:
            Dg = Dg.to(memory_format=torch.channels_last)
        opt = torch.optim.SGD(Dg.parameters(), lr=0.0)
        eng = cg.PrivacyEngine(Dg, batch_size=B, sample_size=1000, noise_multiplier=0.0, max_grad_norm=0.3,
                               num_private_passes=1, auto_clip_and_accum_on_step=False)
        eng.attach(opt)
        eng.backward(d_loss(Dg, real.to(DEV), fake.to(DEV), None))
        assert all(p.grad is None for p in Dg.parameters())          # autograd never touched the weights
        eng.disable_hooks()
        eng.clip(); eng.accum_grads_across_passes(); eng.accumulate_batch()
        opt.step()
        for p, g in zip(Dg.parameters(), ref["grads"]):
            assert rel(p.grad, g) < REL_TOL


def test_backprop_clipper_matches_pure_torch_restatement():
    """Wrapped critic (CUDA l2_clip in forward and on grad-inputs) vs the same clipping written with the
    oracle's l2_clip (reference backprop_clip.py:18-22, 98-103)."""
    from csl_gan_b200.backprop_clip import BackpropClipper
    torch.manual_seed(1)
    D = DD.MNIST_DCRN_D(n_classes=0).to(DEV)
    Dref = copy.deepcopy(D)
    x = torch.randn(6, 1, 28, 28, device=DEV) * 3
    back, fwd = [0.05, 0.02, 0.5], [5.0, 30.0, 10.0]
    bc = BackpropClipper(D, back, fwd, device=DEV, input_size=(1, 1, 28, 28))
    out = D(x)[0]
    out.sum().backward()

    # restatement: clip inputs forward, clip the gradient flowing out of each layer's input backward
    class ClipGrad(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t, c):
            ctx.c = c
            return t.view_as(t)

        @staticmethod
        def backward(ctx, g):
            return O.l2_clip(g, ctx.c), None

    h = x
    layers = list(Dref.blocks) + [Dref.linOut]
    for i, layer in enumerate(layers):
        if i == len(layers) - 1:
            h = h.reshape(h.size(0), -1)
        # the reference wraps the LAYER: clip(x) -> layer -> dummy whose grad_input (= grad wrt layer output) is clipped
        h = ClipGrad.apply(layer(O.l2_clip(h, fwd[i])), back[i])
        if i < len(layers) - 1:
            h = torch.nn.functional.leaky_relu(h, 0.2)
    h.sum().backward()
    assert rel(out, h) < 1e-5
    for (n, p), q in zip(D.named_parameters(), Dref.parameters()):
        assert rel(p.grad, q.grad) < 2e-3, n          # cuDNN TF32-off vs itself: same library, tiny differences
    bc.disable_hooks()
    D.zero_grad()
    D(x)[0].sum().backward()                          # hooks off: only the forward clip remains


def test_engine_rejects_cpu_module_and_batchnorm():
    D, *_ = make("mnist_uncond")
    with pytest.raises(cg.CslGanCudaError):
        cg.PrivacyEngine(D, batch_size=4, sample_size=100, noise_multiplier=1.0, max_grad_norm=1.0)
    bad = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.BatchNorm1d(4)).to(DEV)
    with pytest.raises(NotImplementedError):
        cg.PrivacyEngine(bad, batch_size=4, sample_size=100, noise_multiplier=1.0, max_grad_norm=1.0)


def test_device_mean_sampler_matches_reference_expressions():
    """DeviceMeanSampler on the GPU against the reference's expressions (mean_sampler.py:47-64 class means,
    :75-84 sample): exact with the noise switched off, and the two noise terms have the reference's structure (one
    offset per sample + iid pixel noise) and scale with it on."""
    from csl_gan_b200.mean_sampler import DeviceMeanSampler
    g = torch.Generator().manual_seed(3)
    n_classes, mean_size, n_batches = 3, 20, 4
    batches = [(torch.rand(90, 1, 8, 8, generator=g), torch.randint(0, n_classes, (90,), generator=g)) for _ in range(n_batches)]
    ms = DeviceMeanSampler.from_batches(batches, n_classes, mean_size, 0.0, 60000, device=DEV)
    assert ms.mean_samples.is_cuda and tuple(ms.mean_samples.shape) == (n_classes, n_batches, 1, 8, 8)
    for i, (x, y) in enumerate(batches):
        for c in range(n_classes):
            s = x[y == c]
            want = (s[:mean_size] if len(s) > mean_size else s).sum(dim=0) / mean_size       # reference :57-59
            assert torch.allclose(ms.mean_samples[c, i].cpu(), want, atol=1e-6)
    # sample(): every block of num_samples draws is a permutation of the stored means (reference :76), labels honoured
    labels = torch.randint(0, n_classes, (10,), generator=g)
    r, lab = ms.sample(10, noise_std=0.0, noise_mean_std=0.0, requested_labels=labels)
    assert r.is_cuda and torch.equal(lab.cpu(), labels)
    stored = ms.mean_samples.cpu()
    for blk in range(0, 10, n_batches):
        idx = []
        for k in range(blk, min(blk + n_batches, 10)):
            hit = [j for j in range(n_batches) if torch.equal(r[k].cpu(), stored[labels[k], j])]
            assert len(hit) >= 1
            idx.append(hit[0])
        assert len(set(idx)) == len(idx)                                                     # no repeats inside a block
    # noise structure: r - mean = per-sample constant (std noise_mean_std) + iid pixel noise (std noise_std)
    big = DeviceMeanSampler(torch.zeros(1, 4, 1, 32, 32, device=DEV), 0.1, 100, 60000)
    r, lab = big.sample(512, noise_std=0.05, noise_mean_std=0.2)
    assert lab is None and tuple(r.shape) == (512, 1, 32, 32)
    per_sample = r.mean(dim=(1, 2, 3))
    assert abs(per_sample.std().item() - 0.2) < 0.03
    assert abs((r - per_sample.view(-1, 1, 1, 1)).std().item() - 0.05) < 0.005
    eps, alpha = big.get_privacy_cost(1e-6)
    assert eps > 0 and alpha > 1
