"""Kernel-level parity through the C ABI (ctypes) against the CPU oracle / plain fp64 math."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from csl_gan_b200 import _lib as L  # noqa: E402
from csl_gan_b200 import functional as FN  # noqa: E402
from oracle import dp_oracle as O  # noqa: E402
from oracle import philox as PH  # noqa: E402

DEV = "cuda"


def tf32(x: torch.Tensor) -> torch.Tensor:
    """round-to-nearest (ties away) to 10 explicit mantissa bits, like cvt.rna.tf32.f32"""
    i = x.contiguous().view(torch.int32)
    r = (i + 0x1000) & ~0x1FFF
    return r.view(torch.float32)


def st():
    return L.stream_ptr()


def contract_plain(X, Y, epi=L.EPI_ACCUM, block_n=0, n_split=1, max_ctas=0):
    """out[M][N] = X[M,K] @ Y[N,K]^T through cg_contract (KH=KW=1, C=N, slots of 32 columns)."""
    M, K = X.shape
    N = Y.shape[0]
    assert K % 32 == 0
    units = K // 32
    d = L.ContractDesc()
    d.X, d.x_pitch, d.x_rows, d.x_cols = X.data_ptr(), X.stride(0), M, K
    d.Y, d.y_pitch, d.y_rows, d.y_cols = Y.data_ptr(), Y.stride(0), N, K
    d.M, d.C, d.KH, d.KW = M, N, 1, 1
    d.nkb = 1
    d.x_slot_stride = d.y_slot_stride = 32
    d.group_mode = L.GROUP_SPLITK
    spg = (units + n_split - 1) // n_split
    d.n_groups = (units + spg - 1) // spg
    d.slot_lo, d.slot_hi, d.spg = 0, units, spg
    d.n_seg, d.seg_stride = 1, 1
    out = torch.zeros((M, N), device=DEV)
    d.epi, d.out, d.out_group_stride = epi, out.data_ptr(), 0
    d.block_n, d.max_ctas = block_n, max_ctas
    L.call("cg_contract", C.byref(d), st())
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,N,K,bn,split", [
    (128, 128, 32, 0, 1),       # one MMA k-block
    (128, 128, 256, 0, 1),      # pipeline wraps the 5-stage ring
    (64, 75, 1024, 0, 4),       # ragged M and N, split-K atomics
    (256, 300, 96, 64, 1),      # several m/n tiles, explicit BN
    (1, 16, 64, 16, 2),         # degenerate rows
    (512, 640, 2048, 0, 8),     # many items per CTA
])
def test_tcgen05_gemm_matches_fp64(M, N, K, bn, split):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    X = tf32(torch.randn(M, K, generator=g).to(DEV))
    Y = tf32(torch.randn(N, K, generator=g).to(DEV))
    out = contract_plain(X, Y, block_n=bn, n_split=split)
    ref = (X.double() @ Y.double().t()).float()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err


def test_tcgen05_gemm_persistent_few_ctas():
    """max_ctas=3 forces every CTA through many items: exercises ring/accumulator phase flips."""
    g = torch.Generator().manual_seed(5)
    X = tf32(torch.randn(384, 160, generator=g).to(DEV))
    Y = tf32(torch.randn(500, 160, generator=g).to(DEV))
    out = contract_plain(X, Y, max_ctas=3)
    ref = (X.double() @ Y.double().t()).float()
    assert (out - ref).abs().max().item() / ref.abs().max().item() < 2e-5


def _conv_case(B, Cin, Cout, H, W, k, s, p, d=1, seed=0):
    g = torch.Generator().manual_seed(seed)
    conv = torch.nn.Conv2d(Cin, Cout, k, stride=s, padding=p, dilation=d)
    A = torch.randn(B, Cin, H, W, generator=g)
    Ho = (H + 2 * p - d * (k - 1) - 1) // s + 1
    Wo = (W + 2 * p - d * (k - 1) - 1) // s + 1
    Bp = torch.randn(B, Cout, Ho, Wo, generator=g)
    return conv, A, Bp, Ho, Wo


@pytest.mark.parametrize("B,Cin,Cout,H,W,k,s,p,d", [
    (3, 3, 64, 64, 64, 5, 2, 2, 1),      # D64 blocks.0
    (2, 64, 128, 32, 32, 5, 2, 2, 1),    # D64 blocks.1
    (2, 128, 256, 16, 16, 5, 2, 2, 1),   # D64 blocks.2 (clipped sum on CTA pairs, odd number of half tiles)
    (2, 256, 512, 8, 8, 5, 2, 2, 1),     # D64 blocks.3 (Q = 16 < one k-block; CTA pairs, two 256-row tile pairs)
    (3, 1, 64, 28, 28, 5, 2, 2, 1),      # MNIST DCRN blocks.0 (Q = 196, ragged k tail)
    (2, 64, 128, 14, 14, 5, 2, 2, 1),    # MNIST DCRN blocks.1 (Q = 49)
    (2, 5, 7, 9, 11, 3, 1, 1, 1),        # stride 1
    (2, 4, 6, 13, 12, 3, 2, 0, 2),       # dilation 2, no padding
    (2, 3, 5, 10, 10, 4, 3, 1, 1),       # stride 3, even kernel
])
@pytest.mark.parametrize("path", ["channels_last", "channels_last_from_cl_tensors", "kw_planes", "channels_last_tf32"])
def test_conv_per_sample_grads_store_sumsq_accum(B, Cin, Cout, H, W, k, s, p, d, path):
    """staging + contraction (all three epilogues) vs the oracle's unfold/einsum, on both operand paths:
    channels-last MN-major (main; FP16 operand containers by default, TF32 words in the *_tf32 variant) and kw-plane
    K-major (geometries the main path cannot tile)."""
    from csl_gan_b200.grad_sample import LayerPlan
    conv, A, Bp, Ho, Wo = _conv_case(B, Cin, Cout, H, W, k, s, p, d)
    if path != "kw_planes" and not L.cl_supported(Ho, Wo):
        pytest.skip("window grid not tileable by the channels-last path")
    gw_ref, gb_ref = O.conv2d_grad_sample(conv, A, Bp)                # [B, Cout, Cin, k, k], [B, Cout]
    conv = conv.to(DEV)
    plan = LayerPlan("conv", conv, 0, 1)
    plan.force_legacy = path == "kw_planes"
    plan.use_ghost = False
    plan.use_half = path != "channels_last_tf32"
    path = "channels_last" if path == "channels_last_tf32" else path
    Bpad = 32
    Ag, Bg = A.to(DEV), Bp.to(DEV)
    if path == "channels_last_from_cl_tensors":
        Ag, Bg = Ag.contiguous(memory_format=torch.channels_last), Bg.contiguous(memory_format=torch.channels_last)
    plan.capture_activation(Ag, 0, Bpad, 1)
    plan.capture_backprop(Bg, 0, 1.0)
    assert plan.path == ("kw_planes" if path == "kw_planes" else "channels_last")
    # STORE
    gs = plan.materialize(0, B).cpu()
    scale = gw_ref.abs().max().item()
    assert (gs - gw_ref).abs().max().item() / scale < 3e-3
    for n in range(B):
        rel = (gs[n] - gw_ref[n]).norm() / gw_ref[n].norm()
        assert rel < 1e-3, (n, rel)
    # bias rows
    np.testing.assert_allclose(plan.live_bias_rows()[:B].cpu().numpy(), gb_ref.numpy(), rtol=1e-4, atol=1e-4)
    # SUMSQ
    norm2 = torch.zeros(Bpad, device=DEV)
    plan.weight_norm2(norm2, 0, B)
    torch.cuda.synchronize()
    ref_n2 = gw_ref.reshape(B, -1).double().pow(2).sum(1)
    np.testing.assert_allclose(norm2[:B].cpu().double().numpy(), ref_n2.numpy(), rtol=2e-3)
    # ACCUM with per-sample factors
    f = torch.rand(Bpad, device=DEV) + 0.25
    plan.scale_backprops(f, 0, B)
    out = torch.empty_like(conv.weight)
    plan.weighted_sum(out, 0, B, 148, accumulate=False, factor_row=f)     # factor_row: thin-layer path only
    torch.cuda.synchronize()
    ref = torch.einsum("n,n...->...", f[:B].cpu(), gw_ref)
    assert ((out.cpu() - ref).norm() / ref.norm()).item() < 1e-3


@pytest.mark.parametrize("B,Cin,Cout,H,W,k,s,p", [
    (40, 128, 256, 16, 16, 5, 2, 2),     # Q = 64: 80 k-blocks, 13 tile pairs (the last one half empty)
    (70, 256, 512, 8, 8, 3, 1, 1),       # Q = 64, 3x3 stride 1: 18 half tiles, two 256-row pairs
    (33, 384, 256, 4, 4, 5, 2, 2),       # Q = 4: 8 samples per k-block, ragged last k-block, 3 half tiles per tap
])
@pytest.mark.parametrize("operands", ["f16", "tf32"])
def test_cta_pair_clipped_sum_matches_single_cta_and_reference(B, Cin, Cout, H, W, k, s, p, operands):
    """cl_pair_kernel (cluster of 2, tcgen05 cta_group::2) vs cl_contract_kernel vs the oracle, split-K over
    several groups so the pipeline wraps and both accumulator stages are used."""
    from csl_gan_b200.grad_sample import LayerPlan
    conv, A, Bp, Ho, Wo = _conv_case(B, Cin, Cout, H, W, k, s, p, 1, seed=3)
    gw_ref, _ = O.conv2d_grad_sample(conv, A, Bp)
    conv = conv.to(DEV)
    plan = LayerPlan("conv", conv, 0, 1)
    plan.use_half = operands == "f16"
    Bpad = (B + 31) // 32 * 32
    plan.capture_activation(A.to(DEV).contiguous(memory_format=torch.channels_last), 0, Bpad, 1)
    plan.capture_backprop(Bp.to(DEV).contiguous(memory_format=torch.channels_last), 0, 1.0)
    assert plan.impl is not None and plan.impl.pair
    f = torch.rand(Bpad, device=DEV) + 0.25
    f[B:] = 0
    plan.scale_backprops(f, 0, Bpad)
    ref = torch.einsum("n,n...->...", f[:B].cpu(), gw_ref)
    outs = {}
    for pair in (True, False):
        plan.impl.pair = pair
        for sms in (148, 6):                       # many groups / few CTAs (persistent loop over items)
            out = torch.empty_like(conv.weight)
            plan.weighted_sum(out, 0, B, sms, accumulate=False, factor_row=f)
            torch.cuda.synchronize()
            outs[(pair, sms)] = out.cpu()
            assert ((out.cpu() - ref).norm() / ref.norm()).item() < 1e-3, (pair, sms)
    # same products, different summation order only
    assert ((outs[(True, 148)] - outs[(False, 148)]).norm() / ref.norm()).item() < 1e-5


@pytest.mark.parametrize("B,Cin,Cout,H,W,k,s,p", [
    (5, 256, 512, 8, 8, 5, 2, 2),        # D64 blocks.3: Q = 16, 8 samples per tile, ragged last group
    (3, 128, 256, 16, 16, 5, 2, 2),      # D64 blocks.2: Q = 64, 2 samples per tile
    (9, 6, 10, 8, 8, 3, 2, 1),           # Q = 16, C not a multiple of 4 or 32, O not a multiple of 32
    (4, 8, 8, 4, 4, 3, 1, 1),            # stride 1, Q = 16
    (2, 16, 24, 16, 32, 3, 2, 1),        # Q = 128, one sample per tile, non-square
])
@pytest.mark.parametrize("operands", ["f16", "tf32"])
def test_ghost_norms_match_direct_and_oracle(B, Cin, Cout, H, W, k, s, p, operands):
    from csl_gan_b200.grad_sample import LayerPlan
    conv, A, Bp, Ho, Wo = _conv_case(B, Cin, Cout, H, W, k, s, p)
    gw_ref, _ = O.conv2d_grad_sample(conv, A, Bp)
    ref_n2 = gw_ref.reshape(B, -1).double().pow(2).sum(1).numpy()
    conv = conv.to(DEV)
    out = {}
    for ghost in (True, False):
        plan = LayerPlan("conv", conv, 0, 1)
        plan.use_ghost = ghost
        plan.use_half = operands == "f16"
        plan.capture_activation(A.to(DEV), 1, 32, 2)          # second pass: exercises slot offsets
        plan.capture_backprop(Bp.to(DEV), 1, 1.0)
        assert plan.path == "channels_last" and plan.impl.ghost == ghost
        norm2 = torch.zeros(64, device=DEV)
        plan.weight_norm2(norm2, 1, B)
        torch.cuda.synchronize()
        assert float(norm2[:32].abs().sum()) == 0.0 and float(norm2[32 + B:].abs().sum()) == 0.0
        out[ghost] = norm2[32:32 + B].cpu().double().numpy()
    np.testing.assert_allclose(out[True], ref_n2, rtol=2e-3)
    np.testing.assert_allclose(out[True], out[False], rtol=2e-3)


def test_conv_transpose_per_sample_grads():
    from csl_gan_b200.grad_sample import LayerPlan
    g = torch.Generator().manual_seed(3)
    ct = torch.nn.ConvTranspose2d(6, 4, 4, stride=2, padding=1)
    B = 3
    A = torch.randn(B, 6, 7, 5, generator=g)
    out_shape = ct(A).shape
    Bp = torch.randn(out_shape, generator=g)
    gw_ref, gb_ref = O.conv_transpose2d_grad_sample(ct, A, Bp)
    # anchor the oracle restatement itself on autograd (micro-batch)
    for n in range(B):
        ct.zero_grad()
        (ct(A[n:n + 1]) * Bp[n:n + 1]).sum().backward()
        assert torch.allclose(ct.weight.grad, gw_ref[n], rtol=1e-4, atol=1e-5)
        assert torch.allclose(ct.bias.grad, gb_ref[n], rtol=1e-4, atol=1e-5)
    ct = ct.to(DEV)
    plan = LayerPlan("convT", ct, 0, 1)
    plan.capture_activation(A.to(DEV), 0, 32, 1)
    plan.capture_backprop(Bp.to(DEV), 0, 1.0)
    assert plan.path == "kw_planes"                       # 7x5 window grid: not tileable by the main path
    gs = plan.materialize(0, B).cpu()
    for n in range(B):
        assert ((gs[n] - gw_ref[n]).norm() / gw_ref[n].norm()).item() < 1e-3
    np.testing.assert_allclose(plan.live_bias_rows()[:B].cpu().numpy(), gb_ref.numpy(), rtol=1e-4, atol=1e-4)
    # a tileable geometry goes through the channels-last path (8x8 inputs, roles of the operands swapped)
    ct2 = torch.nn.ConvTranspose2d(6, 4, 4, stride=2, padding=1)
    A2 = torch.randn(B, 6, 8, 8, generator=g)
    Bp2 = torch.randn(ct2(A2).shape, generator=g)
    gw2, gb2 = O.conv_transpose2d_grad_sample(ct2, A2, Bp2)
    ct2 = ct2.to(DEV)
    plan2 = LayerPlan("convT", ct2, 0, 1)
    plan2.use_ghost = False
    plan2.capture_activation(A2.to(DEV), 0, 32, 1)
    plan2.capture_backprop(Bp2.to(DEV), 0, 1.0)
    assert plan2.path == "channels_last"
    gs2 = plan2.materialize(0, B).cpu()
    for n in range(B):
        assert ((gs2[n] - gw2[n]).norm() / gw2[n].norm()).item() < 1e-3
    np.testing.assert_allclose(plan2.live_bias_rows()[:B].cpu().numpy(), gb2.numpy(), rtol=1e-4, atol=1e-4)
    n2 = torch.zeros(32, device=DEV)
    plan2.weight_norm2(n2, 0, B)
    np.testing.assert_allclose(n2[:B].cpu().double().numpy(), gw2.reshape(B, -1).double().pow(2).sum(1).numpy(), rtol=2e-3)


@pytest.mark.parametrize("B,P,Od", [(600, 794, 128), (37, 128, 10), (5, 128, 1), (64, 8192, 1)])
def test_linear_closed_form_norms_and_weighted_sum(B, P, Od):
    from csl_gan_b200.grad_sample import LayerPlan, _round_up
    g = torch.Generator().manual_seed(B)
    lin = torch.nn.Linear(P, Od).to(DEV)
    A = torch.randn(B, P, generator=g)
    Bp = torch.randn(B, Od, generator=g)
    gw_ref, gb_ref = O.linear_grad_sample(A, Bp)
    plan = LayerPlan("lin", lin, 0, 1)
    Bpad = _round_up(B, 32)
    for ps in range(2):                               # two passes, second one is what we check
        plan.capture_activation(A.to(DEV) * (ps + 1), ps, Bpad, 2)
        plan.capture_backprop(Bp.to(DEV), ps, 1.0)
    norm2 = torch.zeros(2 * Bpad, device=DEV)
    plan.weight_norm2(norm2, 1, B)
    ref = (2 * gw_ref).reshape(B, -1).double().pow(2).sum(1)
    np.testing.assert_allclose(norm2[Bpad:Bpad + B].cpu().double().numpy(), ref.numpy(), rtol=1e-4)
    bn2 = torch.zeros(2 * Bpad, device=DEV)
    plan.bias_norm2(bn2, 1, B)
    np.testing.assert_allclose(bn2[Bpad:Bpad + B].cpu().numpy(), gb_ref.pow(2).sum(1).numpy(), rtol=1e-4)
    f = torch.rand(2 * Bpad, device=DEV) + 0.25
    plan.scale_backprops(f, 0, Bpad + B)
    out = torch.empty_like(lin.weight)
    plan.weighted_sum(out, 0, Bpad + B, 148, accumulate=False)
    ref = torch.einsum("n,n...->...", f[:B].cpu(), gw_ref) + 2 * torch.einsum("n,n...->...", f[Bpad:Bpad + B].cpu(), gw_ref)
    assert ((out.cpu() - ref).norm() / ref.norm()).item() < 1e-3
    outb = torch.empty(Od, device=DEV)
    plan.bias_weighted_sum(outb, f, 0, Bpad + B, accumulate=False)
    refb = (f[:B].cpu()[:, None] * gb_ref).sum(0) + (f[Bpad:Bpad + B].cpu()[:, None] * gb_ref).sum(0)
    np.testing.assert_allclose(outb.cpu().numpy(), refb.numpy(), rtol=1e-4, atol=1e-4)
    gs = plan.materialize(1, B).cpu()
    assert ((gs - 2 * gw_ref).norm() / (2 * gw_ref).norm()).item() < 1e-3


def test_clip_factors_flat_and_per_layer():
    n_params, S = 4, 70
    norm2 = (torch.rand(n_params, S) * 30).to(DEV)
    fac = torch.empty(n_params, S, device=DEV)
    nout = torch.empty(n_params, S, device=DEV)
    Cs = torch.tensor([2.0, 3.0, 0.5, 10.0], device=DEV)
    L.call("cg_clip_factors", norm2.data_ptr(), n_params, S, 1, Cs.data_ptr(), 1.0, 0, S, fac.data_ptr(), nout.data_ptr(), st())
    norms = [norm2[k].sqrt().cpu().view(1, S) for k in range(n_params)]
    ref = O.calc_clipping_factors(norms, [2.0, 3.0, 0.5, 10.0], n_params)
    for k in range(n_params):
        np.testing.assert_allclose(fac[k].cpu().numpy(), ref[k][0].numpy(), rtol=1e-6)
        np.testing.assert_allclose(nout[k].cpu().numpy(), norms[k][0].numpy(), rtol=1e-6)
    L.call("cg_clip_factors", norm2.data_ptr(), n_params, S, 0, Cs.data_ptr(), 1.0, 10, S, fac.data_ptr(), nout.data_ptr(), st())
    flat = torch.stack([n[0] for n in norms]).norm(2, dim=0)
    ref = (2.0 / (flat + 1e-6)).clamp(max=1.0)
    ref[:10] = 1.0
    np.testing.assert_allclose(fac[0].cpu().numpy(), ref.numpy(), rtol=2e-6)
    assert ((fac[0] < 0.999).any() and (fac[0] == 1.0).any())


@pytest.mark.parametrize("numel", [1, 10, 1000, 103179 - 101632, 101632, 4 * 1184 * 256 + 17, 3276800])
def test_noise_bit_exact_vs_torch_cuda_generator(numel):
    """The noise term must be bit-identical to torch.normal with a CUDA generator (north_star)."""
    seed, std, Bsz = 1234567, 10.0 * 4.0, 600
    gen = torch.Generator(device=DEV)
    gen.manual_seed(seed)
    burn = torch.normal(0.0, 1.0, (4096,), device=DEV, generator=gen)       # advance the offset first
    state = gen.get_state()
    offset = int.from_bytes(bytes(state[8:16].tolist()), "little")
    assert int.from_bytes(bytes(state[0:8].tolist()), "little") == seed
    summed = torch.randn(numel, device=DEV) * 50
    # reference op order (upstream privacy_engine.step): grad = summed / B; noise /= B; grad += noise
    noise = torch.normal(0.0, std, (numel,), device=DEV, generator=gen)
    ref = summed / Bsz
    ref += noise / Bsz
    out = torch.empty(numel, device=DEV)
    inc = C.c_ulonglong(0)
    L.call("cg_noise_finalize", summed.data_ptr(), out.data_ptr(), numel, float(Bsz), std, float(Bsz), seed, offset,
           C.byref(inc), st())
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    state2 = gen.get_state()
    assert int.from_bytes(bytes(state2[8:16].tolist()), "little") == offset + inc.value
    # pure-noise mode and the numpy restatement of the stream (few-ulp float stage)
    pure = torch.empty(numel, device=DEV)
    L.call("cg_noise_finalize", None, pure.data_ptr(), numel, 0.0, 1.0, 0.0, seed, offset, C.byref(inc), st())
    sm, mt, _, _ = L.device_info()
    if numel <= 200000:
        z = PH.torch_cuda_standard_normal(numel, seed, offset, sm, mt)
        np.testing.assert_allclose(pure.cpu().numpy(), z, rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("rows,cols", [(600, 784), (128, 12288), (7, 5), (1, 3276800), (33, 255)])
def test_row_l2_norm_forward_backward(rows, cols):
    g = torch.Generator().manual_seed(rows)
    t = torch.randn(rows, cols, generator=g).to(DEV).requires_grad_(True)
    n = FN.row_l2_norm(t)
    ref = O.row_l2_norm(t.detach().cpu().double())
    np.testing.assert_allclose(n.detach().cpu().double().numpy(), ref.numpy(), rtol=2e-6)
    go = torch.randn(rows, generator=g).to(DEV)
    (gi,) = torch.autograd.grad(n, t, go, create_graph=True)
    t2 = t.detach().clone().requires_grad_(True)
    (gi_ref,) = torch.autograd.grad(t2.norm(2, dim=1), t2, go, create_graph=True)
    assert torch.allclose(gi, gi_ref, rtol=1e-5, atol=1e-6)
    # double backward (immediate sensitivity differentiates through this node)
    w = torch.randn(rows, cols, generator=g).to(DEV)
    (gg,) = torch.autograd.grad((gi * w).sum(), t)
    (gg_ref,) = torch.autograd.grad((gi_ref * w).sum(), t2)
    assert torch.allclose(gg, gg_ref, rtol=1e-4, atol=1e-5)


def test_l2_clip_matches_reference_golden(golden_dir):
    gz = np.load(os.path.join(golden_dir, "l2_clip.npz"))
    for j in range(4):
        t = torch.from_numpy(gz[f"in_{j}"]).to(DEV)
        for Cc in (0.5, 5.0, 50.0):
            out = FN.l2_clip(t, Cc)
            np.testing.assert_allclose(out.cpu().numpy(), gz[f"out_{j}_C{Cc}"], rtol=2e-6, atol=1e-7)
    t = torch.randn(4, 3, 5, 5, device=DEV, requires_grad=True)
    t2 = t.detach().clone().requires_grad_(True)
    w = torch.randn_like(t)
    (g1,) = torch.autograd.grad((FN.l2_clip(t, 2.0) * w).sum(), t)
    (g2,) = torch.autograd.grad((O.l2_clip(t2, 2.0) * w).sum(), t2)
    assert torch.allclose(g1, g2, rtol=1e-4, atol=1e-5)


def test_vec_max_and_row_stat():
    v = torch.randn(1000, device=DEV)
    assert FN.vec_max(v).item() == v.max().item()
    norms = torch.rand(5, 96, device=DEV)
    out = torch.empty(5, device=DEV)
    L.call("cg_row_stat", norms.data_ptr(), 5, 96, 32, 90, 0, 1.5, out.data_ptr(), st())
    np.testing.assert_allclose(out.cpu().numpy(), (norms[:, 32:90].mean(1) * 1.5).cpu().numpy(), rtol=1e-5)
    L.call("cg_row_stat", norms.data_ptr(), 5, 96, 32, 90, 1, 2.0, out.data_ptr(), st())
    np.testing.assert_allclose(out.cpu().numpy(), (norms[:, 32:90].max(1).values * 2.0).cpu().numpy(), rtol=1e-6)


def test_errors_are_loud():
    with pytest.raises(L.CslGanCudaError):
        L.require_cuda_f32(torch.zeros(3), "cpu tensor")
    d = L.ContractDesc()
    d.KH = 99
    with pytest.raises(L.CslGanCudaError):
        L.call("cg_contract", C.byref(d), st())


@pytest.mark.parametrize("layer", ["conv_q256", "conv_q16_pair", "linear"])
def test_fp16_operands_survive_a_wide_dynamic_range(layer):
    """Adversarial case for the FP16 operand containers: per-sample magnitudes spread over 2^+-25 in
    BOTH operands (far outside FP16's 2^-24 .. 2^16; the per-sample gradient norms then span 2^+-50, about what fp32
    itself can square), plus 2^20 of spread inside every sample.  The exact per-sample
    power-of-two scales keep per-sample norms within 1e-3 of the fp64 reference for EVERY sample, and the clipped sum
    within 1e-3 normwise (with clipping active the factors equalise the samples, so small ones matter)."""
    from csl_gan_b200.grad_sample import LayerPlan
    g = torch.Generator().manual_seed(11)
    B, Bpad = 24, 32
    if layer == "linear":
        mod = torch.nn.Linear(200, 64, bias=False)
        A = torch.randn(B, 200, generator=g)
        Bp = torch.randn(B, 64, generator=g)
    else:
        Cin, Cout, H = (64, 128, 32) if layer == "conv_q256" else (128, 256, 8)
        mod = torch.nn.Conv2d(Cin, Cout, 5, stride=2, padding=2, bias=False)
        A = torch.randn(B, Cin, H, H, generator=g)
        Bp = torch.randn(B, Cout, H // 2, H // 2, generator=g)
    # spread inside a sample: a log-uniform envelope over 2^-20 .. 1 per element
    A = A * torch.exp2(-20 * torch.rand(A.shape, generator=g))
    Bp = Bp * torch.exp2(-20 * torch.rand(Bp.shape, generator=g))
    # spread across samples: 2^-25 .. 2^25, independently for the two operands
    ea = torch.randint(-25, 26, (B,), generator=g).float()
    eb = torch.randint(-25, 26, (B,), generator=g).float()
    ea[0], eb[0], ea[1], eb[1] = 25, 25, -25, -25
    A = A * torch.exp2(ea).view(B, *([1] * (A.dim() - 1)))
    Bp = Bp * torch.exp2(eb).view(B, *([1] * (Bp.dim() - 1)))
    if layer == "linear":
        gw_ref = torch.einsum("ni,nj->nij", Bp.double(), A.double())
    else:
        U = F.unfold(A.double(), 5, padding=2, stride=2)
        gw_ref = torch.einsum("noq,npq->nop", Bp.double().reshape(B, Bp.shape[1], -1), U).reshape(B, *mod.weight.shape)
    n_ref = gw_ref.reshape(B, -1).norm(dim=1)
    mod = mod.to(DEV)
    plan = LayerPlan(layer, mod, 0, None)
    plan.use_half = True
    plan.use_ghost = layer == "conv_q16_pair"
    plan.capture_activation(A.to(DEV), 0, Bpad, 1)
    plan.capture_backprop(Bp.to(DEV), 0, 1.0)
    assert plan.impl is not None and plan.impl.half
    norm2 = torch.zeros(Bpad, device=DEV)
    plan.weight_norm2(norm2, 0, B)
    torch.cuda.synchronize()
    np.testing.assert_allclose(norm2[:B].sqrt().cpu().double().numpy(), n_ref.numpy(), rtol=1e-3)
    # clipping active: C = the median norm, so samples 2^100 apart end up with comparable weight in the sum
    Cthr = n_ref.median().item()
    f = torch.zeros(Bpad, device=DEV)
    f[:B] = (Cthr / (n_ref + 1e-6)).clamp(max=1.0).float().to(DEV)
    plan.scale_backprops(f, 0, Bpad)
    out = torch.empty_like(mod.weight)
    plan.weighted_sum(out, 0, B, 148, accumulate=False, factor_row=f)
    torch.cuda.synchronize()
    ref = torch.einsum("n,n...->...", f[:B].cpu().double(), gw_ref)
    assert ((out.cpu().double() - ref).norm() / ref.norm()).item() < 1e-3
    # materialised per-sample gradients carry the exact scales back too
    gs = plan.materialize(0, B).cpu().double()
    for n in range(B):
        assert ((gs[n] - gw_ref[n]).norm() / gw_ref[n].norm()).item() < 1e-3, n


# ---------------------------------------------------------------------------------------------------------------
# thin first convolution straight from the critic's tensors (csrc/thin.cuh)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,Cin,Cout,H,W,k,s,p,d,layout", [
    (5, 3, 64, 64, 64, 5, 2, 2, 1, "nchw"),        # D64 blocks.0, image as the data loader delivers it
    (5, 3, 64, 64, 64, 5, 2, 2, 1, "nhwc"),        # ... and as a channels_last critic sees it
    (300, 3, 64, 64, 64, 5, 2, 2, 1, "nchw"),      # more samples than SMs: several items per CTA, both image buffers
    (3, 1, 32, 32, 32, 3, 1, 1, 1, "nchw"),        # one channel, stride 1, M = 32 (one chunk)
    (2, 4, 128, 32, 32, 4, 2, 1, 1, "nhwc"),       # even kernel, M = 128 (four chunks), 64 staged rows
    (2, 3, 64, 32, 32, 3, 1, 2, 2, "nchw"),        # dilation 2
])
def test_thin_capture_matches_unfold_einsum(B, Cin, Cout, H, W, k, s, p, d, layout):
    conv, A, Bp, Ho, Wo = _conv_case(B, Cin, Cout, H, W, k, s, p, d, seed=11)
    geom = L.UnfoldGeom(Cin, H, W, k, k, s, s, p, p, d, d, Ho, Wo)
    assert L.thin_direct_ok(geom, Cout)
    scale = 3.0
    a = A.to(DEV)
    if layout == "nhwc":
        a = a.contiguous(memory_format=torch.channels_last)
    bp = Bp.to(DEV).contiguous(memory_format=torch.channels_last)
    Cs = k * k * Cin
    Gs = torch.full((B, Cout * Cs), float("nan"), device=DEV)
    n2 = torch.full((B,), float("nan"), device=DEV)
    bias = torch.full((B, Cout), float("nan"), device=DEV)
    L.call("cg_thin_capture", a.data_ptr(), a.stride(0), a.stride(1), a.stride(2), a.stride(3), bp.data_ptr(), B,
           C.byref(geom), Cout, scale, Gs.data_ptr(), Gs.shape[1], n2.data_ptr(), bias.data_ptr(), st())
    torch.cuda.synchronize()
    # oracle: unfold + einsum in fp64 (reference semantics: opacus _compute_conv_grad_sample)
    U = F.unfold(A.double(), k, dilation=d, padding=p, stride=s)                    # [B][Cin*k*k][Q], rows (c, kh, kw)
    G = scale * torch.einsum("bmq,bpq->bmp", Bp.double().reshape(B, Cout, -1), U)   # [B][Cout][(c, kh, kw)]
    G = G.view(B, Cout, Cin, k, k).permute(0, 1, 3, 4, 2).reshape(B, -1)            # natural layout [m][kh][kw][c]
    err = ((Gs.cpu().double() - G).norm(dim=1) / G.norm(dim=1)).max().item()
    assert err < 1e-3, err
    n2_ref = (G * G).sum(dim=1)
    assert ((n2.cpu().double() - n2_ref).abs() / n2_ref).max().item() < 2e-3
    b_ref = scale * Bp.double().sum(dim=(2, 3))
    assert ((bias.cpu().double() - b_ref).abs().max() / b_ref.abs().max()).item() < 1e-5


def test_small_ops_table_matches_torch():
    """cg_small_ops: one launch for a table of the per-layer scalar / bias operations, each against plain torch."""
    g = torch.Generator().manual_seed(9)
    S = 300
    rows = torch.randn(S, 77, generator=g).to(DEV)
    a, b, c = (torch.rand(S, generator=g).to(DEV) + 0.1 for _ in range(3))
    out_ss = torch.full((S,), -1.0, device=DEV)
    out_cp = torch.full((S,), -1.0, device=DEV)
    out_mul = torch.full((S,), -1.0, device=DEV)
    out_col = torch.zeros(77, device=DEV)
    big_rows = torch.randn(S, 4800, generator=g).to(DEV)
    out_big = torch.zeros(4800, device=DEV)
    mult = torch.full((S,), -1.0, device=DEV)
    scale = torch.zeros(2, device=DEV)
    lo, n = 7, 250
    ops = [L.small_op(L.OP_ROW_SUMSQ, rows, out_ss, S, R=77),
           L.small_op(L.OP_COPY, a, out_cp, S),
           L.small_op(L.OP_MUL, a, out_mul, S, b=b),
           L.small_op(L.OP_WCOLSUM, rows, out_col, n, b=a, R=77, lo=lo),
           L.small_op(L.OP_WCOLSUM, big_rows, out_big, n, b=b, R=4800, lo=lo),
           L.small_op(L.OP_CLIP_MULT, a, mult, n, b=b, c=c, out2=scale, lo=lo)]
    L.small_ops(ops, st())
    torch.cuda.synchronize()
    assert torch.allclose(out_ss, (rows * rows).sum(dim=1), rtol=1e-5)
    assert torch.equal(out_cp, a)
    assert torch.equal(out_mul, a * b)
    assert torch.allclose(out_col, (a[lo:lo + n, None] * rows[lo:lo + n]).sum(dim=0), rtol=1e-4, atol=1e-4)
    assert torch.allclose(out_big, (b[lo:lo + n, None] * big_rows[lo:lo + n]).sum(dim=0), rtol=1e-4, atol=1e-3)
    prod = (a * b * c)[lo:lo + n]
    up = scale[0].item()
    assert up >= prod.max().item() and up / 2 < prod.max().item() * 1.0000001 and np.log2(up) == round(np.log2(up))
    assert torch.allclose(mult[lo:lo + n] * up, prod, rtol=1e-6)
    assert (mult[:lo] == -1).all() and (mult[lo + n:] == -1).all()


def test_thin_capture_two_batches_in_one_launch():
    """cg_thin_capture2: the fake and the real pass of a step as two segments of ONE launch (different batch sizes)."""
    conv, A, Bp, Ho, Wo = _conv_case(7, 3, 64, 64, 64, 5, 2, 2, 1, seed=21)
    conv2, A2, Bp2, _, _ = _conv_case(4, 3, 64, 64, 64, 5, 2, 2, 1, seed=22)
    geom = L.UnfoldGeom(3, 64, 64, 5, 5, 2, 2, 2, 2, 1, 1, Ho, Wo)
    Cs, M, scale = 75, 64, 2.0
    outs = []
    a1, a2 = A.to(DEV), A2.to(DEV)
    b1 = Bp.to(DEV).contiguous(memory_format=torch.channels_last)
    b2 = Bp2.to(DEV).contiguous(memory_format=torch.channels_last)
    Gs = torch.full((11, M * Cs), float("nan"), device=DEV)
    n2 = torch.full((11,), float("nan"), device=DEV)
    bias = torch.full((11, M), float("nan"), device=DEV)
    L.call("cg_thin_capture2", a1.data_ptr(), a2.data_ptr(), a1.stride(0), a1.stride(1), a1.stride(2), a1.stride(3),
           b1.data_ptr(), b2.data_ptr(), 7, 4, C.byref(geom), M, scale, Gs.data_ptr(), Gs[7:].data_ptr(), Gs.shape[1],
           n2.data_ptr(), n2[7:].data_ptr(), bias.data_ptr(), bias[7:].data_ptr(), st())
    torch.cuda.synchronize()
    for (Ai, Bi, lo) in ((A, Bp, 0), (A2, Bp2, 7)):
        nb = Ai.shape[0]
        U = F.unfold(Ai.double(), 5, padding=2, stride=2)
        G = scale * torch.einsum("bmq,bpq->bmp", Bi.double().reshape(nb, M, -1), U)
        G = G.view(nb, M, 3, 5, 5).permute(0, 1, 3, 4, 2).reshape(nb, -1)
        got = Gs[lo:lo + nb].cpu().double()
        assert ((got - G).norm(dim=1) / G.norm(dim=1)).max().item() < 1e-3
        assert ((n2[lo:lo + nb].cpu().double() - (G * G).sum(dim=1)).abs() / (G * G).sum(dim=1)).max().item() < 2e-3
        b_ref = scale * Bi.double().sum(dim=(2, 3))
        assert ((bias[lo:lo + nb].cpu().double() - b_ref).abs().max() / b_ref.abs().max()).item() < 1e-5
