"""CPU-side checks: the C-ABI library loads and exports every declared symbol, host-only entry
points work without a GPU, the product path fails loudly without CUDA, accountant / options / Philox
restatement behave, and the lazy grad_sample protocol objects have the right shapes."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import csl_gan_b200 as cg
from csl_gan_b200 import _lib as L
from csl_gan_b200 import accountant, options
from oracle import philox as PH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from csl_gan_b200 import build
    build.build()                      # nvcc cross-compiles without a GPU
    return L.load()


def test_library_exports_every_symbol_declared_in_the_header(lib):
    hdr = open(os.path.join(ROOT, "include", "cslgan_b200.h")).read()
    declared = set(re.findall(r"\b(cg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/cslgan_b200.h but not exported"
    assert declared == set(L.EXPORTED_SYMBOLS), declared ^ set(L.EXPORTED_SYMBOLS)
    assert lib.cg_version() >= 100


def test_plan_unfold_host_logic(lib):
    # 5x5 stride 2 pad 2 on 64x64 (every conv of the reference critics): two row residues, taps at -1..1
    g, p = L.plan_unfold(3, 64, 64, 5, 5, 2, 2, 2, 2, 1, 1, 32, 32)
    assert (p.n_rho, p.Hs, p.Wop, p.rows, p.slot_stride, p.a_min) == (2, 34, 32, 30, 34 * 32, -1)
    assert list(p.tap_row0)[:5] == [0, 15, 0, 15, 0]
    assert list(p.tap_coloff)[:5] == [0, 0, 32, 32, 64]
    # 14x14 output: window rows padded to 16 so every TMA box starts 16-byte aligned
    g, p = L.plan_unfold(1, 28, 28, 5, 5, 2, 2, 2, 2, 1, 1, 14, 14)
    assert p.Wop == 16 and p.slot_stride == 16 * 16 and list(p.tap_coloff)[:5] == [0, 0, 16, 16, 32]
    # stride 1: a single residue, one plane row per kw
    g, p = L.plan_unfold(8, 10, 10, 3, 3, 1, 1, 1, 1, 1, 1, 10, 10)
    assert p.n_rho == 1 and p.rows == 3 * 8 and p.Hs == 12 and list(p.tap_coloff)[:3] == [0, 12, 24]
    # brute-force: every (kh, oh) maps to the input row the convolution reads
    for (H, k, s, pad, d) in [(64, 5, 2, 2, 1), (13, 3, 2, 0, 2), (10, 4, 3, 1, 1), (9, 3, 1, 1, 1)]:
        Ho = (H + 2 * pad - d * (k - 1) - 1) // s + 1
        g, p = L.plan_unfold(2, H, H, k, k, s, s, pad, pad, d, d, Ho, Ho)
        for kh in range(k):
            j = p.tap_row0[kh] // (k * 2)
            for oh in range(Ho):
                hs = p.tap_coloff[kh] // p.Wop + oh
                assert s * (hs + p.a_min) + p.rho[j] == oh * s - pad + kh * d
    with pytest.raises(L.CslGanCudaError):
        L.plan_unfold(1, 8, 8, 99, 3, 1, 1, 0, 0, 1, 1, 1, 6)


def test_product_path_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    D = cg.discriminators.MNISTVanillaD(n_classes=0)
    with pytest.raises(cg.CslGanCudaError):
        cg.PrivacyEngine(D, batch_size=4, sample_size=100, noise_multiplier=1.0, max_grad_norm=1.0)
    with pytest.raises(cg.CslGanCudaError):
        cg.ISPrivacyEngine(D, batch_size=4, sample_size=100, noise_multiplier=1.0)
    with pytest.raises(cg.CslGanCudaError):
        cg.row_l2_norm(torch.zeros(3, 4))
    with pytest.raises(cg.CslGanCudaError):
        cg.l2_clip(torch.zeros(3, 4), 1.0)


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "csl_gan_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), f"{fn} references the oracle"


def test_philox_known_answer_vectors():
    """Random123 kat_vectors for philox4x32-10."""
    def run(c, k):
        return [int(x) for x in PH.philox4x32_10(np.array(c, dtype=np.uint32), np.array(k, dtype=np.uint32))]
    assert run([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_stream_layout_and_moments():
    n = 200_000
    z = PH.torch_cuda_standard_normal(n, seed=1234, offset=0)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    # offset bookkeeping: grid = min(148*8, ceil(n/256)); increment = ((n-1)/(256*grid*4)+1)*4
    assert PH.torch_cuda_offset_increment(1) == 4
    assert PH.torch_cuda_offset_increment(256 * 1184 * 4) == 4
    assert PH.torch_cuda_offset_increment(256 * 1184 * 4 + 1) == 8
    # a later offset continues a thread's own stream: element i at offset 4 equals what the same
    # thread would draw on its second loop trip at offset 0
    a = PH.torch_cuda_standard_normal(256 * 1184 * 4 + 8, 7, 0)
    b = PH.torch_cuda_standard_normal(8, 7, 4, sm_count=148)
    # thread idx t, second trip -> linear index 256*1184*4 + t (component 0)
    np.testing.assert_array_equal(a[256 * 1184 * 4: 256 * 1184 * 4 + 8], b[:8])


def test_accountant_matches_closed_forms_and_is_monotone():
    # q = 1: RDP of the plain Gaussian mechanism is alpha / (2 sigma^2)
    assert accountant.compute_rdp(1.0, 2.0, 1, 8.0) == pytest.approx(8.0 / (2 * 4.0))
    orders = [1 + x / 10.0 for x in range(1, 100)] + list(range(12, 64))
    r1 = accountant.compute_rdp(0.01, 1.1, 100, orders)
    r2 = accountant.compute_rdp(0.01, 1.1, 200, orders)
    np.testing.assert_allclose(r2, 2 * r1)
    e1, a1 = accountant.get_privacy_spent(orders, r1, 1e-5)
    e2, _ = accountant.get_privacy_spent(orders, r2, 1e-5)
    assert 0 < e1 < e2 and a1 in orders
    e_hi, _ = accountant.get_privacy_spent(orders, accountant.compute_rdp(0.01, 0.6, 100, orders), 1e-5)
    assert e_hi > e1                       # less noise -> more privacy loss
    ec, _ = accountant.get_privacy_spent(orders, r1, 1e-5, conversion="classic")
    assert ec >= e1                        # the improved conversion is never looser
    # integer and fractional orders agree at the boundary
    lo = accountant.compute_rdp(0.02, 1.3, 1, 5.0)
    hi = accountant.compute_rdp(0.02, 1.3, 1, 5.0 + 1e-9)
    assert hi == pytest.approx(lo, rel=1e-5)
    # the well-known DP-SGD reference point (q=256/60000, sigma=1.1, 60 epochs, delta=1e-5) is eps ~ 3.0 (classic ~3.5)
    steps = int(60 * 60000 / 256)
    big = list(orders) + list(range(64, 256))
    eps, _ = accountant.get_privacy_spent(big, accountant.compute_rdp(256 / 60000, 1.1, steps, big), 1e-5)
    assert 2.5 < eps < 3.6


def test_accountant_reproduces_the_published_tf_privacy_table():
    """Independent pin of compute_rdp: the (noise multiplier, epochs) -> epsilon rows of the TensorFlow-Privacy MNIST
    DP-SGD tutorial (n = 60000, batch 256, delta = 1e-5; its RDP accountant is the Mironov-Talwar-Zhang 2019 sampled
    Gaussian bound with the classic conversion over the same default orders): 1.19, 3.01, 7.10."""
    orders = [1 + x / 10.0 for x in range(1, 100)] + list(range(12, 64))
    q = 256 / 60000
    for sigma, epochs, eps_pub in [(1.3, 15, 1.19), (1.1, 60, 3.01), (0.7, 45, 7.10)]:
        rdp = accountant.compute_rdp(q, sigma, epochs * 60000 // 256, orders)
        eps, _ = accountant.get_privacy_spent(orders, rdp, 1e-5, conversion="classic")
        assert round(eps, 2) == eps_pub, (sigma, epochs, eps)
        eps_i, _ = accountant.get_privacy_spent(orders, rdp, 1e-5)
        assert eps_i < eps                       # the improved (Balle et al. 2020) conversion is tighter


def test_options_accept_every_flag_of_the_reference_cli():
    """An unmodified reference command line parses (reference options.py:116-206 defines 109 flag spellings); the I/O,
    resume, logging-cadence and tm/sv flags are carried on the namespace and ignored by the D step."""
    o = options.parse(["CelebA", "-d", "/data/celeba", "-lp", "/data/list_attr.txt", "-la", "Male", "-o", "out", "-rp", "ckpt",
                       "-re", "3", "-ka", "sigma", "d_lr", "-nw", "4", "--train_d_until_threshold", "0.5", "--download_mnist",
                       "--smooth_sens_t", "0.02", "--tm_m", "3", "--tm_max_val", "1", "--tm_min_val", "-1",
                       "--tm_rho_per_epoch", "5", "--tm_sens_compute_bs", "64", "--save_every", "1", "--log_every", "100",
                       "--sample_every", "1000", "--sample_num", "16", "-p", "-dpm", "gc", "-nms", "32", "-gd", "cuda:0",
                       "-dd", "cuda:0", "-bss", "32", "-wd", "0.0", "-wi", "10", "-eb", "8.0"])
    assert o.data_path == "/data/celeba" and o.keep_args == ["sigma", "d_lr"] and o.log_every == 100 and o.tm_m == 3
    assert set(options.IGNORED_FLAGS) <= set(vars(o))
    have = {s for a in options.build_parser()._actions for s in (a.option_strings or [a.dest])}
    ref_flags = """--weights_seed --manual_seed dataset -d --data_path -lp --label_path -la --label_attr --model --im_size
        --download_mnist -o --output_dir -rp --resume_path -re --resume_epochs -ka --keep_args -ne --n_epochs --d_lr --g_lr
        -wd --weight_decay -bs --batch_size -bss --batch_split_size -tss --train_set_size -gd --g_device -dd --d_device
        -nw --num_workers --g_latent_dim --n_d_steps --train_d_until_threshold -cond --conditional --g_label_emb_mode
        --d_label_emb_mode --conditional_arch --aux_loss_type --aux_loss_scalar --aux_penalty --d_fake_aux_loss --adam_b1
        --adam_b2 --penalty -pss --public_set_size -nms --num_mean_samples -pupd --penalty_use_public_data -wi --warmup_iter
        --mean_sample_size --mean_sample_noise_std --delta --sigma -eb --epsilon_budget -dpm --dp_mode -ispp
        --imm_sens_per_param -issv --imm_sens_scaling_vec -issm --imm_sens_scaling_mode -gcs --grad_clip_split -gcm
        --grad_clip_mode -c --clipping_param -cpl --clipping_param_per_layer -as --adaptive_scalar --adaptive_stat
        --smooth_sens_t --tm_m --tm_max_val --tm_min_val --tm_rho_per_epoch --tm_sens_compute_bs -bpc --backprop_clip
        --bpc_back_clip_param --bpc_back_clip_param_pl --bpc_forward_clip_param --bpc_forward_clip_param_pl -bpcaas
        --bpc_auto_activation_scale -bpcawgs --bpc_auto_weight_grad_scale --bpc_during_g_train --save_every --log_every
        --sample_every --sample_num -p --profile_training""".split()
    assert len(ref_flags) == 109 and not [f for f in ref_flags if f not in have]


def test_options_defaults_and_derived_flags():
    o = options.parse(["MNIST", "-dpm", "gc", "--conditional", "--sigma", "10"])
    assert (o.batch_size, o.clipping_param, o.sigma, o.grad_clip_mode, o.grad_clip_split) == (600, 4.0, 10.0, "standard", True)
    assert o.use_dp and o.per_sample_grad and o.is_acgan and o.use_aux_loss and not o.use_grad_clip_per_layer
    o = options.parse(["CelebA", "-dpm", "gc", "-gcm", "adaptive-pl", "-nms", "32"])
    assert o.use_grad_clip_per_layer and o.clipping_param_per_layer == [1000, 200, 1000, 100, 1000, 100, 1000, 5, 2500]
    assert o.penalty == ["WGAN-GP"] and o.adaptive_scalar == 1.5 and o.batch_size == 128
    o = options.parse(["CelebA", "-dpm", "is", "-ispp", "False", "-nms", "32"])
    assert o.imm_sens_per_param is True        # reference quirk: False counts as unset (options.py:95)
    assert not o.per_sample_grad
    with pytest.raises(Exception):
        options.parse(["CelebA", "-dpm", "gc"])                    # penalty on public data needs -nms / -pss
    with pytest.raises(NotImplementedError):
        options.parse(["MNIST", "-dpm", "tm"])
    with pytest.raises(Exception):
        options.parse(["MNIST", "-dpm", "is", "-ispp", "True", "-issm", "constant-pl"])


def test_rel_lazy_grad_sample_shapes_do_not_need_a_gpu():
    class FakeEngine:
        _cur_B = 7
        _params = [torch.zeros(3, 4)]
        def _n_passes_view(self):
            return 2
    v = cg.GradSampleView(FakeEngine(), 0)
    assert v.size(1) == 7 and tuple(v.shape) == (2, 7, 3, 4) and len(v) == 2 and v.dim() == 4


def test_backprop_clipper_bounds_bookkeeping():
    """grad_l2_bounds as reference backprop_clip.py:71-96 derives them (shapes from a dry forward)."""
    import math
    from csl_gan_b200.backprop_clip import BackpropClipper
    torch.manual_seed(0)
    D = cg.discriminators.MNISTVanillaD(n_classes=0)
    # explicit parameters: Linear weight bound = C_in * C_back, bias bound = C_back
    bc = BackpropClipper(D, [0.01, 0.02], [20, 5], wrap=False)
    assert bc.grad_l2_bounds == pytest.approx([20 * 0.01, 0.01, 5 * 0.02, 0.02])
    # automatic parameters (reference :62-80)
    bc = BackpropClipper(D, None, None, auto_activation_scale=0.2, auto_weight_grad_scale=1e-3, wrap=False)
    c_in1 = math.sqrt(784 * 0.2 ** 2)
    w1 = math.sqrt(128 * 784 * 1e-3 ** 2)
    assert bc.input_clip_params[0] == pytest.approx(c_in1)
    assert bc.grad_l2_bounds[0] == pytest.approx(w1) and bc.grad_l2_bounds[1] == pytest.approx(w1 / c_in1)
    # conv critic: weight bound = C_in * sqrt(out pixels) * C_back, bias bound = C_back * out pixels
    # (the reference takes prod(out_shape[1:]) of a batch-stripped shape, i.e. H*W, backprop_clip.py:79-93)
    Dc = cg.discriminators.MNIST_DCRN_D(n_classes=0)
    bc = BackpropClipper(Dc, [0.01] * 3, [20] * 3, wrap=False)
    n1 = 14 * 14
    assert bc.grad_l2_bounds[0] == pytest.approx(20 * math.sqrt(n1) * 0.01) and bc.grad_l2_bounds[1] == pytest.approx(0.01 * n1)
    assert len(bc.grad_l2_bounds) == 5


def test_device_mean_sampler_shapes_and_privacy_cost():
    from csl_gan_b200.mean_sampler import DeviceMeanSampler
    g = torch.Generator().manual_seed(0)
    ms = DeviceMeanSampler(torch.rand(2, 5, 3, 8, 8, generator=g), noise_std=0.12, mean_size=1000, dataset_size=180000,
                           smallest_class_size=70000, generator=g)
    x, y = ms.sample(12)
    assert tuple(x.shape) == (12, 3, 8, 8) and tuple(y.shape) == (12,)
    lab = torch.tensor([1, 0, 1])
    x2, y2 = ms.sample(3, noise_std=0, noise_mean_std=0, requested_labels=lab)
    assert torch.equal(y2, lab)
    for i in range(3):        # with the noise off a sample IS one of the stored class means
        assert any(torch.equal(x2[i], ms.mean_samples[lab[i], k]) for k in range(5))
    eps, alpha = ms.get_privacy_cost(1e-6)
    assert eps > 0
    one = DeviceMeanSampler(torch.rand(1, 4, 1, 8, 8), 0.2, 100, 1000)
    assert one.sample(9)[1] is None


def test_split_k_group_count_keeps_the_last_wave_full():
    """cl_plan._pick_split_k: items = tiles x groups are dealt to one persistent CTA (or CTA pair) per SM (pair)."""
    from csl_gan_b200.cl_plan import _pick_split_k
    for n_tiles, units, ctas in ((100, 512, 148), (26, 2048, 148), (7, 8192, 148), (13, 2048, 74), (50, 512, 74),
                                 (1, 32, 148), (3, 5, 148), (1000, 64, 148)):
        g = _pick_split_k(n_tiles, units, ctas)
        assert 1 <= g <= max(1, units // 8)                      # at least 8 k-blocks per group
        items = n_tiles * g
        waves = -(-items // ctas)
        eff = items / (waves * ctas)
        # never worse than the old rule (about two items per CTA) by more than the per-group penalty
        g_old = max(1, min((2 * ctas) // n_tiles, units // 4 if units >= 4 else 1))
        items_old = n_tiles * g_old
        eff_old = items_old / (-(-items_old // ctas) * ctas)
        assert eff >= eff_old - 0.004 * 64, (n_tiles, units, ctas, g, eff, eff_old)
    assert _pick_split_k(100, 512, 148) >= 4                     # the 4x4 layer of the CelebA critic: 1.35 waves before


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the CPU arm of the bench contract): ONE JSON line on stdout, whatever libraries
    print; the keys the driver reads are present.  MNIST workload so the CPU suite stays fast."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "mnist_gc",
                           "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, proc.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"] == "mnist_gc" and d["gpu_launches"] == 0


def test_layer_geometry_matches_torch_output_shapes():
    """ClLayerPlan.geometry (host logic): window grids of Conv2d and the swapped roles of ConvTranspose2d."""
    import torch
    from torch import nn
    from csl_gan_b200.cl_plan import ClLayerPlan
    for conv, x in ((nn.Conv2d(3, 8, 5, stride=2, padding=2), torch.zeros(2, 3, 64, 64)),
                    (nn.Conv2d(4, 6, 3, stride=1, padding=0, dilation=2), torch.zeros(2, 4, 13, 12)),
                    (nn.Conv2d(5, 7, (3, 5), stride=(1, 2), padding=(1, 2)), torch.zeros(1, 5, 9, 11))):
        y = conv(x)
        Cn, H, W, kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo, M = ClLayerPlan.geometry(conv, "conv", x.shape)
        assert (Cn, H, W, M) == (x.shape[1], x.shape[2], x.shape[3], conv.out_channels)
        assert (Ho, Wo) == tuple(y.shape[2:]) and (kh, kw) == tuple(conv.kernel_size)
    for ct, x in ((nn.ConvTranspose2d(6, 4, 4, stride=2, padding=1), torch.zeros(2, 6, 8, 8)),
                  (nn.ConvTranspose2d(3, 5, 3, stride=2, padding=0, output_padding=1), torch.zeros(1, 3, 5, 7))):
        y = ct(x)
        # the per-sample gradient of a transposed conv is the conv contraction with the roles swapped: the
        # "unfolded" operand is the layer OUTPUT's gradient [out_channels, H, W], the plain operand the input
        Cn, H, W, kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo, M = ClLayerPlan.geometry(ct, "convT", x.shape)
        assert (Cn, H, W) == tuple(y.shape[1:]) and (Ho, Wo) == tuple(x.shape[2:]) and M == ct.in_channels


def test_thin_direct_geometry_check_is_pure_host_logic(lib):
    """cg_thin_direct_ok (csrc/thin.cuh coverage) needs no GPU: the CelebA first conv is covered, a window grid that
    is not a multiple of 64 positions, too many staged rows or an odd number of backprop chunks are not."""
    def geom(C, H, k, s, p, d=1):
        Ho = (H + 2 * p - d * (k - 1) - 1) // s + 1
        return L.UnfoldGeom(C, H, H, k, k, s, s, p, p, d, d, Ho, Ho)
    assert L.thin_direct_ok(geom(3, 64, 5, 2, 2), 64)           # CelebA D64 blocks.0
    assert L.thin_direct_ok(geom(1, 32, 3, 1, 1), 32)
    assert not L.thin_direct_ok(geom(3, 40, 5, 2, 2), 64)       # 20 x 20 = 400 positions: not a multiple of 64
    assert not L.thin_direct_ok(geom(8, 64, 5, 2, 2), 64)       # 200 staged rows > 128
    assert not L.thin_direct_ok(geom(3, 64, 5, 2, 2), 96)       # three 32-channel chunks
    assert not L.thin_direct_ok(geom(3, 64, 5, 2, 2), 60)


def test_batched_entry_points_validate_their_tables_without_a_gpu(lib):
    """cg_small_ops / cg_scale_slots_h_multi reject over-long or malformed tables before touching the device."""
    with pytest.raises(L.CslGanCudaError):
        L.call("cg_small_ops", (L.SmallOp * 33)(), 33, None)
    with pytest.raises(L.CslGanCudaError):
        L.call("cg_small_ops", None, 3, None)
    arr = (L.SmallOp * 1)()
    arr[0].op, arr[0].n = 99, 5
    with pytest.raises(L.CslGanCudaError):
        L.call("cg_small_ops", arr, 1, None)                      # null operands / unknown op
    with pytest.raises(L.CslGanCudaError):
        L.call("cg_scale_slots_h_multi", (L.ScaleSeg * 9)(), 9, None)
    assert L.call("cg_small_ops", arr, 0, None) is None or True   # empty table: nothing to do


def test_symmetric_flat_is_unavailable_without_an_nccl_group():
    """The fused exchange needs an initialised NCCL group with at least two ranks; everything else keeps NCCL / no
    exchange at all (single process)."""
    from csl_gan_b200.dist import SymmetricFlat
    assert SymmetricFlat.available() is False
