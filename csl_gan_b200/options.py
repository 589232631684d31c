"""Command-line options of the DP discriminator step.

Keeps the hot-path flags of reference options.py:113-206 (names, short forms, choices, defaults and
the per-dataset default tables :11-91) and the derived fields the training step reads (:230-235,
240-256).  Every flag spelling of the reference CLI (options.py:116-206) parses, so an unmodified reference command
line is accepted; dataset / output-directory / resume / logging-cadence flags are carried but ignored
(`IGNORED_FLAGS`; SURVEY.md §8: only the D step is rebuilt), and `parse()` takes argv and touches no files.

Reference quirk kept on purpose: in `fill_defaults` a value of False counts as "unset"
(options.py:95), so e.g. CelebA cannot switch `-ispp` off.
"""
from __future__ import annotations

import argparse
import random
from argparse import Namespace
from typing import Optional, Sequence

DATASET_DEFAULTS = {
    "MNIST": dict(
        model="Vanilla", im_size=28, n_epochs=10000, g_lr=2e-4, d_lr=2e-4, batch_size=600, batch_split_size=60,
        train_set_size=60000, g_latent_dim=100, n_d_steps=1, g_label_emb_mode="concat", d_label_emb_mode="concat",
        aux_loss_type="cross_entropy", adam_b1=0.9, adam_b2=0.999, penalty=[], mean_sample_size=5000,
        mean_sample_noise_std=0.22, delta=1e-5, sigma=5.0, grad_clip_mode="standard", clipping_param=4.0,
        imm_sens_scaling_mode="standard", n_classes=10, weights_seed=42),
    "CelebA": dict(
        model="DeepConvResNet", im_size=64, n_epochs=1000, g_lr=1e-4, d_lr=1e-4, batch_size=128, batch_split_size=32,
        train_set_size=180000, public_set_size=0, g_latent_dim=128, n_d_steps=5, g_label_emb_mode="concat",
        d_label_emb_mode="concat", aux_loss_type="wasserstein", adam_b1=0.0, adam_b2=0.9, penalty=["WGAN-GP"],
        mean_sample_size=1000, mean_sample_noise_std=0.12, delta=1e-6, sigma=0.5,
        imm_sens_scaling_vec=[20, 2, 15, 1.5, 10, 1.5, 10, 1, 30], imm_sens_scaling_mode="standard",
        imm_sens_per_param=True, grad_clip_mode="standard", clipping_param=200,
        clipping_param_per_layer=[1000, 200, 1000, 100, 1000, 100, 1000, 5, 2500], n_classes=2, gp_lambda=10),
}


def str2bool(v):
    if isinstance(v, bool):
        return v
    s = v.lower()
    if s in ("yes", "true", "t", "y", "1"):
        return True
    if s in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("Boolean value expected.")


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description="DP discriminator step (csl-gan options, hot-path subset)")
    a = ap.add_argument
    a("--weights_seed", type=int, default=42)
    a("--manual_seed", type=int, default=-1)
    a("dataset", type=str, choices=["MNIST", "CelebA"])
    # dataset / output / resume plumbing of reference options.py:121-133: accepted so that an unmodified reference
    # command line parses; the D step never reads them (IGNORED_FLAGS below)
    a("-d", "--data_path", type=str, default=None)
    a("-lp", "--label_path", type=str, default=None)
    a("-la", "--label_attr", type=str, default=None)
    a("--download_mnist", default=False, action="store_true")
    a("-o", "--output_dir", type=str, default=None)
    a("-rp", "--resume_path", type=str, default=None)
    a("-re", "--resume_epochs", type=int, default=0)
    a("-ka", "--keep_args", type=str, nargs="*", default=[])
    a("--model", type=str, choices=["Vanilla", "DeepConvResNet"], default=None)
    a("--im_size", type=int, default=None, choices=[64, 48, 28])
    a("-ne", "--n_epochs", type=int, default=None)
    a("--d_lr", type=float, default=None)
    a("--g_lr", type=float, default=None)
    a("-wd", "--weight_decay", type=float, default=0)
    a("-bs", "--batch_size", type=int, default=None)
    a("-bss", "--batch_split_size", type=int, default=None)
    a("-tss", "--train_set_size", type=int, default=None)
    a("-gd", "--g_device", type=str, default="cpu")
    a("-dd", "--d_device", type=str, default="cpu")
    a("-nw", "--num_workers", type=int, default=8)
    a("--g_latent_dim", type=int, default=None)
    a("--n_d_steps", type=int, default=None)
    a("--train_d_until_threshold", type=float, default=1e10)
    a("-cond", "--conditional", action="store_true", default=False)
    a("--g_label_emb_mode", type=str, choices=["embed", "concat"], default=None)
    a("--d_label_emb_mode", type=str, choices=["embed", "concat"], default=None)
    a("--conditional_arch", type=str, choices=["CGAN", "ACGAN", "WCGAN"], default="ACGAN")
    a("--aux_loss_type", type=str, choices=["wasserstein", "cross_entropy"], default=None)
    a("--aux_loss_scalar", type=float, default=1)
    a("--aux_penalty", type=str2bool, default=True)
    a("--d_fake_aux_loss", type=str2bool, default=True)
    a("--adam_b1", type=float, default=None)
    a("--adam_b2", type=float, default=None)
    a("--penalty", type=str, nargs="*", choices=[None, "WGAN-GP", "WGAN-GP1", "DRAGAN", "DRAGAN1"], default=None)
    a("-pss", "--public_set_size", type=int, default=0)
    a("-nms", "--num_mean_samples", type=int, default=0)
    a("-pupd", "--penalty_use_public_data", type=str2bool, default=True)
    a("-wi", "--warmup_iter", type=int, default=0)
    a("--mean_sample_size", type=int, default=None)
    a("--mean_sample_noise_std", type=int, default=None)
    a("--delta", type=float, default=None)
    a("--sigma", type=float, default=None)
    a("-eb", "--epsilon_budget", type=float, default=None)
    a("-dpm", "--dp_mode", type=str, choices=["gc", "is", "tm", "sv"], default=None)
    a("-ispp", "--imm_sens_per_param", type=str2bool, default=False)
    a("-issv", "--imm_sens_scaling_vec", type=float, nargs="*", default=None)
    a("-issm", "--imm_sens_scaling_mode", type=str, choices=["standard", "constant-pl", "moving-avg-pl"], default=None)
    a("--moving_avg_beta", type=float, default=0.9,
      help="beta of the moving-avg-pl scaling update (the reference reads opt.moving_avg_beta at train.py:249 but never defines it)")
    a("-gcs", "--grad_clip_split", type=str2bool, default=True)
    a("-gcm", "--grad_clip_mode", type=str, choices=["standard", "adaptive", "constant-pl", "adaptive-pl"], default=None)
    a("-c", "--clipping_param", type=float, default=None)
    a("-cpl", "--clipping_param_per_layer", type=float, nargs="*", default=None)
    a("-as", "--adaptive_scalar", type=float, default=1.5)
    a("--adaptive_stat", choices=["mean", "max"], default="mean")
    # trimmed-mean / smooth-sensitivity knobs (reference options.py:183-188): parsed, rejected in derive() with dp_mode
    a("--smooth_sens_t", type=float, default=0.01)
    a("--tm_m", type=int, default=None)
    a("--tm_max_val", type=float, default=None)
    a("--tm_min_val", type=float, default=None)
    a("--tm_rho_per_epoch", type=float, default=10)
    a("--tm_sens_compute_bs", type=float, default=None)
    a("-bpc", "--backprop_clip", type=str2bool, default=False)
    a("--bpc_back_clip_param", type=float, default=0.01)
    a("--bpc_back_clip_param_pl", type=float, nargs="*", default=None)
    a("--bpc_forward_clip_param", type=float, default=20)
    a("--bpc_forward_clip_param_pl", type=float, nargs="*", default=None)
    a("-bpcaas", "--bpc_auto_activation_scale", type=float, default=0.2)
    a("-bpcawgs", "--bpc_auto_weight_grad_scale", type=float, default=1e-3)
    a("--bpc_during_g_train", type=str2bool, default=True)
    a("--save_every", type=int, default=None)
    a("--log_every", type=int, default=None)
    a("--sample_every", type=int, default=None)
    a("--sample_num", type=int, default=None)
    a("-p", "--profile_training", default=False, action="store_true")
    return ap


# flags of the reference CLI that the D step does not read (data loading, output directories, resume, logging
# cadence, dp_mode tm/sv knobs); they parse and are carried on the namespace untouched
IGNORED_FLAGS = ("data_path", "label_path", "label_attr", "download_mnist", "output_dir", "resume_path", "resume_epochs",
                 "keep_args", "num_workers", "train_d_until_threshold", "smooth_sens_t", "tm_m", "tm_max_val",
                 "tm_min_val", "tm_rho_per_epoch", "tm_sens_compute_bs", "save_every", "log_every", "sample_every",
                 "sample_num")


def fill_defaults(opt: Namespace, table: dict) -> None:
    for key, val in table.items():
        cur = opt.__dict__.get(key, None)
        if cur is None or cur is False:                      # reference options.py:95
            opt.__dict__[key] = val


def derive(opt: Namespace) -> Namespace:
    """Derived flags and incompatibility checks (reference options.py:230-256)."""
    opt.use_dp = opt.dp_mode is not None
    opt.use_grad_clip_per_layer = opt.grad_clip_mode not in ("standard", "adaptive")
    opt.per_sample_grad = opt.dp_mode in ("gc", "tm", "sv")
    opt.is_acgan = opt.conditional and opt.conditional_arch == "ACGAN"
    opt.use_aux_loss = opt.conditional and opt.conditional_arch in ("ACGAN", "WCGAN")
    if opt.conditional_arch == "WCGAN" and opt.aux_penalty:
        opt.aux_penalty = False
    if opt.dp_mode in ("tm", "sv"):
        raise NotImplementedError("dp_mode tm/sv are 'very experimental' in the reference (README.md:11) and out of scope")
    if opt.imm_sens_per_param and opt.imm_sens_scaling_mode not in (None, "standard"):
        raise Exception("Calculating IS per parameter does not require per parameter scaling. "
                        "Scaling estimates per-parameter calculation.")
    if opt.public_set_size > 0 and opt.num_mean_samples > 0:
        raise Exception("Both public data partition and mean samples were configured, please select only one.")
    if (len(opt.penalty) > 0 and opt.use_dp and opt.penalty_use_public_data and opt.public_set_size < 1
            and opt.num_mean_samples < 1):
        raise Exception("In order to enable gradient penalty using public data, please enable mean sampling by "
                        "setting num_mean_samples or public data by setting public_set_size.")
    if (opt.g_label_emb_mode != "concat" or opt.d_label_emb_mode != "concat") and opt.model == "Vanilla":
        raise Exception("Vanilla model with embedded labels not implemented")
    return opt


def parse(argv: Optional[Sequence[str]] = None) -> Namespace:
    opt = build_parser().parse_args(argv)
    fill_defaults(opt, DATASET_DEFAULTS[opt.dataset])
    derive(opt)
    if opt.manual_seed < 0:
        opt.manual_seed = random.randint(1, 1000000)
    return opt
