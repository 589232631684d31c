#!/usr/bin/env python
"""bench.py -- per-sample clipped gradients / second for one DP discriminator step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--batch B] [--impl reference]

Metric (BASELINE.json): per-sample clipped grads/sec (DP D-step), plus the fraction of roofline.
  value   B_private_total / t_dp : the DP machinery only, captured (activation, grad_output) tensors
          already resident in HBM -> staging -> per-sample norms -> clip factors -> clipped sum ->
          (allreduce) -> Philox noise -> p.grad.  2 passes (fake + real) are contracted per sample.
  e2e     the same metric through the public API (DiscriminatorStep): pinned-host images -> H2D ->
          D forward fake+real -> backward with capture hooks -> clip -> accumulate -> noise + Adam ->
          D2H of the loss, everything inside the timed region.
  roofline  the tcgen05 contraction kernels: algorithmic FLOPs (n_passes * B * F_psg, counted once even
          though norms and the clipped sum each run the contraction) / their CUDA-event time per step,
          against the matmul peak of the operand type (MEASURED_PEAKS.json for 16-bit operands; a TF32
          matmul measured live, with its own clock sample, for TF32 operands).
  verify  outside the timed regions: the engine's clipped sums for exactly the tensors that are timed
          against the CPU oracle on the same tensors (N = 1) or against autograd's batch gradient when no
          factor clips (any N); at N > 1 also: replicas bit-identical, allreduce == sum of the local parts.
  cpu_baseline  the CPU oracle (restatement of the opacus-fork path, see oracle/dp_oracle.py) on all
          host cores over a bounded number of steps of the same workload (same batch size, 64-sample chunks).
`--impl reference` prints the CPU arm alone (the reference's arithmetic lives in an un-vendored
dependency, so the oracle port is the only runnable statement of it: kind = "port"); it runs the stated
batch size and honours --steps / --warmup (a wall-clock cap, when it bites, is declared in `config`).
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CPU_CHUNK = 64      # samples per oracle chunk: B x |theta| floats are materialised per chunk, as the reference would per batch
F_PSG = {"d64": 324_419_584, "mnist": 206_080}            # sum_layers 2*O*P*Q per (sample, pass)
CELEBA_CPL = [1000, 200, 1000, 100, 1000, 100, 1000, 5, 2500]   # reference options.py:80


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["celeba_d64_gc", "mnist_gc"], default="celeba_d64_gc")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch size (weak scaling)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads / cpu baseline")
    ap.add_argument("--no-graph", action="store_true", help="run the e2e step eagerly instead of as one CUDA graph")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------
def make_workload(name: str, B: int, device, seed: int = 0):
    from csl_gan_b200 import discriminators as DD
    torch.manual_seed(42)                                   # reference weights_seed default
    g = torch.Generator().manual_seed(1000 + seed)
    if name == "celeba_d64_gc":
        D = DD.CelebA_DCRN_D64(n_classes=0)
        real = torch.rand(B, 3, 64, 64, generator=g) * 2 - 1
        fake = torch.tanh(torch.randn(B, 3, 64, 64, generator=g))
        y = None
        cfg = dict(C=CELEBA_CPL, sigma=0.5, sample_size=180000, fpsg=F_PSG["d64"])
    else:
        D = DD.MNISTVanillaD(n_classes=10, conditional_arch="ACGAN", aux_loss_type="cross_entropy")
        real = torch.rand(B, 1, 28, 28, generator=g)
        fake = torch.sigmoid(torch.randn(B, 1, 28, 28, generator=g))
        y = torch.randint(0, 10, (B,), generator=g)
        cfg = dict(C=4.0, sigma=10.0, sample_size=60000, fpsg=F_PSG["mnist"])
    return D.to(device), real, fake, y, cfg


def d_loss(D, real, fake, y):
    of, af = D(fake, y)
    orr, ar = D(real, y)
    loss = D.real_loss(orr) + D.fake_loss(of)
    if ar is not None:
        loss = loss + D.aux_loss(ar, y) + D.aux_loss(af, y, fake=True)
    return loss


def grab_captures(D, real, fake, y):
    """Run fake+real forward/backward once with plain hooks and keep every layer's (input, grad_output), plus the
    batch gradient autograd computes in that same backward (fp32 library math: it is the independent reference of
    the no-clipping identity  sum_n G_n == B * grad(mean loss)  that verify() checks)."""
    from csl_gan_b200.privacy_engine import SUPPORTED_LAYERS
    passes, count, handles = [], {}, []
    old_tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False

    def mk(name):
        def fwd(layer, inp, out):
            k = count.get(name, 0)
            count[name] = k + 1
            while len(passes) <= k:
                passes.append({})
            passes[k][name] = [inp[0].detach().clone(), None]
            out.register_hook(lambda gr, k=k: passes[k][name].__setitem__(1, gr.detach().clone()))
        return fwd
    for name, m in D.named_modules():
        if isinstance(m, SUPPORTED_LAYERS):
            handles.append(m.register_forward_hook(mk(name)))
    d_loss(D, real, fake, y).backward()
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_tf32
    for h in handles:
        h.remove()
    batch_grads = [p.grad.detach().clone() for p in D.parameters()]
    for p in D.parameters():
        p.grad = None
    return [{n: (a, g) for n, (a, g) in d.items()} for d in passes], batch_grads


def _rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-5 * max(1.0, b.numel() ** 0.5))).item()


def oracle_sums_from_captures(D_cpu, caps_cpu, B, C, chunk=CPU_CHUNK):
    """Clipped sums of the CPU oracle from the captured tensors themselves (grad samplers -> per-sample norms ->
    clip factors -> weighted sum, oracle/dp_oracle.py), in `chunk`-sample pieces.  Returns (sums, any factor < 1)."""
    from oracle import dp_oracle as O
    layers = {n: m for n, m in D_cpu.named_modules() if isinstance(m, O.SUPPORTED)}
    params = [p for p in D_cpu.parameters() if p.requires_grad]
    pidx = {id(p): k for k, p in enumerate(params)}
    per_layer = isinstance(C, (list, tuple))
    sums = [torch.zeros_like(p) for p in params]
    any_clip = False
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        per = [[] for _ in params]
        for layer_caps in caps_cpu:
            for name, (A, G) in layer_caps.items():
                layer = layers[name]
                gw, gb = O.layer_grad_sample(layer, A[lo:hi], G[lo:hi] * B)       # loss is a batch mean -> x B
                per[pidx[id(layer.weight)]].append(gw)
                if getattr(layer, "bias", None) is not None:
                    per[pidx[id(layer.bias)]].append(gb)
        gs = [torch.stack(x, dim=0) for x in per]                                 # [n_passes, n, *shape]
        norms = O.calc_sample_norms(gs, flat=not per_layer)
        fac = O.calc_clipping_factors(norms, C, len(params))
        any_clip = any_clip or any(bool((f < 1.0).any()) for f in fac)
        for k in range(len(params)):
            sums[k] += O.weighted_sum(fac[k], gs[k]).sum(dim=0)
    return sums, any_clip


def verify(eng, D, caps, batch_grads, cfg, B, world, rank, dist_on, wl, tol=1e-3):
    """Outside the timed regions: is what gets timed also right?  """
    out = {"tol": tol}
    overlap = getattr(eng, "overlap_allreduce", False)
    eng.overlap_allreduce = False                  # first the plain route: local sums, then ONE allreduce in step()
    eng.ingest_captures(caps)
    fac = eng.clipping_factors()
    any_clip = bool((fac < 1.0).any().item())
    eng.clip()
    eng.accum_grads_across_passes()
    eng.accumulate_batch()
    local = [p.summed_grad.detach().clone() for p in D.parameters()]
    out["any_factor_clips"] = any_clip
    if not any_clip:
        # no factor clips (the CelebA defaults): the clipped sum is B x the batch gradient autograd computed
        # (cuDNN wgrad, fp32) in the very backward the captures were taken from
        out["vs_autograd_max_rel_err"] = max(_rel_err(s, g * B) for s, g in zip(local, batch_grads))
    if world == 1:
        import copy as _copy
        D_cpu = _copy.deepcopy(D).to("cpu").to(memory_format=torch.contiguous_format)
        caps_cpu = [{n: (a.cpu().contiguous(), g.cpu().contiguous()) for n, (a, g) in lc.items()} for lc in caps]
        t0 = time.perf_counter()
        ref, ref_clip = oracle_sums_from_captures(D_cpu, caps_cpu, B, cfg["C"])
        out["vs_oracle_max_rel_err"] = max(_rel_err(s, r) for s, r in zip(local, ref))
        out["vs_oracle"] = (f"oracle/dp_oracle.py on the same captured tensors, B={B} x 2 passes in {CPU_CHUNK}-sample "
                            f"chunks ({time.perf_counter() - t0:.1f} s)")
        out["oracle_any_factor_clips"] = ref_clip
    # noise-free step: p.grad must be allreduce(local sums) / global batch on every rank
    sigma = eng.noise_multiplier
    eng.noise_multiplier = 0.0
    eng.step()
    eng.noise_multiplier = sigma
    eng.steps -= 1
    got = torch.cat([p.grad.detach().reshape(-1) for p in D.parameters()])
    if dist_on:
        import torch.distributed as dist
        ref_flat = torch.cat([t.reshape(-1) for t in local])
        dist.all_reduce(ref_flat)                                             # plain NCCL sum of the local parts
        want = ref_flat / float(B * world)                                    # same (logical) element order as `got`
        out["allreduce_vs_sum_of_parts_max_rel_err"] = _rel_err(got, want)
        chk = got.view(torch.int32).to(torch.int64).sum().reshape(1)
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        out["replicas_identical"] = all(int(c.item()) == int(allc[0].item()) for c in allc)
        if overlap:
            # the timed route: clip() all-reduces finished buckets while the remaining GEMMs run; same gradients
            eng.overlap_allreduce = True
            eng.ingest_captures(caps)
            eng.clip()
            eng.accum_grads_across_passes()
            eng.accumulate_batch()
            eng.noise_multiplier = 0.0
            eng.step()
            eng.noise_multiplier = sigma
            eng.steps -= 1
            got2 = torch.cat([p.grad.detach().reshape(-1) for p in D.parameters()])
            out["overlapped_vs_single_allreduce_max_rel_err"] = _rel_err(got2, got)
    else:
        out["step_vs_local_sum_max_rel_err"] = _rel_err(got, torch.cat([t.reshape(-1) for t in local]) / float(B))
    errs = [v for k, v in out.items() if k.endswith("max_rel_err")]
    out["ok"] = bool(all(e < tol for e in errs) and out.get("replicas_identical", True))
    eng.overlap_allreduce = overlap
    for p in D.parameters():
        p.grad = None
    return out


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def timed(fn, steps, warmup, dist_on):
    """W warm-up calls, then exactly K calls between CUDA events; barrier + synchronize on both sides;
    max over ranks."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist_on:
        torch.distributed.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if dist_on:
        torch.distributed.barrier()
    ms = e0.elapsed_time(e1)
    if dist_on:
        t = torch.tensor([ms], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = t.item()
    return ms


def bandwidth_kernels(hbm_peak_gbs):
    """Achieved HBM GB/s of the bandwidth-bound kernels (noise add, per-sample row norm, l2_clip) on buffers
    larger than the 126 MB L2, against the measured copy bandwidth.  Algorithmic bytes:
    noise 8 B per parameter element, row norm 4 B per input element, l2_clip 8 B per element."""
    import ctypes as C
    from csl_gan_b200 import _lib as L
    from csl_gan_b200 import functional as FN
    out = {}

    def time_it(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3

    n = 96 * 1024 * 1024                                   # 384 MB per fp32 buffer
    a = torch.randn(n, device="cuda")
    b = torch.empty_like(a)
    inc = C.c_ulonglong(0)
    t = time_it(lambda: L.call("cg_noise_finalize", a.data_ptr(), b.data_ptr(), n, 512.0, 0.5, 512.0, 1, 0,
                               C.byref(inc), L.stream_ptr()))
    out["noise_finalize"] = {"elements": n, "GB/s": 8 * n / t / 1e9}
    rows = 8192
    x = a[: rows * 12288].view(rows, 12288)                # 8192 rows of a [B, 3*64*64] input gradient: 403 MB
    t = time_it(lambda: FN.row_l2_norm(x))
    out["row_l2_norm"] = {"shape": [rows, 12288], "GB/s": 4 * x.numel() / t / 1e9}
    t = time_it(lambda: FN.l2_clip(x, 50.0))
    out["l2_clip"] = {"shape": [rows, 12288], "GB/s": 8 * x.numel() / t / 1e9}
    for v in out.values():
        v["frac_of_measured_hbm_peak"] = v["GB/s"] / hbm_peak_gbs if hbm_peak_gbs else None
    return out


def mnist_gc_configs(dev, args, peaks):
    """BASELINE.json configs[0] (MNIST conditional vanilla GAN, dp_mode=gc, sigma=10, bs 600) measured like the main
    workload (value = DP machinery with resident captures, e2e = DiscriminatorStep from pinned host buffers), plus the
    same DP machinery at B = 65536, where the path is bandwidth work: its fraction of the measured HBM roofline with the
    algorithmic bytes of SURVEY.md 8(d) = (1050 + 139) * 4 B per (sample, pass) + 2 * |theta| * 4 B per step."""
    out = {}
    m = measure_gc("mnist_gc", 600, dev, args.steps, args.warmup, 0, 1, False, not args.no_graph, do_verify=True)
    out["mnist_gc_bs600"] = {"argv": "MNIST --conditional --dp_mode gc --sigma 10", "batch": 600,
                             "value": m["value"], "unit": "samples/s", "ms_per_step": m["t_dp"], "e2e": m["e2e"],
                             "launches_per_step": m["launches"], "launch_mode": m["launch_mode"], "verify": m["verify"],
                             "kernel_ms_per_step": m["kernel_ms_per_step"]}
    m.clear()
    Bbig = 65536
    m = measure_gc("mnist_gc", Bbig, dev, 10, 3, 0, 1, False, not args.no_graph, do_verify=False, do_e2e=False)
    bytes_step = 2 * Bbig * 4756 + 2 * 103179 * 4
    gbs = bytes_step / (m["t_dp"] * 1e-3) / 1e9
    out["mnist_gc_bs65536_dp_only"] = {"batch": Bbig, "value": m["value"], "unit": "samples/s", "ms_per_step": m["t_dp"],
                                       "algorithmic_bytes_per_step": bytes_step, "GB/s": gbs,
                                       "frac_of_measured_hbm_peak": gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                                       "kernel_ms_per_step": m["kernel_ms_per_step"]}
    m.clear()
    return out


def other_configs(dev, args, peaks, only=None):
    """The other BASELINE.json configs.  configs[0] (MNIST gc) like the main workload (mnist_gc_configs); configs[1..3]
    through the public API (DiscriminatorStep, inputs resident on the device): MNIST dp_mode=is, CelebA gc adaptive-pl
    with mean samples, CelebA is per-parameter with gradient penalty.  Reported as samples/s of the private batch;
    parity for these paths is in tests/."""
    from csl_gan_b200 import discriminators as DD
    from csl_gan_b200 import options as OPT
    from csl_gan_b200.dstep import DiscriminatorStep, setup_privacy_engine
    torch.backends.cudnn.benchmark = True
    out = {}
    if only is None:
        try:
            out.update(mnist_gc_configs(dev, args, peaks))
        except Exception as exc:                                   # never lose the main line over an extra
            out["mnist_gc_error"] = f"{type(exc).__name__}: {exc}"[:300]
            torch.cuda.synchronize()
    cases = {
        "mnist_is_bs600": (["MNIST", "--conditional", "--dp_mode", "is", "--sigma", "10"], 600),
        "celeba_gc_adaptive_pl_bs128": (["CelebA", "-nms", "32", "--dp_mode", "gc", "-gcm", "adaptive-pl"], 128),
        "celeba_is_per_param_gp_bs128": (["CelebA", "-nms", "32", "--dp_mode", "is", "-ispp", "True"], 128),
    }
    for name, (argv, B) in cases.items():
        if only is not None and name not in only:
            continue
        o = OPT.parse(argv + ["-bs", str(B), "-tss", "180000", "--manual_seed", "3"])
        ncls = o.n_classes if o.conditional else 0
        D = DD.build_discriminator(o.dataset, o.model, n_classes=ncls, im_size=o.im_size, emb_mode=o.d_label_emb_mode,
                                   conditional_arch=o.conditional_arch, aux_loss_type=o.aux_loss_type,
                                   aux_loss_scalar=o.aux_loss_scalar, weights_seed=o.weights_seed, device=dev)
        shape = (1, 28, 28) if o.dataset == "MNIST" else (3, o.im_size, o.im_size)
        if o.dataset == "CelebA" and not (o.dp_mode == "is" and os.environ.get("CSLGAN_IS_CL", "1") == "0"):
            D = D.to(memory_format=torch.channels_last)
        opt_d = torch.optim.Adam(D.parameters(), lr=o.d_lr, betas=(o.adam_b1, o.adam_b2), capturable=True)
        eng = setup_privacy_engine(o, D, opt_d)
        g = torch.Generator().manual_seed(0)
        real = torch.rand((B,) + shape, generator=g).to(dev) * 2 - 1
        fake = torch.rand((B,) + shape, generator=g).to(dev) * 2 - 1
        y = torch.randint(0, ncls, (B,), generator=g).to(dev) if ncls else None
        pub = torch.rand((B,) + shape, generator=g).to(dev) * 0.5

        def public_batch(n, labels, pub=pub, y=y):
            return pub[:n], (labels if labels is not None else y)

        step = DiscriminatorStep(o, D, opt_d, eng, public_batch=public_batch)
        for _ in range(3):
            step(real, y, fake, y, use_dp=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_it = 10
        e0.record()
        for _ in range(n_it):
            step(real, y, fake, y, use_dp=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_it
        out[name] = {"argv": " ".join(argv), "batch": B, "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3),
                     "mode": "eager DiscriminatorStep, device-resident inputs"}
        # the same step as one CUDA graph (these small-batch steps are launch-latency bound)
        try:
            from csl_gan_b200.dstep import GraphedDiscriminatorStep
            runner = GraphedDiscriminatorStep(step, (real, y, fake, y), warmup=2)
            for _ in range(3):
                runner(real, y, fake, y)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n_it):
                runner(real, y, fake, y)
            e1.record()
            torch.cuda.synchronize()
            msg = e0.elapsed_time(e1) / n_it
            out[name]["graph_ms_per_step"] = msg
            out[name]["graph_samples_per_s"] = B / (msg * 1e-3)
        except Exception as exc:                      # a step with a host read cannot be captured: report why
            out[name]["graph_error"] = f"{type(exc).__name__}: {exc}"[:200]
            torch.cuda.synchronize()
    return out


def cpu_step_rate(workload: str, B: int, steps: int, warmup: int, max_wall_s: float = 0.0):
    """Full DP D-step of the CPU oracle (fwd fake+real, backward with grad-sample hooks, norms, clip,
    weighted sum, accumulate, noise) on all host cores at the STATED batch size.  The batch goes through in
    64-sample chunks that accumulate into p.summed_grad (per-sample work is independent; the reference would
    materialise B x |theta| floats at once: 17.7 GB at B = 512), one noise draw and /B at the end.
    Returns (samples/s, ms/step, cores, timed steps actually run)."""
    from oracle import dp_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    D, real, fake, y, cfg = make_workload(workload, B, "cpu")
    gen = torch.Generator().manual_seed(1)
    chunk = min(B, CPU_CHUNK if workload == "celeba_d64_gc" else B)
    eng = O.OracleGCEngine(D, batch_size=chunk, noise_multiplier=cfg["sigma"], max_grad_norm=cfg["C"],
                           accum_passes=False, num_private_passes=1)

    def step():
        for p in D.parameters():
            p.grad = None
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            eng.batch_size = hi - lo
            eng.enable_hooks()
            d_loss(D, real[lo:hi], fake[lo:hi], None if y is None else y[lo:hi]).backward()
            eng.disable_hooks()
            eng.clip()
            eng.accum_grads_across_passes()
            eng.accumulate_batch()
        eng.step_grads(lambda k, std, shape: torch.normal(0.0, std, shape, generator=gen))
    t_w = time.perf_counter()
    for _ in range(warmup):
        step()
    per = (time.perf_counter() - t_w) / max(warmup, 1)
    if max_wall_s > 0 and warmup > 0 and per * steps > max_wall_s:
        steps = max(1, int(max_wall_s / per))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    eng.remove()
    return B / dt, dt * 1e3, cores, steps


# ----------------------------------------------------------------------------------------------
REF_MAX_WALL_S = 240.0     # declared cap on the timed region of the CPU arm


def run_reference(args, rank):
    """CPU arm: the oracle port on the host cores at the stated batch size, --steps / --warmup honoured."""
    if rank != 0:
        return
    B = args.batch or (512 if args.workload == "celeba_d64_gc" else 600)
    warm = max(1, args.warmup)
    rate, ms, cores, steps = cpu_step_rate(args.workload, B, args.steps, warm, REF_MAX_WALL_S)
    capped = steps != args.steps
    sample = (f"full DP D-step of the oracle port at B={B} per step (fake + real pass, {CPU_CHUNK}-sample chunks "
              f"accumulating into summed_grad), {steps} timed steps after {warm} warm-up")
    line = {
        "impl": "reference", "metric": "per-sample clipped grads/sec (DP D-step)", "value": rate,
        "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "per_gpu_batch": B, "n_passes": 2, "device": "cpu",
                   "max_wall_s": REF_MAX_WALL_S, "steps_capped_by_wall_clock": capped},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def measure_gc(wl, B, dev, steps, warmup, rank, world, dist_on, want_graph, do_verify, do_e2e=True):
    """One gc workload on this rank's GPU: t_dp (DP machinery, captured tensors resident in HBM; CUDA-graph replay),
    kernel attribution, verification against the oracle, and e2e through DiscriminatorStep from pinned host buffers."""
    import csl_gan_b200 as cg
    from csl_gan_b200 import _lib as L
    from csl_gan_b200.dstep import DiscriminatorStep
    from csl_gan_b200 import options as OPT

    D, real_h, fake_h, y_h, cfg = make_workload(wl, B, dev, seed=rank)
    if wl == "celeba_d64_gc":
        D = D.to(memory_format=torch.channels_last)          # cuDNN's native tensor-core layout for the critic itself
    real_pin, fake_pin = real_h.pin_memory(), fake_h.pin_memory()
    y_dev = None if y_h is None else y_h.to(dev)
    opt_d = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.0, 0.9) if wl == "celeba_d64_gc" else (0.9, 0.999),
                             capturable=True, fused=os.environ.get("CSLGAN_FUSED_ADAM", "1") == "1")
    eng = cg.PrivacyEngine(D, batch_size=B, sample_size=cfg["sample_size"], noise_multiplier=cfg["sigma"],
                           max_grad_norm=cfg["C"], accum_passes=False, num_private_passes=1,
                           auto_clip_and_accum_on_step=False, data_parallel=dist_on,
                           overlap_allreduce=dist_on and os.environ.get("CSLGAN_OVERLAP", "0") == "1")
    eng.disable_hooks()
    eng.attach(opt_d)
    eng._set_seed(1234)
    out = {"allreduce": ("bucketed, overlapped with the clipped-sum GEMMs" if eng.overlap_allreduce else
                         "one allreduce in step()") if dist_on else None,
           "clipping": "per-layer" if isinstance(cfg["C"], list) else "flat", "sigma": cfg["sigma"],
           "operand_dtype": eng.operand_dtype,
           "arithmetic": ("FP16 tensor-core operands (10-bit mantissa like TF32; exact per-sample power-of-two scale, "
                          "round-to-nearest staged), fp32 accumulation in TMEM and fp32 everywhere else"
                          if eng.operand_dtype == "f16" else
                          "TF32 tensor-core operands (round-to-nearest staged), fp32 accumulation and fp32 everywhere else")}

    # ---- value: DP machinery with captured tensors resident in HBM ---------------------------------
    caps, batch_grads = grab_captures(D, real_h.to(dev), fake_h.to(dev), y_dev)
    out["verify"] = verify(eng, D, caps, batch_grads, cfg, B, world, rank, dist_on, wl) if do_verify else None

    def dp_only():
        eng.ingest_captures(caps)
        eng.clip()
        eng.accum_grads_across_passes()
        eng.accumulate_batch()
        eng.step()

    # the DP machinery is captured once into a CUDA graph and replayed (launch-bound at MNIST sizes); with several
    # ranks the graph contains the NCCL allreduce of engine.step() (CSLGAN_GRAPH_DIST=0 keeps multi-rank runs eager)
    use_graph = want_graph and (not dist_on or os.environ.get("CSLGAN_GRAPH_DIST", "1") == "1")
    launches0 = L.launch_count
    dp_only()
    out["launches"] = L.launch_count - launches0                                # ABI launch calls per step
    if dist_on and getattr(eng, "_symm", None) is not None:
        sy = eng._symm
        out["allreduce"] = ("fused with the noise: one kernel over NVLink peer-mapped symmetric memory "
                            f"({'multicast' if sy.use_multicast else 'peer loads / stores'}), two cross-rank barriers")
    timed_fn = dp_only
    dp_graph = None
    if use_graph:
        eng.enable_graph_safe_rng()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                dp_only()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        dp_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(dp_graph):
            dp_only()
        timed_fn = dp_graph.replay
    ms_dp = timed(timed_fn, steps, warmup, dist_on)
    out["t_dp"] = ms_dp / steps
    out["value"] = B * world / (out["t_dp"] * 1e-3)
    out["launch_mode"] = "cuda-graph replay" if use_graph else "eager"

    # ---- kernel attribution: CUDA events around every ABI call, one extra pass of K steps -----------
    L.set_profile(True)
    for _ in range(steps):
        dp_only()
    torch.cuda.synchronize()
    prof = L.profile_summary()
    L.set_profile(False)
    per_step = {k: (ms / steps, n // steps) for k, (ms, n) in prof.items()}
    # the contraction kernels: channels-last MN-major GEMM (main), ghost-norm Gram kernel, kw-plane GEMM (odd
    # geometries) -- tensor bound -- and the thin first conv's capture+contraction kernel, which reads the critic's
    # fp32 tensors and is HBM bound by design (3 % of the FLOPs, a third of the operand bytes)
    contract_calls = ("cg_cl_contract", "cg_ghost_norm", "cg_contract")
    out["t_contract"] = sum(per_step.get(k, (0.0, 0))[0] for k in contract_calls)
    out["n_contract"] = sum(per_step.get(k, (0.0, 0))[1] for k in contract_calls)
    thin_calls = ("cg_thin_capture", "cg_thin_capture2")
    out["t_thin"] = sum(per_step.get(k, (0.0, 0))[0] for k in thin_calls)
    out["n_thin"] = sum(per_step.get(k, (0.0, 0))[1] for k in thin_calls)
    out["flops_step"] = 2 * B * cfg["fpsg"]
    out["kernel_ms_per_step"] = {k: round(v[0], 4) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1][0])}
    if not do_e2e:
        out["_keep"] = (dp_graph, eng)
        return out

    # ---- e2e: public API, host buffers, H2D/D2H inside the timed region -----------------------------
    argv = (["CelebA", "-dpm", "gc", "-gcm", "constant-pl", "--penalty"] if wl == "celeba_d64_gc"
            else ["MNIST", "-dpm", "gc", "--conditional", "--sigma", "10"])
    o = OPT.parse(argv + ["-bs", str(B)])
    o.penalty = []
    stepper = DiscriminatorStep(o, D, opt_d, eng)

    copy_stream = torch.cuda.Stream()
    staged = {}

    def upload():
        """H2D of one step's inputs from pinned memory on the copy stream (double-buffered: it overlaps the
        compute of the step in flight; every step still pays for its own copy inside the timed region)."""
        with torch.cuda.stream(copy_stream):
            staged["r"] = real_pin.to(dev, non_blocking=True)
            staged["f"] = fake_pin.to(dev, non_blocking=True)
            staged["ev"] = torch.cuda.Event()
            staged["ev"].record(copy_stream)

    upload()
    # one CUDA graph for the whole step (critic fwd/bwd, capture, norms, clip, allreduce, noise, Adam)
    runner = None
    if use_graph:
        from csl_gan_b200.dstep import GraphedDiscriminatorStep
        runner = GraphedDiscriminatorStep(stepper, (staged["r"], y_dev, staged["f"], y_dev), warmup=3)

    def e2e_step():
        torch.cuda.current_stream().wait_event(staged["ev"])
        r, f = staged["r"], staged["f"]
        r.record_stream(torch.cuda.current_stream())
        f.record_stream(torch.cuda.current_stream())
        if runner is not None:
            res = runner(r, y_dev, f, y_dev)                        # copies into the graph's static inputs, then replays
            upload()                                                # next step's inputs, overlapped
        else:
            upload()
            res = stepper(r, y_dev, f, y_dev, use_dp=True)
        return (res.d_real_loss + res.d_fake_loss).item()          # D2H read of the step's result

    ms_e2e = timed(e2e_step, steps, warmup, dist_on)
    out["e2e"] = {"value": B * world / (ms_e2e / steps * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e / steps,
                  "h2d_bytes_per_step": real_pin.numel() * 4 + fake_pin.numel() * 4, "d2h_bytes_per_step": 4,
                  "mode": "cuda-graph replay of DiscriminatorStep" if use_graph else "eager DiscriminatorStep",
                  "optimizer": "torch.optim.Adam(capturable=True, fused=%s)" % (os.environ.get("CSLGAN_FUSED_ADAM", "1") == "1")}
    out["_keep"] = (dp_graph, runner, eng)
    return out


def matmul_peak(dtype):
    """cuBLAS matmul peak of one operand type, measured the way MEASURED_PEAKS.json measures bf16 (8192^3, best of
    12), with the SM clock sampled while it runs."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    td = {"tf32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[dtype]
    a = torch.randn(8192, 8192, device="cuda").to(td)
    b = torch.randn(8192, 8192, device="cuda").to(td)
    best = 1e9
    with ClockSampler(torch.cuda.current_device()) as ck:
        for _ in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = old
    del a, b
    return 2 * 8192 ** 3 / (best * 1e-3) / 1e12, ck.summary()


def roofline_block(m, wl, B, peaks):
    """Contraction kernels vs the matmul peak of their operand type.  16-bit operands: the driver-written bf16 peak of
    MEASURED_PEAKS.json (fp16 and bf16 share the tensor-core rate); TF32 operands: a TF32 matmul measured live."""
    dt = m["operand_dtype"]
    live, live_clk = matmul_peak("tf32" if dt == "tf32" else "f16")
    if dt != "tf32" and peaks.get("bf16_tflops"):
        peak, src = peaks["bf16_tflops"], "MEASURED_PEAKS.json bf16_tflops (burst; driver-written)"
    else:
        peak, src = live, f"{dt} torch.matmul 8192^3 best of 12, measured live"
    t_c = m["t_contract"]
    flops = m["flops_step"]
    thin = None
    if m.get("t_thin", 0.0) > 0 and wl == "celeba_d64_gc":
        # the first conv (3 -> 64 channels, 5x5, 32x32 windows) runs in cg_thin_capture: its FLOPs leave the tensor
        # roofline, and it gets its own HBM roofline: bp 4*1024*64 + image 4*3*64*64 read, Gs 4*4800 written per unit
        thin_flops = 2 * B * (2 * 64 * 75 * 1024)
        flops -= thin_flops
        thin_bytes = 2 * B * (4 * 1024 * 64 + 4 * 3 * 64 * 64 + 4 * 4800)
        hbm = peaks.get("hbm_gbs")
        ach = thin_bytes / (m["t_thin"] * 1e-3) / 1e9
        thin = {"bound": "hbm", "kernel": "thin_direct_kernel (capture + contraction + norm of the first conv, tcgen05 kind::tf32)",
                "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": (ach / hbm) if hbm else None,
                "algorithmic_bytes_per_step": thin_bytes, "ms_per_step": m["t_thin"], "launches_per_step": m["n_thin"],
                "algorithmic_flops_per_step": thin_flops}
    achieved = flops / (t_c * 1e-3) / 1e12 if t_c > 0 else None
    traffic, tnote = None, "no tracked ncu traffic record for this workload / batch"
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if rec.get("workload") == wl and rec.get("per_gpu_batch") == B:
            traffic = rec.get("contraction_dram_bytes_per_step")
            tnote = (f"dram__bytes_read+write over the contraction launches of one step, ncu --set full at commit "
                     f"{rec.get('commit')} ({rec.get('source')}); whole DP step: {rec.get('whole_step_dram_bytes')} B; "
                     f"algorithmic operand bytes per step = {2 * B * (1032196 if wl == 'celeba_d64_gc' else 4756)}")
    except Exception:
        pass
    return {"bound": "tensor",
            "kernel": "cl_contract_kernel + cl_pair_kernel (cta_group::2) + ghost2_norm_kernel (tcgen05, TMA-fed)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
            "peak_source": src, "live_matmul_peak": {"dtype": dt if dt == "tf32" else "f16", "tflops": live, "clocks": live_clk},
            "bf16_peak_measured": peaks.get("bf16_tflops"),
            "algorithmic_flops_per_step": flops, "contract_ms_per_step": t_c,
            "thin_layer": thin,
            "contract_launches_per_step": m["n_contract"],
            "whole_dp_frac": m["flops_step"] / (m["t_dp"] * 1e-3) / 1e12 / peak,
            "traffic": traffic, "traffic_note": tnote}


_JSON_OUT = None


def emit(line: dict):
    """The one JSON line goes to the process's ORIGINAL stdout; fd 1 itself is pointed at stderr for the rest
    of the run so that library chatter (NCCL prints its version banner on stdout) cannot end up next to it."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    args = parse_args()
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the DP hot path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_on = world > 1
    if dist_on:
        torch.distributed.init_process_group("nccl", device_id=dev)

    wl = args.workload
    B = args.batch or (512 if wl == "celeba_d64_gc" else 600)
    torch.backends.cudnn.benchmark = True                     # reference train.py:28
    clk = ClockSampler(local)
    clk.__enter__()                      # sampled across both timed regions (t_dp and e2e)
    m = measure_gc(wl, B, dev, args.steps, args.warmup, rank, world, dist_on, not args.no_graph, do_verify=True)
    clk.__exit__(None, None, None)
    clocks = clk.summary()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        line = {
            "metric": "per-sample clipped grads/sec (DP D-step)", "value": m["value"], "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["t_dp"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": m["operand_dtype"],
            "data": "synthetic",
            "config": {"workload": wl, "per_gpu_batch": B, "global_batch": B * world, "n_passes": 2,
                       "contractions_per_step": 2 * B * world, "clipping": m["clipping"],
                       "sigma": m["sigma"], "parallelism": f"dp{world}", "allreduce": m["allreduce"],
                       "arithmetic": m["arithmetic"],
                       "l2": "staged operands per step exceed the 126 MB L2 (no flush needed)" if wl == "celeba_d64_gc"
                             else "working set fits in L2; MNIST is launch-latency bound (SURVEY.md §8d)"},
            "e2e": m["e2e"],
            "gpu_launches": m["launches"] * args.steps,
            "gpu_launches_per_step": m["launches"],
            "launch_mode": m["launch_mode"],
            "kernel_ms_per_step": m["kernel_ms_per_step"],
            "roofline": roofline_block(m, wl, B, peaks),
            "verify": m["verify"],
            "clocks": clocks,
        }
        if not args.no_extras and world == 1:
            line["bandwidth_kernels"] = bandwidth_kernels(peaks.get("hbm_gbs"))
            line["other_configs"] = other_configs(dev, args, peaks)
            rate, ms, cores, nst = cpu_step_rate(wl, B, 3, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
                                    "sample": f"full DP D-step of the CPU oracle at the bench batch size B={B} "
                                              f"({CPU_CHUNK}-sample chunks), {nst} timed steps after 1 warm-up "
                                              f"({ms:.0f} ms/step); compare with e2e"}
        emit(line)
    if dist_on:
        # CUDA graphs that hold NCCL kernels must be gone before the communicator is torn down; even so NCCL's
        # teardown was seen to hang after graph replays, so the ranks synchronise and leave without it
        m.clear()
        torch.cuda.synchronize()
        torch.distributed.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
