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
  roofline  the tcgen05 contraction kernel: algorithmic FLOPs (n_passes * B * F_psg, counted once even
          though norms and the clipped sum each run the contraction) / its CUDA-event time per step,
          against the TF32 matmul peak measured live the way MEASURED_PEAKS.json measures bf16.
  cpu_baseline  the CPU oracle (restatement of the opacus-fork path, see oracle/dp_oracle.py) on all
          host cores over a bounded sample of the same workload.
`--impl reference` prints the CPU arm alone (the reference's arithmetic lives in an un-vendored
dependency, so the oracle port is the only runnable statement of it: kind = "port").
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
    """Run fake+real forward/backward once with plain hooks and keep every layer's (input, grad_output)."""
    from csl_gan_b200.privacy_engine import SUPPORTED_LAYERS
    passes, count, handles = [], {}, []

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
    for h in handles:
        h.remove()
    for p in D.parameters():
        p.grad = None
    return [{n: (a, g) for n, (a, g) in d.items()} for d in passes]


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


def measure_tf32_peak():
    """TF32 matmul peak, measured the way MEASURED_PEAKS.json measures bf16 (8192^3, best of 10)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(8192, 8192, device="cuda")
    b = torch.randn(8192, 8192, device="cuda")
    best = 1e9
    for _ in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = old
    del a, b
    return 2 * 8192 ** 3 / (best * 1e-3) / 1e12


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


def other_configs(dev, only=None):
    """BASELINE.json configs[1..3] through the public API (DiscriminatorStep, eager, inputs resident on the
    device): MNIST dp_mode=is, CelebA gc adaptive-pl with mean samples, CelebA is per-parameter with gradient
    penalty.  Reported as samples/s of the private batch; parity for these paths is in tests/."""
    from csl_gan_b200 import discriminators as DD
    from csl_gan_b200 import options as OPT
    from csl_gan_b200.dstep import DiscriminatorStep, setup_privacy_engine
    torch.backends.cudnn.benchmark = True
    out = {}
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
        if o.dataset == "CelebA":
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


def cpu_step_rate(workload: str, B: int, steps: int, warmup: int):
    """Full DP D-step of the CPU oracle (fwd fake+real, backward with grad-sample hooks, norms, clip,
    weighted sum, accumulate, noise) on all host cores; returns (samples/s, ms/step, cores)."""
    from oracle import dp_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    D, real, fake, y, cfg = make_workload(workload, B, "cpu")
    gen = torch.Generator().manual_seed(1)
    eng = O.OracleGCEngine(D, batch_size=B, noise_multiplier=cfg["sigma"], max_grad_norm=cfg["C"],
                           accum_passes=False, num_private_passes=1)

    def step():
        for p in D.parameters():
            p.grad = None
        eng.enable_hooks()
        d_loss(D, real, fake, y).backward()
        eng.disable_hooks()
        eng.clip()
        eng.accum_grads_across_passes()
        eng.accumulate_batch()
        eng.step_grads(lambda k, std, shape: torch.normal(0.0, std, shape, generator=gen))
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    eng.remove()
    return B / dt, dt * 1e3, cores


# ----------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """CPU arm: the oracle port on the host cores, bounded sample of the same workload per step."""
    if rank != 0:
        return
    B_full = args.batch or (512 if args.workload == "celeba_d64_gc" else 600)
    B = min(B_full, 64 if args.workload == "celeba_d64_gc" else 600)
    steps = max(1, min(args.steps, 5 if args.workload == "celeba_d64_gc" else 20))
    warm = max(1, min(args.warmup, 1 if args.workload == "celeba_d64_gc" else 3))
    rate, ms, cores = cpu_step_rate(args.workload, B, steps, warm)
    sample = f"full DP D-step of the oracle port, B={B} of {B_full}, {steps} timed steps after {warm} warm-up"
    line = {
        "impl": "reference", "metric": "per-sample clipped grads/sec (DP D-step)", "value": rate,
        "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "per_gpu_batch": B_full, "n_passes": 2, "device": "cpu"},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


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

    import csl_gan_b200 as cg
    from csl_gan_b200 import _lib as L
    from csl_gan_b200.dstep import DiscriminatorStep
    from csl_gan_b200 import options as OPT

    wl = args.workload
    B = args.batch or (512 if wl == "celeba_d64_gc" else 600)
    torch.backends.cudnn.benchmark = True                     # reference train.py:28
    D, real_h, fake_h, y_h, cfg = make_workload(wl, B, dev, seed=rank)
    if wl == "celeba_d64_gc":
        D = D.to(memory_format=torch.channels_last)          # cuDNN's native tensor-core layout for the critic itself
    real_pin, fake_pin = real_h.pin_memory(), fake_h.pin_memory()
    y_dev = None if y_h is None else y_h.to(dev)
    opt_d = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.0, 0.9) if wl == "celeba_d64_gc" else (0.9, 0.999),
                             capturable=True, fused=os.environ.get("CSLGAN_FUSED_ADAM", "1") == "1")
    eng = cg.PrivacyEngine(D, batch_size=B, sample_size=cfg["sample_size"], noise_multiplier=cfg["sigma"],
                           max_grad_norm=cfg["C"], accum_passes=False, num_private_passes=1,
                           auto_clip_and_accum_on_step=False, data_parallel=dist_on)
    eng.disable_hooks()
    eng.attach(opt_d)
    eng._set_seed(1234)

    # ---- value: DP machinery with captured tensors resident in HBM ---------------------------------
    caps = grab_captures(D, real_h.to(dev), fake_h.to(dev), y_dev)

    def dp_only():
        eng.ingest_captures(caps)
        eng.clip()
        eng.accum_grads_across_passes()
        eng.accumulate_batch()
        eng.step()

    # the DP machinery is captured once into a CUDA graph and replayed (launch-bound at MNIST sizes); with several
    # ranks the graph contains the NCCL allreduce of engine.step() (CSLGAN_GRAPH_DIST=0 keeps multi-rank runs eager)
    use_graph = not args.no_graph and (not dist_on or os.environ.get("CSLGAN_GRAPH_DIST", "1") == "1")
    launches0 = L.launch_count
    dp_only()
    launches = L.launch_count - launches0                                       # ABI launch calls per step
    timed_fn = dp_only
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
    clk = ClockSampler(local)
    clk.__enter__()                      # sampled across both timed regions (t_dp below, e2e further down)
    ms_dp = timed(timed_fn, args.steps, args.warmup, dist_on)
    t_dp = ms_dp / args.steps
    value = B * world / (t_dp * 1e-3)

    # ---- kernel attribution: CUDA events around every ABI call, one extra pass of K steps -----------
    L.set_profile(True)
    for _ in range(args.steps):
        dp_only()
    torch.cuda.synchronize()
    prof = L.profile_summary()
    L.set_profile(False)
    per_step = {k: (ms / args.steps, n // args.steps) for k, (ms, n) in prof.items()}
    # the contraction kernels: channels-last MN-major GEMM (main), ghost-norm Gram kernel, kw-plane GEMM (odd geometries)
    contract_calls = ("cg_cl_contract", "cg_ghost_norm", "cg_contract")
    t_contract = sum(per_step.get(k, (0.0, 0))[0] for k in contract_calls)
    n_contract = sum(per_step.get(k, (0.0, 0))[1] for k in contract_calls)
    flops_step = 2 * B * cfg["fpsg"]

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

    ms_e2e = timed(e2e_step, args.steps, args.warmup, dist_on)
    e2e_value = B * world / (ms_e2e / args.steps * 1e-3)
    h2d = real_pin.numel() * 4 + fake_pin.numel() * 4
    clk.__exit__(None, None, None)
    clocks = clk.summary()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf32_peak = measure_tf32_peak()
        achieved = flops_step / (t_contract * 1e-3) / 1e12 if t_contract > 0 else None
        roof = {"bound": "tensor", "kernel": "cl_contract_kernel + cl_pair_kernel (cta_group::2) + ghost2_norm_kernel (tcgen05 kind::tf32, TMA-fed)", "achieved": achieved,
                "peak": tf32_peak, "unit": "TFLOP/s", "frac": (achieved / tf32_peak) if achieved else None,
                "peak_source": "TF32 torch.matmul 8192^3 best of 12, measured live (MEASURED_PEAKS.json has no TF32 figure)",
                "bf16_peak_measured": peaks.get("bf16_tflops"),
                "frac_of_bf16_peak": (achieved / peaks["bf16_tflops"]) if achieved and peaks.get("bf16_tflops") else None,
                "algorithmic_flops_per_step": flops_step, "contract_ms_per_step": t_contract,
                "contract_launches_per_step": n_contract,
                "whole_dp_frac": flops_step / (t_dp * 1e-3) / 1e12 / tf32_peak,
                # dram__bytes_read+write summed over the contraction launches of one step, from the ncu --set full
                # capture in profiles/r1_ncu_full_contraction_kernels_final.txt (B=512/GPU CelebA workload only)
                "traffic": 2.57e9 if (wl == "celeba_d64_gc" and B == 512) else None,
                "traffic_note": "per step (8 contraction launches, profiles/r1_ncu_full_contraction_kernels_final.txt); algorithmic operand bytes per step = "
                                f"{2 * B * (1032196 if wl == 'celeba_d64_gc' else 4756)}"}
        line = {
            "metric": "per-sample clipped grads/sec (DP D-step)", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dp,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32",
            "data": "synthetic",
            "config": {"workload": wl, "per_gpu_batch": B, "global_batch": B * world, "n_passes": 2,
                       "contractions_per_step": 2 * B * world, "clipping": "per-layer" if isinstance(cfg["C"], list) else "flat",
                       "sigma": cfg["sigma"], "parallelism": f"dp{world}",
                       "arithmetic": "TF32 tensor-core operands (round-to-nearest staged), fp32 accumulation and fp32 everywhere else",
                       "l2": "staged operands per step exceed the 126 MB L2 (no flush needed)" if wl == "celeba_d64_gc"
                             else "working set fits in L2; MNIST is launch-latency bound (SURVEY.md §8d)"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "mode": "cuda-graph replay of DiscriminatorStep" if use_graph else "eager DiscriminatorStep",
                    "optimizer": "torch.optim.Adam(capturable=True, fused=%s)" % (os.environ.get("CSLGAN_FUSED_ADAM", "1") == "1")},
            "gpu_launches": launches * args.steps,
            "gpu_launches_per_step": launches,
            "launch_mode": "cuda-graph replay" if use_graph else "eager",
            "kernel_ms_per_step": {k: round(v[0], 4) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1][0])},
            "roofline": roof,
            "clocks": clocks,
        }
        if not args.no_extras and world == 1:
            line["bandwidth_kernels"] = bandwidth_kernels(peaks.get("hbm_gbs"))
            line["other_configs"] = other_configs(dev)
            Bc = 64 if wl == "celeba_d64_gc" else 600
            rate, ms, cores = cpu_step_rate(wl, Bc, 3, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
                                    "sample": f"full DP D-step of the CPU oracle, B={Bc}, 3 timed steps after 1 warm-up "
                                              f"({ms:.0f} ms/step); compare with e2e"}
        emit(line)
    if dist_on:
        # CUDA graphs that hold NCCL kernels must be gone before the communicator is torn down; even so NCCL's
        # teardown was seen to hang after graph replays, so the ranks synchronise and leave without it
        runner = dp_graph = None
        torch.cuda.synchronize()
        torch.distributed.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        if use_graph:
            os._exit(0)
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
