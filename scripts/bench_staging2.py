"""Capture-kernel microbenchmark on the CelebA D64 tensors (B=512, channels_last): us and GB/s (fp32 read + staged write)
per call for the FP16 path, fused (single pass) vs two-pass (CSLGAN_FUSED_STAGE=0 in a second process)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from csl_gan_b200.grad_sample import LayerPlan

B = 512
dev = "cuda"
layers = [("blocks.0", 3, 64, 64), ("blocks.1", 64, 128, 32), ("blocks.2", 128, 256, 16), ("blocks.3", 256, 512, 8)]
print("route", os.environ.get("CSLGAN_STAGE", "sweep"), "per", os.environ.get("CSLGAN_SWEEP_PER", "2048"))
for name, cin, cout, h in layers:
    conv = torch.nn.Conv2d(cin, cout, 5, stride=2, padding=2).to(dev)
    act = torch.randn(B, cin, h, h, device=dev)
    if cin > 3:
        act = act.contiguous(memory_format=torch.channels_last)      # (the image batch arrives NCHW)
    bp = torch.randn(B, cout, h // 2, h // 2, device=dev).contiguous(memory_format=torch.channels_last)
    plan = LayerPlan(name, conv, 0, 1)
    plan.capture_activation(act, 0, B, 2)
    plan.capture_backprop(bp, 0, float(B))
    for what, fn, t in (("Y", lambda: plan.capture_activation(act, 0, B, 2), act), ("X", lambda: plan.capture_backprop(bp, 0, float(B)), bp)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        big = torch.empty(64 * 1024 * 1024, device=dev)
        ts = []
        for _ in range(5):
            big.zero_()                      # flush L2
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = sorted(ts)[len(ts) // 2]
        staged = plan.impl.Yt if what == "Y" else plan.impl.Xt
        out_bytes = staged.numel() * staged.element_size() / 2          # one of two passes
        print(f"{name} {what}: {us:7.1f} us  in {t.numel()*4/1e6:6.1f} MB out {out_bytes/1e6:6.1f} MB  "
              f"{(t.numel()*4 + out_bytes)/us/1e3:6.0f} GB/s  half={plan.impl.half} merged={plan.impl.plan.merged}")
