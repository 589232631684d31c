"""Norm phase of the CelebA 16x16 layer (64 -> 128 channels, 5x5, stride 2): resident-backprop / shifted-window kernel
(CSLGAN_RESIDENT = 4 + flags) against the tap-per-box kernel (CSLGAN_RESIDENT=0) and an fp64 einsum; time per launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from csl_gan_b200.grad_sample import LayerPlan

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = "cuda"
torch.manual_seed(0)
conv = torch.nn.Conv2d(64, 128, 5, stride=2, padding=2).to(dev)
act = torch.randn(B, 64, 32, 32, device=dev).contiguous(memory_format=torch.channels_last)
bp = torch.randn(B, 128, 16, 16, device=dev).contiguous(memory_format=torch.channels_last)
plan = LayerPlan("blocks.1", conv, 0, 1)
plan.capture_activation(act, 0, B, 1)
plan.capture_backprop(bp, 0, 1.0)
n2 = torch.zeros(B, device=dev)
plan.weight_norm2(n2, 0, B)
torch.cuda.synchronize()
nb = min(B, 8)
U = F.unfold(act[:nb].double(), 5, padding=2, stride=2)
G = torch.einsum("bmq,bpq->bmp", bp[:nb].double().reshape(nb, 128, -1), U)
ref = (G * G).sum(dim=(1, 2))
err = ((n2[:nb].double() - ref).abs() / ref).max().item()
ts = []
for _ in range(5):
    n2.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.weight_norm2(n2, 0, B); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"RESIDENT={os.environ.get('CSLGAN_RESIDENT', 'default')}: max rel err of the norms {err:.2e}, {sorted(ts)[2]:.1f} us per launch (B={B})")

# ---- clipped sum of the same layer: factor-scaled backprops, ONE split-K GEMM over all samples
fac = torch.rand(B, device=dev) * 0.9 + 0.1
w = conv.weight.detach().to(memory_format=torch.channels_last)
out = torch.zeros_like(w)
props = torch.cuda.get_device_properties(0)
plan.scale_backprops(fac, 0, B)
plan.weighted_sum(out, 0, B, props.multi_processor_count, accumulate=False, factor_row=fac)
torch.cuda.synchronize()
Gs = torch.einsum("b,bmq,bpq->mp", fac[:nb].double(), bp[:nb].double().reshape(nb, 128, -1), U)
if nb == B:
    ref_w = Gs.view(128, 64, 5, 5)
    err_s = ((out.double() - ref_w).norm() / ref_w.norm()).item()
else:
    # full reference through autograd-free conv identity: sum_b f_b G_b = wgrad of (f * bp)
    ref_w = torch.nn.grad.conv2d_weight(act.double().contiguous(), w.shape, (bp.double() * fac.double().view(-1, 1, 1, 1)).contiguous(), stride=2, padding=2)
    err_s = ((out.double() - ref_w).norm() / ref_w.norm()).item()
ts = []
for _ in range(5):
    out.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.weighted_sum(out, 0, B, props.multi_processor_count, accumulate=False, factor_row=fac, prezeroed=True); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"RESIDENT={os.environ.get('CSLGAN_RESIDENT', 'default')}: clipped sum rel err {err_s:.2e}, {sorted(ts)[2]:.1f} us per launch (B={B})")
