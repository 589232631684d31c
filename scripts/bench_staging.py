"""Per-layer timing of the capture (staging) kernels on the bench workload's shapes (CelebA D64, B per pass).
usage: python scripts/bench_staging.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import csl_gan_b200 as cg
from csl_gan_b200 import discriminators as DD

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = "cuda"
D = DD.CelebA_DCRN_D64(n_classes=0).to(dev).to(memory_format=torch.channels_last)
opt = torch.optim.SGD(D.parameters(), lr=0.0)
eng = cg.PrivacyEngine(D, batch_size=B, sample_size=100000, noise_multiplier=0.5, max_grad_norm=1.0,
                       num_private_passes=1, auto_clip_and_accum_on_step=False)
eng.attach(opt)
acts, grads = {}, {}
hooks = []
for plan in eng._plans:
    def fh(mod, inp, out, plan=plan):
        acts[plan.name] = inp[0].detach()
        out.register_hook(lambda g, plan=plan: grads.__setitem__(plan.name, g.detach()))
    hooks.append(plan.layer.register_forward_hook(fh))
x = torch.randn(B, 3, 64, 64, device=dev).contiguous(memory_format=torch.channels_last)
out = D(x, None)[0]
out.mean().backward()
for h in hooks:
    h.remove()
torch.cuda.synchronize()


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


tot = 0.0
for plan in eng._plans:
    a, g = acts[plan.name], grads[plan.name]
    plan.capture_activation(a, 0, eng.Bpad, eng.max_passes)
    ta = timeit(lambda: plan.capture_activation(a, 0, eng.Bpad, eng.max_passes))
    tg = timeit(lambda: plan.capture_backprop(g, 0, float(B)))
    impl = plan.impl
    ya = impl.Yt.numel() * 4 / eng.max_passes if impl is not None else 0
    xa = impl.Xt.numel() * 4 / eng.max_passes if impl is not None else 0
    print(f"{plan.name:12s} act {tuple(a.shape)!s:22s} {ta:7.1f} us  ({(a.numel()*4 + ya)/ta/1e6:5.2f} TB/s r+w)   "
          f"bp {tuple(g.shape)!s:22s} {tg:7.1f} us  ({(g.numel()*4 + xa)/tg/1e6:5.2f} TB/s r+w)")
    tot += ta + tg
print(f"total per pass {tot:.1f} us")
