"""Eager end-to-end DiscriminatorStep steps of the CelebA D64 gc workload (B per GPU = argv[1], default 512) for an ncu
launch list: critic forward/backward (cuDNN), capture, norms, clip, noise, Adam -- every kernel its own launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import csl_gan_b200 as cg
from csl_gan_b200 import options as OPT
from csl_gan_b200.dstep import DiscriminatorStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
D, real, fake, y, cfg = bench.make_workload("celeba_d64_gc", B, dev)
D = D.to(memory_format=torch.channels_last)
opt_d = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.0, 0.9), capturable=True, fused=True)
eng = cg.PrivacyEngine(D, batch_size=B, sample_size=cfg["sample_size"], noise_multiplier=cfg["sigma"],
                       max_grad_norm=cfg["C"], accum_passes=False, num_private_passes=1, auto_clip_and_accum_on_step=False)
eng.disable_hooks()
eng.attach(opt_d)
eng._set_seed(1)
o = OPT.parse(["CelebA", "-dpm", "gc", "-gcm", "constant-pl", "--penalty", "-bs", str(B)])
o.penalty = []
step = DiscriminatorStep(o, D, opt_d, eng)
r, f = real.to(dev), fake.to(dev)
for _ in range(iters):
    res = step(r, None, f, None, use_dp=True)
torch.cuda.synchronize()
print("ok")
