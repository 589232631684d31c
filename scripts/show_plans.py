import sys; sys.path.insert(0, ".")
import torch
import csl_gan_b200 as cg
from csl_gan_b200 import discriminators as DD
D = DD.CelebA_DCRN_D64(n_classes=0).cuda().to(memory_format=torch.channels_last)
eng = cg.PrivacyEngine(D, batch_size=32, sample_size=1000, noise_multiplier=0.0, max_grad_norm=1.0, auto_clip_and_accum_on_step=False)
x = torch.randn(32, 3, 64, 64, device="cuda")
D(x, None)[0].mean().backward()
for pl in eng._plans:
    im = pl.impl
    print(pl.name, "merged", im.plan.merged, "Cs", im.plan.Cs, "taps", im.plan.n_taps, "thin", im.thin, "pair", im.pair, "ghost", im.ghost)
