"""Eager DP steps of the CelebA D64 workload for ncu (no CUDA graph: every kernel is its own launch).
usage: python scripts/profile_dp.py [B] [iters] [operands]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 3:
    os.environ["CSLGAN_OPERANDS"] = sys.argv[3]
import torch
import bench
import csl_gan_b200 as cg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
D, real, fake, y, cfg = bench.make_workload("celeba_d64_gc", B, dev)
D = D.to(memory_format=torch.channels_last)
eng = cg.PrivacyEngine(D, batch_size=B, sample_size=cfg["sample_size"], noise_multiplier=cfg["sigma"],
                       max_grad_norm=cfg["C"], accum_passes=False, num_private_passes=1, auto_clip_and_accum_on_step=False)
eng.disable_hooks()
eng._set_seed(1)
caps, _ = bench.grab_captures(D, real.to(dev), fake.to(dev), None)
torch.cuda.synchronize()
for _ in range(iters):
    eng.ingest_captures(caps)
    eng.clip(); eng.accum_grads_across_passes(); eng.accumulate_batch(); eng.step()
torch.cuda.synchronize()
print("ok", eng.operand_dtype)
