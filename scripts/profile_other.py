"""Eager DiscriminatorStep of one of bench.py's other_configs for an ncu launch list.
usage: python scripts/profile_other.py <config name> [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import argparse
import torch
import bench

name = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
os.environ["CSLGAN_PROFILE_ITERS"] = str(iters)
args = argparse.Namespace(steps=2, warmup=1)
dev = torch.device("cuda", 0)
out = bench.other_configs(dev, args, {}, only=[name])
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in out[name].items() if k != "argv"})
