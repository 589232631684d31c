import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
args = argparse.Namespace(steps=2, warmup=1)
out = bench.other_configs(torch.device("cuda", 0), args, {}, only=["celeba_is_per_param_gp_bs128"])
v = out["celeba_is_per_param_gp_bs128"]
print("IS_CL", os.environ.get("CSLGAN_IS_CL", "1"), {k: round(x, 3) for k, x in v.items() if isinstance(x, float)}, v.get("graph_error"))
