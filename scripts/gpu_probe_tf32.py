"""Probe: what does tcgen05 kind::tf32 do with the low 13 mantissa bits of fp32 operands (truncate / round / keep)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_kernels import contract_plain
for k in range(8, 24):
    a = 1.0 + 2.0 ** -k
    X = torch.full((128, 32), a, device="cuda")
    Y = torch.ones((128, 32), device="cuda")
    o = contract_plain(X, Y)
    print(f"a=1+2^-{k}: seen as 1+{(o[0,0].item()/32-1):.3e}  (exact {2.0**-k:.3e})")
for bits in (0x1000, 0x1001, 0x0FFF, 0x1FFF, 0x1800):
    X = torch.full((128, 32), 1.0, device="cuda").view(torch.int32).add_(bits).view(torch.float32)
    o = contract_plain(X, torch.ones((128, 32), device="cuda"))
    print(f"low bits {bits:#06x}: seen as 1+{(o[0,0].item()/32-1):.6e}   rn would be 1+{(((bits+0x1000)&~0x1FFF)*2.0**-23):.6e}")
