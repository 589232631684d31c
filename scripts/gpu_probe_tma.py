"""Probe: does a 2-D SWIZZLE_128B TMA box tolerate a start column that is not 16-byte aligned?"""
import ctypes as C
import subprocess
import sys

import torch

def run(off):
    from csl_gan_b200 import _lib as L
    M, N, K = 128, 128, 64
    g = torch.Generator().manual_seed(1)
    X = torch.randn(M, K, generator=g).cuda()
    Yfull = torch.randn(N, K + 64, generator=g).cuda()
    i = X.view(torch.int32); X = ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    i = Yfull.view(torch.int32); Yfull = ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    d = L.ContractDesc()
    d.X, d.x_pitch, d.x_rows, d.x_cols = X.data_ptr(), X.stride(0), M, K
    d.Y, d.y_pitch, d.y_rows, d.y_cols = Yfull.data_ptr(), Yfull.stride(0), N, K + 64
    d.M, d.C, d.KH, d.KW = M, N, 1, 1
    d.tap_coloff[0] = off
    d.nkb = 1
    d.x_slot_stride = d.y_slot_stride = 32
    d.group_mode, d.n_groups = L.GROUP_SPLITK, 1
    d.slot_lo, d.slot_hi, d.spg = 0, K // 32, K // 32
    d.n_seg, d.seg_stride = 1, 1
    out = torch.zeros(M, N, device="cuda")
    d.epi, d.out = L.EPI_ACCUM, out.data_ptr()
    L.call("cg_contract", C.byref(d), L.stream_ptr())
    torch.cuda.synchronize()
    ref = (X.double() @ Yfull[:, off:off + K].double().t()).float()
    print("off", off, "relerr", ((out - ref).abs().max() / ref.abs().max()).item(), flush=True)

if len(sys.argv) > 1:
    run(int(sys.argv[1]))
else:
    for off in (0, 32, 4, 8, 2, 1, 14):
        r = subprocess.run([sys.executable, __file__, str(off)], capture_output=True, text=True, timeout=120)
        print((r.stdout + r.stderr).strip().splitlines()[-1] if (r.stdout + r.stderr).strip() else "no output", flush=True)
