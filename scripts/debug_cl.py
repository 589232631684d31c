import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from csl_gan_b200 import _lib as L
from csl_gan_b200.cl_plan import ClLayerPlan
dev = "cuda"
def tf32(x):
    i = x.contiguous().view(torch.int32); return ((i + 0x1000) & ~0x1FFF).view(torch.float32)

# (a) plain GEMM through the CL kernel: Linear geometry, Q=1, k-block = 32 slots
def chunked(t):
    """[rows, ch] -> [ceil(ch/32), rows, 32] zero padded"""
    rows, ch = t.shape
    n = (ch + 31) // 32
    out = torch.zeros(n, rows, 32, device=t.device)
    for i in range(n):
        w = min(32, ch - 32 * i)
        out[i, :, :w] = t[:, 32 * i:32 * i + w]
    return out.contiguous()

def gemm(M, P, S, n_groups=1):
    g = torch.Generator().manual_seed(M + P)
    X = tf32(torch.randn(S, M, generator=g).to(dev))
    Y = tf32(torch.randn(S, P, generator=g).to(dev))
    Xt, Yt = chunked(X), chunked(Y)
    geom = L.UnfoldGeom(P, 1, 1, 1, 1, 1, 1, 0, 0, 1, 1, 1, 1)
    plan = L.plan_cl(geom, False)
    d = L.ClDesc()
    d.Xt, d.xt_pitch, d.xt_rows, d.M = Xt.data_ptr(), 32, S, M
    d.Yt, d.n_slots_total = Yt.data_ptr(), S
    d.group_mode, d.n_groups, d.slot_lo, d.slot_hi = L.GROUP_SPLITK, n_groups, 0, S
    out = torch.zeros(M, P, device=dev)
    d.epi, d.out, d.out_group_stride, d.max_ctas = L.EPI_ACCUM, out.data_ptr(), 0, 0
    L.call("cg_cl_contract", C.byref(d), C.byref(geom), C.byref(plan), L.stream_ptr())
    torch.cuda.synchronize()
    ref = (X.double().t() @ Y.double()).float()
    print(f"gemm M={M} P={P} S={S} groups={n_groups}: relerr {((out-ref).abs().max()/ref.abs().max()).item():.3e}", flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
if which == "gemm":
    gemm(128, 32, 32)
    gemm(128, 256, 64)
    gemm(64, 100, 96, 2)
    gemm(300, 700, 640, 3)
    gemm(10, 8192, 1024, 4)
