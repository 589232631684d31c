"""How fast can the critic's own forward/backward be?  (benchmark mode, channels_last, skipping wgrad)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

B = 512
dev = torch.device("cuda", 0)

def run(tag, cl, bench_mode, skip_wgrad, tf32=True):
    torch.backends.cudnn.benchmark = bench_mode
    torch.backends.cudnn.allow_tf32 = tf32
    D, real_h, fake_h, y, cfg = bench.make_workload("celeba_d64_gc", B, dev)
    r, f = real_h.to(dev), fake_h.to(dev)
    if cl:
        D = D.to(memory_format=torch.channels_last)
        r, f = r.contiguous(memory_format=torch.channels_last), f.contiguous(memory_format=torch.channels_last)
    outs = []
    h = D.blocks[0].register_forward_hook(lambda m, i, o: outs.append(o))
    def step():
        outs.clear()
        for p in D.parameters(): p.grad = None
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = D.real_loss(D(r)[0]) + D.fake_loss(D(f)[0])
        e1.record()
        if skip_wgrad:
            torch.autograd.backward(loss, inputs=list(outs))
        else:
            loss.backward()
        e2.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), e1.elapsed_time(e2)
    for _ in range(5): step()
    fw = bw = 0
    for _ in range(10):
        a, b = step(); fw += a; bw += b
    print(f"{tag:50s} fwd {fw/10:.2f} ms  bwd {bw/10:.2f} ms", flush=True)
    h.remove()

run("nchw", False, False, False)
run("nchw benchmark", False, True, False)
run("nchw benchmark skip-wgrad", False, True, True)
run("channels_last benchmark", True, True, False)
run("channels_last benchmark skip-wgrad", True, True, True)
run("nchw benchmark skip-wgrad fp32(no tf32)", False, True, True, tf32=False)
