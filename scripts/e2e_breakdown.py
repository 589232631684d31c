"""Where does the end-to-end D step spend its time?  CUDA-event brackets around each phase."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import csl_gan_b200 as cg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
D, real_h, fake_h, y, cfg = bench.make_workload("celeba_d64_gc", B, dev)
D = D.to(memory_format=torch.channels_last)
real_pin, fake_pin = real_h.pin_memory(), fake_h.pin_memory()
opt = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.0, 0.9))
eng = cg.PrivacyEngine(D, batch_size=B, sample_size=180000, noise_multiplier=0.5, max_grad_norm=cfg["C"],
                       num_private_passes=1, auto_clip_and_accum_on_step=False)
eng.attach(opt)

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def step(hooks=True, collect=None):
    marks = [("start", ev())]
    r = real_pin.to(dev, non_blocking=True); f = fake_pin.to(dev, non_blocking=True)
    marks.append(("h2d", ev()))
    for p in D.parameters(): p.grad = None
    (eng.enable_hooks() if hooks else eng.disable_hooks())
    of, _ = D(f); orr, _ = D(r)
    loss = D.real_loss(orr) + D.fake_loss(of)
    marks.append(("forward(+act staging)", ev()))
    (eng.backward(loss) if hooks else loss.backward())
    marks.append(("backward(+bp staging)", ev()))
    eng.disable_hooks()
    if hooks:
        eng.clip(); marks.append(("clip", ev()))
        eng.accum_grads_across_passes(); eng.accumulate_batch(); marks.append(("accumulate", ev()))
        opt.step(); marks.append(("noise+adam", ev()))
    l = loss.item(); marks.append(("d2h", ev()))
    torch.cuda.synchronize()
    if collect is not None:
        for (n0, e0), (n1, e1) in zip(marks, marks[1:]):
            collect[n1] = collect.get(n1, 0.0) + e0.elapsed_time(e1)

for hooks in (True, False):
    for _ in range(3): step(hooks)
    acc = {}; N = 10
    t0 = time.perf_counter()
    for _ in range(N): step(hooks, acc)
    wall = (time.perf_counter() - t0) / N * 1e3
    print(f"hooks={hooks} wall {wall:.2f} ms/step; " + ", ".join(f"{k} {v / N:.2f}" for k, v in acc.items()), flush=True)
