import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_engine as T
import csl_gan_b200 as cg
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
name, B = sys.argv[1], int(sys.argv[2])
D, shape, ncls, lo = T.make(name)
real, fake, y = T.batch(shape, ncls, lo, B, seed=B)
ref = T.run_oracle(copy.deepcopy(D), real, fake, y, B, 1e9)
Dg = copy.deepcopy(D).cuda()
eng = cg.PrivacyEngine(Dg, batch_size=B, sample_size=1000, noise_multiplier=0.0, max_grad_norm=1e9, num_private_passes=1, auto_clip_and_accum_on_step=False)
T.d_loss(Dg, real.cuda(), fake.cuda(), None if y is None else y.cuda()).backward()
got = eng.per_sample_norms().cpu()
names = [n for n, _ in Dg.named_parameters()]
for k, b in enumerate(ref["per_param_norms"]):
    e = ((got[k] - b).abs() / (b.abs() + 1e-12))
    bad = (e > 1e-3).nonzero().tolist()
    print(names[k], "max relerr %.2e" % e.max().item(), "bad", bad[:6], "ref", [round(b[i[0], i[1]].item(), 6) for i in bad[:3]], "got", [round(got[k][i[0], i[1]].item(), 6) for i in bad[:3]])

# conditioning check for the bias rows: fp64 sums of the GPU's own grad_outputs vs what the engine staged
Dg2 = copy.deepcopy(D).cuda()
caps = {}
for i, blk in enumerate(Dg2.blocks):
    def fh(m, inp, out, i=i):
        out.register_hook(lambda g, i=i: caps.setdefault(i, []).append(g.detach()))
    blk.register_forward_hook(fh)
T.d_loss(Dg2, real.cuda(), fake.cuda(), None).backward()
for i in (0, 1):
    g_real = caps[i][0]                      # backward order: real pass first
    exact = (g_real.double().sum((2, 3)) * B).norm(dim=1)            # fp64 sum of the GPU tensors
    f32 = (g_real.sum((2, 3)) * B).norm(dim=1)                       # torch fp32 sum of the same tensors
    ours = got[2 * i + 1][1].double()
    cpu = ref["per_param_norms"][2 * i + 1][1].double()
    amp = (g_real.double().abs().sum((2, 3)) * B).norm(dim=1) / exact
    print(f"blocks.{i}.bias: ours vs fp64(GPU grads) {((ours - exact.cpu()).abs() / exact.cpu()).max().item():.2e}; "
          f"torch fp32 vs fp64 {((f32.double() - exact).abs() / exact).max().item():.2e}; "
          f"CPU oracle vs fp64(GPU grads) {((cpu - exact.cpu()).abs() / exact.cpu()).max().item():.2e}; cancellation x{amp.max().item():.0f}")
