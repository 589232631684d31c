"""2-GPU NCCL check (skipped with fewer than 2 devices): the sharded gc step with one allreduce and
shared-seed noise equals the single-GPU step on the full batch."""
import copy
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _step(D, real, fake, B, dev, data_parallel, seed=77, sigma=2.0, C=1.5, overlap=False, fused=True):
    import csl_gan_b200 as cg
    D = copy.deepcopy(D).to(dev)
    opt = torch.optim.SGD(D.parameters(), lr=0.0)
    eng = cg.PrivacyEngine(D, batch_size=B, sample_size=60000, noise_multiplier=sigma, max_grad_norm=C,
                           num_private_passes=1, auto_clip_and_accum_on_step=False, data_parallel=data_parallel,
                           overlap_allreduce=overlap, fused_allreduce=fused)
    eng.attach(opt)
    eng._set_seed(seed)
    (D.real_loss(D(real.to(dev))[0]) + D.fake_loss(D(fake.to(dev))[0])).backward()
    eng.disable_hooks()
    eng.clip(); eng.accum_grads_across_passes(); eng.accumulate_batch()
    opt.step()
    torch.cuda.synchronize()
    return [p.grad.detach().cpu() for p in D.parameters()], eng


def _worker(rank, world, port, B, ret):
    import torch.distributed as dist
    from csl_gan_b200 import discriminators as DD
    from csl_gan_b200.dist import shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(5)
    real = torch.rand(B, 1, 28, 28, generator=g)
    fake = torch.rand(B, 1, 28, 28, generator=g) * 0.5
    torch.manual_seed(42)
    D = DD.MNIST_DCRN_D(n_classes=0)
    lo, hi = shard_range(B, rank, world)
    grads, eng = _step(D, real[lo:hi], fake[lo:hi], hi - lo, f"cuda:{rank}", True)
    # the accountant sees the GLOBAL sampling rate, whatever the shard sizes
    assert eng.global_batch_size == B and abs(eng.sample_rate - B / 60000) < 1e-12
    # every replica holds bit-identical gradients (same allreduce result, same Philox stream)
    flat = torch.cat([g.reshape(-1) for g in grads]).to(f"cuda:{rank}")
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    identical = all(torch.equal(both[0], b) for b in both[1:])
    # default route on an NVSwitch box: the exchange and the noise in ONE kernel over peer-mapped symmetric memory
    # (cg_noise_finalize_allreduce); it must give what the NCCL allreduce + noise kernel give (same Philox counters,
    # the sum formed in rank order instead of NCCL's)
    from csl_gan_b200.dist import SymmetricFlat
    if SymmetricFlat.available():
        assert eng._symm is not None, getattr(eng, "_symm_error", "symmetric memory not set up")
        grads_n, eng_n = _step(D, real[lo:hi], fake[lo:hi], hi - lo, f"cuda:{rank}", True, fused=False)
        assert eng_n._symm is None
        identical = identical and all(torch.allclose(a, b, rtol=1e-5, atol=1e-7) for a, b in zip(grads, grads_n))
    # the bucketed, overlapped allreduce (clip() reduces finished layers while the others are still contracting) gives
    # the same gradients as the single allreduce in step()
    grads_o, eng_o = _step(D, real[lo:hi], fake[lo:hi], hi - lo, f"cuda:{rank}", True, overlap=True)
    assert eng_o._reduce_buckets() is not None
    identical = identical and all(torch.allclose(a, b, rtol=1e-6, atol=1e-8) for a, b in zip(grads, grads_o))
    if rank == 0:
        full, _ = _step(D, real, fake, B, "cuda:0", False)
        err = max(((a - b).norm() / (b.norm() + 1e-12)).item() for a, b in zip(grads, full))
        ret.put((err, identical))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_sharded_step_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 15, ret)) for r in range(2)]     # 8 + 7: unequal shards
    for p in procs:
        p.start()
    err, identical = ret.get(timeout=500)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the clipped sums agree to fp32 summation order; the noise term is identical on both paths
    assert err < 1e-4, err
    assert identical
