"""Philox4x32-10 + cuRAND Box-Muller restated in numpy.  TEST INFRASTRUCTURE ONLY.

Restates the random stream behind `torch.normal(0, std, shape, generator=<cuda gen>)`,
i.e. what the fork's `_generate_noise` draws once per parameter tensor on the patched
`optimizer.step()` (reference train.py:484; upstream opacus privacy_engine.step):

  * counter-based generator ...... Philox4x32-10 (Salmon et al., Random123), as used by
                                   cuRAND `curandStatePhilox4_32_10_t`
  * stream layout ................ ATen/native/cuda/DistributionTemplates.h:50-92
                                   (installed torch headers): block 256, unroll 4,
                                   grid = min(SMs * (maxThreadsPerSM/256), ceil(numel/256)),
                                   thread `idx` -> curand_init(seed, idx, offset); element
                                   li = idx + ii * blockDim * gridDim takes component ii of
                                   the float4 drawn in that loop trip; generator offset then
                                   advances by ((numel-1)/(256*grid*4)+1)*4
  * uniform -> normal ............ cuRAND curand_normal.h:70-87 `_curand_box_muller`

The integer stream is exact (checked against the Random123 known-answer vectors in
tests/test_oracle_philox.py). The float stage uses numpy's fp32 log/sin/cos, which may
differ from the device's logf/__sincosf by a few ulp, so CPU-side checks of the normals
use a tolerance; bit-exactness of the product kernel is asserted on the GPU against
torch's own CUDA generator.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

TWO_POW32_INV = np.float32(2.3283064e-10)
TWO_POW32_INV_2PI = np.float32(np.float32(2.3283064e-10) * np.float32(6.2831855))


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr [..., 4] uint32, key [..., 2] uint32 -> [..., 4] uint32."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint64)
    k1 = key[..., 1].astype(np.uint64)
    for r in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [(hi1 ^ c[1] ^ k0) & MASK, lo1, (hi0 ^ c[3] ^ k1) & MASK, lo0]
        if r < 9:
            k0 = (k0 + np.uint64(W0)) & MASK
            k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack(c, axis=-1).astype(np.uint32)


def curand4(seed: int, subsequence: np.ndarray, offset4: np.ndarray) -> np.ndarray:
    """Output of curand4() for a state made by curand_init(seed, subsequence, 4*offset4):
    counter = (offset4 as 64 bit in x,y ; subsequence as 64 bit in z,w)."""
    subsequence = np.asarray(subsequence, dtype=np.uint64)
    offset4 = np.asarray(offset4, dtype=np.uint64)
    subsequence, offset4 = np.broadcast_arrays(subsequence, offset4)
    ctr = np.stack([offset4 & MASK, offset4 >> np.uint64(32),
                    subsequence & MASK, subsequence >> np.uint64(32)], axis=-1).astype(np.uint32)
    key = np.empty(ctr.shape[:-1] + (2,), dtype=np.uint32)
    key[..., 0] = seed & 0xFFFFFFFF
    key[..., 1] = (seed >> 32) & 0xFFFFFFFF
    return philox4x32_10(ctr, key)


def box_muller(x: np.ndarray, y: np.ndarray):
    u = x.astype(np.float32) * TWO_POW32_INV + np.float32(TWO_POW32_INV / np.float32(2))
    v = y.astype(np.float32) * TWO_POW32_INV_2PI + np.float32(TWO_POW32_INV_2PI / np.float32(2))
    s = np.sqrt(np.float32(-2.0) * np.log(u, dtype=np.float32), dtype=np.float32)
    return (np.sin(v, dtype=np.float32) * s).astype(np.float32), (np.cos(v, dtype=np.float32) * s).astype(np.float32)


def torch_cuda_grid(numel: int, sm_count: int = 148, max_threads_per_sm: int = 2048):
    block = 256
    grid = min(sm_count * (max_threads_per_sm // block), (numel + block - 1) // block)
    return block, grid


def torch_cuda_offset_increment(numel: int, sm_count: int = 148, max_threads_per_sm: int = 2048) -> int:
    block, grid = torch_cuda_grid(numel, sm_count, max_threads_per_sm)
    return ((numel - 1) // (block * grid * 4) + 1) * 4


def torch_cuda_standard_normal(numel: int, seed: int, offset: int, sm_count: int = 148,
                               max_threads_per_sm: int = 2048) -> np.ndarray:
    """The N(0,1) stream `tensor.normal_()` would consume for `numel` fp32 elements with
    generator state (seed, offset) on a device with the given SM geometry."""
    assert offset % 4 == 0
    block, grid = torch_cuda_grid(numel, sm_count, max_threads_per_sm)
    nthreads = block * grid
    out = np.empty(numel, dtype=np.float32)
    li = np.arange(numel, dtype=np.int64)
    trip = li // (nthreads * 4)
    rem = li % (nthreads * 4)
    comp = rem // nthreads
    idx = rem % nthreads
    r = curand4(seed, idx.astype(np.uint64), (np.uint64(offset // 4) + trip.astype(np.uint64)))
    n01, n23 = box_muller(r[..., 0], r[..., 1]), box_muller(r[..., 2], r[..., 3])
    comps = np.stack([n01[0], n01[1], n23[0], n23[1]], axis=-1)
    out[:] = comps[np.arange(numel), comp]
    return out
