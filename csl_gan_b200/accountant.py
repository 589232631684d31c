"""RDP accountant for the sampled Gaussian mechanism (API row a21).

Replaces `privacy_engine.get_privacy_spent` / `opacus.privacy_analysis.compute_rdp` as called
from reference train.py:294-295, 588, mean_sampler.py:91-92 and budget_analysis.py:79-80.
Pure CPU scalar math (Mironov, Talwar, Zhang 2019, "Renyi Differential Privacy of the Sampled
Gaussian Mechanism"); it is not on the GPU hot path but the engine API needs it.

Attribution: the log-space evaluation (`_log_add`, `_log_sub`, `_compute_log_a_int`, `_compute_log_a_frac`,
`_compute_rdp_one`) restates, from memory and with the same private function names, the public algorithm of
TensorFlow-Privacy's `rdp_accountant.py` as carried by pytorch/opacus 0.x `privacy_analysis.py` (Apache-2.0); the
fork's copy is not available here.  Pinned independently by the published epsilon table of the TensorFlow-Privacy
MNIST tutorial (tests/test_cpu_host.py::test_accountant_reproduces_the_published_tf_privacy_table).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple, Union

import numpy as np
from scipy import special


def _log_add(logx: float, logy: float) -> float:
    a, b = min(logx, logy), max(logx, logy)
    if a == -np.inf:
        return b
    return math.log1p(math.exp(a - b)) + b


def _log_sub(logx: float, logy: float) -> float:
    if logx < logy:
        raise ValueError("The result of subtraction must be non-negative.")
    if logy == -np.inf:
        return logx
    if logx == logy:
        return -np.inf
    try:
        return math.log(math.expm1(logx - logy)) + logy
    except OverflowError:
        return logx


def _log_erfc(x: float) -> float:
    return math.log(2) + special.log_ndtr(-x * 2 ** 0.5)


def _compute_log_a_int(q: float, sigma: float, alpha: int) -> float:
    log_a = -np.inf
    for i in range(alpha + 1):
        log_coef_i = math.log(special.binom(alpha, i)) + i * math.log(q) + (alpha - i) * math.log(1 - q)
        s = log_coef_i + (i * i - i) / (2 * (sigma ** 2))
        log_a = _log_add(log_a, s)
    return float(log_a)


def _compute_log_a_frac(q: float, sigma: float, alpha: float) -> float:
    log_a0, log_a1 = -np.inf, -np.inf
    i = 0
    z0 = sigma ** 2 * math.log(1 / q - 1) + 0.5
    while True:
        coef = special.binom(alpha, i)
        log_coef = math.log(abs(coef))
        j = alpha - i
        log_t0 = log_coef + i * math.log(q) + j * math.log(1 - q)
        log_t1 = log_coef + j * math.log(q) + i * math.log(1 - q)
        log_e0 = math.log(0.5) + _log_erfc((i - z0) / (math.sqrt(2) * sigma))
        log_e1 = math.log(0.5) + _log_erfc((z0 - j) / (math.sqrt(2) * sigma))
        log_s0 = log_t0 + (i * i - i) / (2 * (sigma ** 2)) + log_e0
        log_s1 = log_t1 + (j * j - j) / (2 * (sigma ** 2)) + log_e1
        if coef > 0:
            log_a0 = _log_add(log_a0, log_s0)
            log_a1 = _log_add(log_a1, log_s1)
        else:
            log_a0 = _log_sub(log_a0, log_s0)
            log_a1 = _log_sub(log_a1, log_s1)
        i += 1
        if max(log_s0, log_s1) < -30:
            break
    return _log_add(log_a0, log_a1)


def _compute_rdp_one(q: float, sigma: float, alpha: float) -> float:
    if q == 0:
        return 0.0
    if sigma == 0:
        return np.inf
    if q == 1.0:
        return alpha / (2 * sigma ** 2)
    if np.isinf(alpha):
        return np.inf
    if float(alpha).is_integer():
        return _compute_log_a_int(q, sigma, int(alpha)) / (alpha - 1)
    return _compute_log_a_frac(q, sigma, alpha) / (alpha - 1)


def compute_rdp(q: float, noise_multiplier: float, steps: Union[int, float],
                orders: Union[Sequence[float], float]) -> Union[np.ndarray, float]:
    """RDP of `steps` compositions of the sampled Gaussian mechanism at each order."""
    if isinstance(orders, (int, float)):
        return _compute_rdp_one(q, noise_multiplier, orders) * steps
    return np.array([_compute_rdp_one(q, noise_multiplier, a) for a in orders]) * steps


def get_privacy_spent(orders: Union[Sequence[float], float], rdp: Union[Sequence[float], float],
                      delta: float, conversion: str = "improved") -> Tuple[float, float]:
    """(epsilon, optimal order).  conversion="improved" is the Balle et al. 2020 bound used by
    opacus >= 0.12; "classic" is rdp - log(delta)/(alpha-1) (Mironov 2017)."""
    orders_vec = np.atleast_1d(np.asarray(orders, dtype=np.float64))
    rdp_vec = np.atleast_1d(np.asarray(rdp, dtype=np.float64))
    if len(orders_vec) != len(rdp_vec):
        raise ValueError("orders and rdp must have the same length")
    if conversion == "classic":
        eps = rdp_vec - math.log(delta) / (orders_vec - 1)
    else:
        eps = (rdp_vec - (np.log(delta) + np.log(orders_vec)) / (orders_vec - 1)
               + np.log((orders_vec - 1) / orders_vec))
    if np.isnan(eps).all():
        return np.inf, np.nan
    idx = int(np.nanargmin(eps))
    return float(eps[idx]), float(orders_vec[idx])
