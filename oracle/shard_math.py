"""TEST INFRASTRUCTURE ONLY - numpy restatement of the engine's sample-shard
reduction (csrc/mppi_kernels.cuh: block / shard partials and combine_partials).

A shard that owns samples [lo, hi) reduces them to
    xmax = max_k x_k,   S = sum_k exp(x_k - xmax),   N[t,d] = sum_k exp(x_k - xmax) u_k[t,d]
with x_k = -c_k / lambda (fp32, like torch.softmax(-costs / lambda) of
src/pi_mpc/mppi.py:376), and any number of such partials combine exactly like
one global softmax:  opt = sum_p N_p e^{xmax_p - xmax} / sum_p S_p e^{xmax_p - xmax}.
"""
from __future__ import annotations

import numpy as np


def shard_partial(costs: np.ndarray, perturbed: np.ndarray, lam: float) -> np.ndarray:
    """costs [k], perturbed [k,T,du] -> flat partial [2 + T*du] (fp32 header, fp64 math)."""
    x = (-costs.astype(np.float32)) / np.float32(lam)
    xmax = x.max()
    e = np.exp((x - xmax).astype(np.float32)).astype(np.float64)
    n = np.tensordot(e, perturbed.astype(np.float64), axes=(0, 0)).reshape(-1)
    return np.concatenate([[xmax, e.sum()], n]).astype(np.float64)


def combine_partials(parts: np.ndarray) -> np.ndarray:
    """parts [n, 2 + E] -> weighted mean control sequence, flat [E]."""
    xmax = parts[:, 0].max()
    scale = np.exp((parts[:, 0] - xmax).astype(np.float32)).astype(np.float64)
    S = (parts[:, 1] * scale).sum()
    return (parts[:, 2:] * scale[:, None]).sum(axis=0) / S
