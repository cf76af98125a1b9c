"""Host-side metrics with the reference's names (mirror of src/inference/metrics.py:247-312,428-446).  Small
numpy reductions over predictions the CUDA library produced (Ensemble.predict / the fused online state)."""
from __future__ import annotations

import math

import numpy as np


def pointwise_lppd(lvals: np.ndarray, y: np.ndarray, task: str) -> np.ndarray:
    """metrics.py:247-293: lvals [..., n_obs, K] -> log predictive density [..., n_obs]."""
    if str(getattr(task, 'value', task)).lower().startswith('regr'):
        mu = lvals[..., 0]
        sigma = np.clip(np.exp(lvals[..., 1]), 1e-6, 1e6)
        return -np.square(y - mu) / (2 * sigma * sigma) - np.log(sigma) - math.log(math.sqrt(2 * math.pi))
    m = lvals.max(axis=-1, keepdims=True)
    lse = m[..., 0] + np.log(np.exp(lvals - m).sum(axis=-1))
    yi = np.broadcast_to(np.asarray(y).astype(np.int64), lvals.shape[:-1])
    return np.take_along_axis(lvals, yi[..., None], axis=-1)[..., 0] - lse


def lppd(lppd_pointwise: np.ndarray) -> float:
    """metrics.py:296-312: mean_obs logsumexp_{chain,sample}(lp, b = 1/(C*S))."""
    flat = lppd_pointwise.reshape(-1, lppd_pointwise.shape[-1]).astype(np.float64)
    m = flat.max(axis=0)
    return float((m + np.log(np.exp(flat - m).sum(axis=0) / flat.shape[0])).mean())


def running_lppd(lppd_pointwise: np.ndarray) -> np.ndarray:
    """metrics.py:428-446."""
    e = np.exp(lppd_pointwise.astype(np.float64))
    cs = np.cumsum(e, axis=-2)
    cnt = np.arange(1, e.shape[-2] + 1).reshape((1, -1, 1))
    return np.log(cs / cnt).mean(axis=-1).mean(axis=0)


# ---- chain diagnostics with the reference's names (metrics.py:226-244, 354-425, 449-523) -------------------------
# numpy / torch in, numpy out; the arithmetic runs in torch on the GPU when one is present (mile_b200/diagnostics.py).
def _run(fn, x, *args, **kw):
    import torch
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    if not t.is_cuda and torch.cuda.is_available():
        t = t.cuda()
    return fn(t, *args, **kw).cpu().numpy()


def rank_normalize_array(samples):
    from . import diagnostics as dg
    s = np.asarray(samples)
    return _run(dg.rank_normalize_array, s.reshape(-1, 1)).reshape(s.shape)   # overall ranks of the whole array


def between_chain_var(x):
    from . import diagnostics as dg
    return _run(dg.between_chain_var, x)


def within_chain_var(x):
    from . import diagnostics as dg
    return _run(dg.within_chain_var, x)


def effective_sample_size(x, rank_normalize: bool = True):
    from . import diagnostics as dg
    return _run(dg.chain_effective_sample_size, x, rank_normalize)


def running_mean(x, axis: int):
    from . import diagnostics as dg
    return _run(dg.running_mean, x, axis)


def gelman_split_r_hat(samples, n_splits: int, rank_normalize: bool = True):
    from . import diagnostics as dg
    return _run(dg.gelman_split_r_hat, samples, n_splits, rank_normalize)


def split_chain_r_hat(samples, n_splits: int, rank_normalize: bool = True):
    from . import diagnostics as dg
    return _run(dg.split_chain_r_hat, samples, n_splits, rank_normalize)
