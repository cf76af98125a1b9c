"""Effective sample size used by the phase-3 L adaptation (src/training/warmup.py:442-463 ->
blackjax.diagnostics.effective_sample_size == numpyro's): FFT autocovariance + Geyer initial positive /
initial monotone sequence.  Runs on torch tensors on whatever device they live on (the [tune3, C, d] position
buffer stays on the GPU; torch.fft is cuFFT -- a once-per-run library call, not part of the hot loop)."""
from __future__ import annotations

import math

import torch


def next_fast_len(n: int) -> int:
    while True:
        m = n
        for p in (2, 3, 5):
            while m % p == 0:
                m //= p
        if m == 1:
            return n
        n += 1


def effective_sample_size(x: torch.Tensor) -> torch.Tensor:
    """x [chains, samples, dim] -> ess [dim] (all chains pooled, as blackjax does)."""
    n_chains, n = x.shape[0], x.shape[1]
    assert n > 1
    dt = x.dtype
    mean = x.mean(dim=1, keepdim=True)
    cen = x - mean
    m = next_fast_len(2 * n)
    f = torch.fft.rfft(cen, n=m, dim=1)
    acov = torch.fft.irfft(f * torch.conj(f), n=m, dim=1)[:, :n].to(dt) / n
    mean_acov = acov.mean(dim=0, keepdim=True)
    mean_var0 = mean_acov[:, :1] * n / (n - 1.0)
    weighted_var = mean_var0 * (n - 1.0) / n
    if n_chains > 1:
        weighted_var = weighted_var + mean.var(dim=0, unbiased=True, keepdim=True)
    n_even = n - n % 2
    rho = torch.cat([torch.ones_like(mean_var0), 1.0 - (mean_var0 - mean_acov[:, 1:n_even]) / weighted_var], dim=1)
    rho = rho[0]                                   # [lag, dim]
    rho_even, rho_odd = rho[0::2].clone(), rho[1::2].clone()
    T = rho_even.shape[0]
    mask0 = (rho_even + rho_odd) > 0
    mask = torch.cumprod(mask0.to(torch.int32), dim=0).bool()      # initial positive sequence
    max_t = (mask.sum(dim=0) - 1).clamp(min=0)
    rho_odd = torch.where(mask, rho_odd, torch.zeros_like(rho_odd))
    sel = max_t + 1
    cols = torch.arange(rho_even.shape[1], device=x.device)
    inb = sel < T                                                   # JAX drops out-of-bounds scatter updates
    mask_even = mask.clone()
    selc = sel.clamp(max=T - 1)
    upd = rho_even[selc, cols] > 0
    mask_even[selc[inb], cols[inb]] = upd[inb]
    rho_even = torch.where(mask_even, rho_even, torch.zeros_like(rho_even))
    rho_sum = rho_even + rho_odd
    run_min = torch.cummin(rho_sum, dim=0).values                   # initial monotone sequence
    upd_mask = rho_sum > run_min
    rho_even_f = torch.where(upd_mask, run_min / 2, rho_even)
    rho_odd_f = torch.where(upd_mask, run_min / 2, rho_odd)
    ess_raw = n_chains * n
    last = rho_even_f[selc, cols]                                   # JAX clamps out-of-bounds gathers
    tau = -1.0 + 2.0 * (rho_even_f + rho_odd_f).sum(dim=0) - last
    tau = torch.clamp(tau, min=1.0 / math.log10(ess_raw))
    return ess_raw / tau


# ------------------------------------------------------------------------------------------------------------------
# Chain diagnostics of the report (mirror of src/inference/metrics.py:226-244, 354-425, 449-523).
# They consume the [n_chains, n_samples, dim] sample tensor where it lives (the kept-sample buffer of the sampler is
# [S, C, d] in HBM: pass `samples_dev.permute(1, 0, 2)`), instead of re-reading 12 000 npz files (utils.py:131-175).
# ------------------------------------------------------------------------------------------------------------------
def _as_tensor(x) -> torch.Tensor:
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x)


def rank_normalize_array(samples: torch.Tensor) -> torch.Tensor:
    """metrics.py:226-244 applied along axis 0 of a [n, ...] tensor: overall ranks (average over ties) ->
    (r - 0.375) / (n + 0.25) -> standard normal quantile."""
    x = _as_tensor(samples)
    flat = x.reshape(x.shape[0], -1)
    n = flat.shape[0]
    order = torch.argsort(flat, dim=0, stable=True)
    ranks = torch.empty_like(flat, dtype=torch.float64)
    ar = torch.arange(1, n + 1, device=flat.device, dtype=torch.float64).unsqueeze(1).expand_as(flat)
    ranks.scatter_(0, order, ar)
    srt = torch.gather(flat, 0, order)
    if n > 1 and bool((srt[1:] == srt[:-1]).any()):       # ties: average ranks like scipy.stats.rankdata
        for j in torch.nonzero((srt[1:] == srt[:-1]).any(dim=0)).flatten().tolist():
            col = srt[:, j]
            _, inv, cnt = torch.unique_consecutive(col, return_inverse=True, return_counts=True)
            end = torch.cumsum(cnt, 0).to(torch.float64)
            avg = end - (cnt.to(torch.float64) - 1.0) / 2.0
            ranks[order[:, j], j] = avg[inv]
    p = (ranks - 0.375) / (n + 0.25)
    z = math.sqrt(2.0) * torch.erfinv(2.0 * p - 1.0)
    return z.to(x.dtype if x.dtype.is_floating_point else torch.float64).reshape(x.shape)


def _rank_normalize_pooled(x: torch.Tensor) -> torch.Tensor:
    """`jnp.apply_along_axis(rank_normalize_array, 0, x.reshape(-1, ...))`: per parameter over all chains x samples."""
    return rank_normalize_array(x.reshape(-1, *x.shape[2:])).reshape(x.shape)


def between_chain_var(x) -> torch.Tensor:
    """metrics.py:354-367."""
    return _as_tensor(x).mean(dim=1).var(dim=0, unbiased=True)


def within_chain_var(x) -> torch.Tensor:
    """metrics.py:370-383."""
    return _as_tensor(x).var(dim=1, unbiased=True).mean(dim=0)


def chain_effective_sample_size(x, rank_normalize: bool = True) -> torch.Tensor:
    """metrics.py:386-405 (`effective_sample_size` there): ESS of every chain separately -> [n_chains, dim]."""
    x = _as_tensor(x)
    if rank_normalize:
        x = _rank_normalize_pooled(x)
    return torch.stack([effective_sample_size(c[None]) for c in x])


def running_mean(x, axis: int) -> torch.Tensor:
    """metrics.py:408-425 (its count vector is shaped for axis == 1 of a 3-D array)."""
    x = _as_tensor(x)
    count = torch.arange(1, x.shape[axis] + 1, device=x.device, dtype=x.dtype)[None, :, None]
    return torch.cumsum(x, dim=axis) / count


def gelman_split_r_hat(samples, n_splits: int, rank_normalize: bool = True) -> torch.Tensor:
    """metrics.py:449-497."""
    import warnings
    x = _as_tensor(samples)
    n_chains = x.shape[0]
    n_samples = x.shape[1] / n_splits
    if n_samples % 1 != 0:
        raise ValueError('Number of samples must be divisible by n_splits')
    if n_samples < 50:
        warnings.warn('Number of samples should be at least 50x the number of splits', UserWarning)
    if rank_normalize:
        x = _rank_normalize_pooled(x)
    splits = x.reshape(n_chains * n_splits, -1, *x.shape[2:])
    wcv, bcv = within_chain_var(splits), between_chain_var(splits)
    return torch.sqrt((((n_samples - 1) / n_samples) * wcv + bcv) / wcv)


def split_chain_r_hat(samples, n_splits: int, rank_normalize: bool = True) -> torch.Tensor:
    """metrics.py:500-523: R-hat of every chain against its own splits -> [n_chains, dim]."""
    x = _as_tensor(samples)
    return torch.stack([gelman_split_r_hat(c[None], n_splits, rank_normalize) for c in x])
