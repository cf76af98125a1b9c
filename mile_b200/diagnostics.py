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
