"""MCLMC warmup (mirror of src/training/warmup.py:155-568: custom_mclmc_warmup, mclmc_find_L_and_step_size,
make_L_step_size_adaptation, make_adaptation_L, handle_nans).  Phases 1+2 (HOT LOOP A) run inside the
persistent CUDA kernel (mile_mclmc_tune); phase 3 (HOT LOOP B) is the sampling kernel capturing every
position into an HBM buffer, followed by the effective sample size of every series on the device (csrc/mile_ess.cuh;
`phase3_L` keeps the torch.fft form for callers that hold the positions themselves)."""
from __future__ import annotations

import logging

import numpy as np

from .engine import Ensemble
from .probabilistic import unwrap_posterior
from .types import AdaptationAlgorithm, AdaptationResults, IntegratorState, MCLMCAdaptationState, key_to_seed, split

logger = logging.getLogger(__name__)

PHASE_RATIO = (0.8, 0.1, 0.1)      # warmup.py:543
LFACTOR = 0.4                      # warmup.py:224
CHUNK = 10000                      # MCLMC steps per kernel launch


def run_warmup(ens: Ensemble, theta0: np.ndarray, rng_key, num_steps: int, *, desired_energy_var_start,
               desired_energy_var_end, trust_in_estimate, num_effective_samples, step_size_init,
               fft_params_limit: int = 2000, fft_samples_limit: int = 10000, diagonal_preconditioning: bool = False,
               active=None):
    """custom_mclmc_warmup(...).run for all chains of `ens` at once (no torch on this path).  Returns (step_size [C], L [C]); with
    `diagonal_preconditioning` the preconditioner stays set on `ens` (read it with `ens.get_sqrt_diag_cov()`, clear it
    with `ens.set_sqrt_diag_cov(None)` -- the reference's sampling phase does not use it, sampling.py:291)."""
    seed = key_to_seed(rng_key)
    tune1, tune2, tune3 = (int(num_steps * r) for r in PHASE_RATIO)        # warmup.py:555-557
    # blackjax.mcmc.mclmc.init with the SAME key as the tuning (warmup.py:539-541,552)
    ens.init(theta0, seed=seed)
    part1_key, part2_key = split(seed, 2)                                  # warmup.py:210
    ens.tune_reset(step_size_init)                                         # warmup.py:204-209
    cfg = ens.tune_cfg(tune1, tune2, desired_energy_var_start, desired_energy_var_end, trust_in_estimate,
                       num_effective_samples)
    done = 0
    while done < tune1 + tune2:                                            # HOT LOOP A
        n = min(CHUNK, tune1 + tune2 - done)
        ens.tune(n, done, cfg, seed=part1_key)
        done += n
    logger.debug('warmup phases 1+2 done (%d steps)', tune1 + tune2)
    if tune2 != 0:
        ens.tune_finish_phase2()                                           # L = sqrt(sum var), warmup.py:387-390
        if diagonal_preconditioning:                                       # warmup.py:391-401
            eps, _, _ = ens.get_tuning()
            ens.precondition_from_moments()                                # sqrt_diag_cov = sqrt(var), L = sqrt(d)
            _, Ld, _ = ens.get_tuning()
            # "readjust the stepsize": tune2 // 3 steps of phase-1 style adaptation (mask = 1) with the preconditioned
            # kernel; run_steps restarts the adaptive state (0, 0, inf), the streaming averages and the step counter
            steps = tune2 // 3
            ens.tune_reset(0.0)
            ens.set_tuning(step_size=eps, L=Ld)
            final_key = split(part1_key, 2)[1]
            if steps > 0:
                ens.tune(steps, 0, cfg, seed=final_key)
    eps, L, _ = ens.get_tuning()
    if tune3 != 0:                                                         # HOT LOOP B, warmup.py:408-465
        # every position stays in HBM; the effective sample size of every (chain, parameter) series is computed there
        # (csrc/mile_ess.cuh) with the reference's subsampling rules (warmup.py:442-456)
        sel = np.arange(ens.d, dtype=np.int32) if active is None else np.asarray(active, dtype=np.int32)
        pidx = None if active is None else sel
        if sel.size > fft_params_limit:
            rng = np.random.default_rng(int(part2_key) & ((1 << 63) - 1))
            pidx = np.sort(rng.permutation(sel)[:fft_params_limit]).astype(np.int32)
        sidx = None
        if tune3 > fft_samples_limit:
            sidx = np.linspace(0, tune3 - 1, fft_samples_limit).astype(np.int32)
        ess = ens.phase3_ess(tune3, eps, L, seed=part2_key, param_idx=pidx, sample_idx=sidx)     # [C, n_selected]
        logger.debug('warmup phase 3 done (%d steps, ESS of %d series)', tune3, ess.size)
        L = (LFACTOR * eps.astype(np.float64) * np.mean(float(tune3) / ess.astype(np.float64), axis=1)).astype(np.float32)
        ens.set_tuning(L=L)
    return eps.astype(np.float32), L.astype(np.float32)


def phase3_L(positions, step_size, seed: int = 0, fft_params_limit: int = 2000, fft_samples_limit: int = 10000,
             device: int | None = None) -> np.ndarray:
    """make_adaptation_L's epilogue (warmup.py:442-463) for all chains: positions [tune3, C, d] (torch CUDA tensor
    or numpy) -> L [C] = 0.4 * eps_c * mean_i(tune3 / ESS_ci), with the reference's subsampling of parameters
    (> 2000, random permutation) and of samples (> 10000, linspace)."""
    import torch
    from .diagnostics import effective_sample_size
    if not torch.is_tensor(positions):
        positions = torch.from_numpy(np.ascontiguousarray(positions, dtype=np.float32)).to(
            f'cuda:{0 if device is None else device}')
    dev = positions.device
    tune3, C, d = positions.shape
    flat = positions
    if d > fft_params_limit:                                           # warmup.py:442-449
        g = torch.Generator(device='cpu').manual_seed(int(seed) & ((1 << 63) - 1))
        perm = torch.randperm(d, generator=g)[:fft_params_limit].to(dev)
        flat = flat[:, :, perm]
    if tune3 > fft_samples_limit:                                      # warmup.py:450-456
        idx = torch.linspace(0, tune3 - 1, fft_samples_limit).to(torch.int64).to(dev)
        flat = flat[idx]
    eps = np.broadcast_to(np.asarray(step_size, np.float32), (C,))
    Ls = []
    for c in range(C):
        ess = effective_sample_size(flat[:, c][None])                  # [1, samples, dim]
        Ls.append(LFACTOR * float(eps[c]) * float(torch.mean(float(tune3) / ess)))
    return np.asarray(Ls, np.float32)


def custom_mclmc_warmup(logdensity_fn, diagonal_preconditioning: bool = True, desired_energy_var_start: float = 5e-4,
                        desired_energy_var_end: float = 5e-4, trust_in_estimate: float = 1.5,
                        num_effective_samples: int = 100, step_size_init: float = 0.005) -> AdaptationAlgorithm:
    """warmup.py:486-568 (same argument names and defaults)."""
    model, x, y = unwrap_posterior(logdensity_fn)
    spec = model.spec

    def run(rng_key, position, num_steps: int = 1000):
        theta0 = spec.ravel(position)
        batched = theta0.ndim == 2
        theta0 = theta0.reshape(-1, spec.n_params)
        ens = model.make_ensemble(theta0.shape[0], x, y)
        try:
            eps, L = run_warmup(ens, theta0, rng_key, num_steps, desired_energy_var_start=desired_energy_var_start,
                                desired_energy_var_end=desired_energy_var_end, trust_in_estimate=trust_in_estimate,
                                num_effective_samples=num_effective_samples, step_size_init=step_size_init,
                                diagonal_preconditioning=diagonal_preconditioning)
            th, u, lp, g = ens.get_state()
            sdc = ens.get_sqrt_diag_cov()
        finally:
            ens.close()
        un = (lambda a: spec.unravel(a)) if batched else (lambda a: spec.unravel(a[0]))
        state = IntegratorState(un(th), un(u), lp if batched else lp[0], un(g))
        params = MCLMCAdaptationState(L if batched else L[0], eps if batched else eps[0], sdc if batched else sdc[0])
        return AdaptationResults(state, params)

    return AdaptationAlgorithm(run)
