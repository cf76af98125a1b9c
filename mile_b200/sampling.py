"""Sampler driver (mirror of src/training/sampling.py:32-292): inference_loop / warmup_mclmc with the
reference's call signatures and on-disk side effects (warmup_params.txt, samples/{id}/sample_{n}.npz,
samples/info.pkl).  All chains of the wave run in ONE persistent CUDA kernel instead of one XLA-CPU device
per chain."""
from __future__ import annotations

import logging
import pickle
from pathlib import Path

import numpy as np

from .callbacks import SampleWriter
from .config import Sampler, SamplerConfig
from .engine import lppd_from_state
from .probabilistic import unwrap_posterior
from .types import key_to_seed, split
from .warmup import run_warmup

logger = logging.getLogger(__name__)

CHUNK = 1000  # MCLMC steps per sampling launch (rounded to a multiple of n_thinning)
NUTS_CHUNK = 500  # NUTS transitions per sampling launch
# Where the kept positions go: 'npz' = the reference layout samples/{chain}/sample_{n}.npz (default: drop-in),
# 'store' = one [C, S, d] array under <exp>/samples_store (sample_store.py; `python -m mile_b200.sample_store export`
# reproduces the npz layout afterwards), 'both'.
import os as _os
SAMPLE_FORMAT = _os.environ.get('MILE_SAMPLE_FORMAT', 'npz')


def inference_loop(unnorm_log_posterior, config: SamplerConfig, rng_key, init_params: dict, step_ids,
                   saving_path: Path, saving_path_warmup: Path | None = None, _frozen=None):
    """sampling.py:32-216: warmup -> warmup_params.txt -> n_samples steps with thinned saves -> info.pkl.
    (`_frozen`: bool [d] mask of parameters that are not sampled -- partition_sampling.partition_inference_loop.)"""
    info = {}
    step_ids = [int(s) for s in np.atleast_1d(np.asarray(step_ids))]
    n_devices = len(step_ids)
    rng_key, warmup_key, sample_key = split(rng_key, 3)                      # sampling.py:65
    assert config.warmup_steps > 0, 'Number of warmup steps must be greater than 0.'
    if config.name not in (Sampler.MCLMC, Sampler.NUTS):
        raise NotImplementedError(f'{config.name} does not have a warmup implemented.')
    nuts = config.name == Sampler.NUTS
    saving_path = Path(saving_path)
    model, x, y = unwrap_posterior(unnorm_log_posterior)
    spec = model.spec
    theta0 = spec.ravel(init_params).reshape(-1, spec.n_params)
    if theta0.shape[0] != n_devices:
        raise ValueError(f'init_params carry {theta0.shape[0]} chains but step_ids has {n_devices}')
    ens = model.make_ensemble(n_devices, x, y)
    # global chain ids key the noise streams (the reference splits one key per chain, sampling.py:181-184): waves / ranks
    # that own different step_ids never share a stream even with the same rng_key
    ens.set_option('chain_base', min(step_ids))
    if _frozen is not None:
        ens.set_frozen_mask(_frozen)
    try:
        logger.info('> Starting Warmup sampling...')
        if nuts:                                                             # sampling.py:70-81 (no warmup_params.txt)
            warmup_nuts(None, config, warmup_key, init_params, step_ids, unnorm_log_posterior, n_devices,
                        saving_path=saving_path_warmup, _ensemble=ens)
            eps = L = None
            saving_path.mkdir(parents=True, exist_ok=True)
        else:
            eps, L = warmup_mclmc(config, warmup_key, init_params, unnorm_log_posterior, n_devices, _ensemble=ens,
                                  _active=None if _frozen is None else np.flatnonzero(~np.asarray(_frozen, bool)))
            saving_path.mkdir(parents=True, exist_ok=True)
            with open(saving_path.parent / 'warmup_params.txt', 'w') as f:      # sampling.py:92-97
                f.write(','.join(str(np.float32(v)) for v in eps) + '\n')
                f.write(','.join(str(np.float32(v)) for v in L) + '\n')
        logger.info('> Warmup sampling completed successfully.')

        logger.info(f'> Starting {config.name.value} Sampling...')
        thin = int(config.n_thinning)
        chunk = max(thin, (NUTS_CHUNK if nuts else CHUNK) // thin * thin)
        nuts_info = []
        fmt = SAMPLE_FORMAT
        if fmt not in ('npz', 'store', 'both'):
            raise ValueError(f'MILE_SAMPLE_FORMAT must be npz, store or both, not {fmt!r}')
        writer = SampleWriter(spec, saving_path, step_ids) if fmt in ('npz', 'both') else None
        store = None
        if fmt in ('store', 'both'):
            from .sample_store import SampleStore
            store = SampleStore.create(saving_path.parent / 'samples_store', spec, step_ids,
                                       -(-int(config.n_samples) // thin))
        fused_lppd = ens.n_test > 0
        if fused_lppd:
            ens.lppd_reset()
        seed = key_to_seed(sample_key)
        done = 0
        while done < config.n_samples:                                       # HOT LOOP C, sampling.py:134-177
            n = min(chunk, config.n_samples - done)
            if nuts:
                samples, ninfo = ens.nuts_sample(n, step_base=done, n_thinning=thin, seed=seed, info=True, lppd=fused_lppd)
                nuts_info.append(ninfo)
            else:
                samples, _ = ens.sample(n, eps, L, step_base=done, n_thinning=thin, seed=seed, lppd=fused_lppd)
            first = -(-done // thin)
            kept = [(first + k) * thin for k in range(samples.shape[0])]
            if writer is not None:
                writer.submit(samples, kept)
            if store is not None:
                store.append(samples, kept)
            done += n
        logger.debug('sampling launches done')
        if writer is not None:
            writer.close()
        if store is not None:
            store.close()
        logger.debug('sample files closed')
        if nuts:                                                             # sampling.py:200-210: [n_devices, n_samples] each
            ni = np.concatenate(nuts_info, axis=0).transpose(1, 0, 2)
            info.update({'num_integration_steps': ni[..., 0].astype(np.int32), 'acceptance_rate': ni[..., 1],
                         'num_trajectory_expansions': ni[..., 2].astype(np.int32), 'is_divergent': ni[..., 3] > 0.5,
                         'energy': ni[..., 4], 'is_turning': ni[..., 5] > 0.5})
        if fused_lppd:
            m, s, cnt = ens.lppd_state()
            info['lppd'] = lppd_from_state(m, s, n_devices * cnt)
            np.savez_compressed(saving_path.parent / f'lppd_state_{step_ids[0]}.npz', m=m, s=s, count=cnt,
                                chains=np.asarray(step_ids))
            logger.info(f"> fused posterior-predictive LPPD over {n_devices} chains x {cnt} samples: {info['lppd']:.5f}")
        logger.info(f'> {config.name.value} Sampling completed successfully.')
    finally:
        ens.close()
    with open(saving_path / 'info.pkl', 'wb') as f:                          # sampling.py:213-215
        pickle.dump({k: v for k, v in info.items() if k != 'lppd'}, f)
    return info


def warmup_mclmc(config: SamplerConfig, rng_key, init_params: dict, unnorm_log_posterior, n_devices: int,
                 _ensemble=None, _active=None):
    """sampling.py:258-292.  Returns (step_size [n_devices], L [n_devices]); the warmed-up chain state stays
    in the ensemble (`use_warmup_as_init`) when one is passed in, else (state, parameters) like the reference."""
    model, x, y = unwrap_posterior(unnorm_log_posterior)
    spec = model.spec
    theta0 = spec.ravel(init_params).reshape(-1, spec.n_params)
    kw = dict(desired_energy_var_start=config.desired_energy_var_start,
              desired_energy_var_end=config.desired_energy_var_end, trust_in_estimate=config.trust_in_estimate,
              num_effective_samples=config.num_effective_samples, step_size_init=config.step_size_init)
    dp = bool(config.diagonal_preconditioning)
    if _ensemble is not None:
        out = run_warmup(_ensemble, theta0, rng_key, config.warmup_steps, diagonal_preconditioning=dp,
                         active=_active, **kw)
        # sampling.py:291 returns only step_size and L: the sampling phase of the reference runs WITHOUT the preconditioner
        _ensemble.set_sqrt_diag_cov(None)
        return out
    from .warmup import custom_mclmc_warmup
    res = custom_mclmc_warmup(unnorm_log_posterior, diagonal_preconditioning=dp, **kw).run(
        rng_key, init_params, config.warmup_steps)
    return res.state, {'step_size': res.parameters.step_size, 'L': res.parameters.L}


def warmup_nuts(kernel, config: SamplerConfig, rng_key, init_params: dict, step_ids, unnorm_log_posterior, n_devices: int,
                saving_path=None, _ensemble=None):
    """sampling.py:220-262: window adaptation of step size and diagonal mass matrix for every chain.  Returns
    (warmup_state, parameters) like the reference; with `_ensemble` the state and the parameters stay on the device."""
    from .nuts import custom_window_adaptation, run_window_adaptation
    if _ensemble is not None:
        model, x, y = unwrap_posterior(unnorm_log_posterior)
        theta0 = model.spec.ravel(init_params).reshape(-1, model.spec.n_params)
        eps, imm = run_window_adaptation(_ensemble, theta0, rng_key, config.warmup_steps, saving_path=saving_path,
                                         step_ids=step_ids)
        return None, {'step_size': eps, 'inverse_mass_matrix': imm}
    ids = np.repeat(np.asarray(step_ids), config.warmup_steps).reshape(n_devices, -1)        # sampling.py:255
    res = custom_window_adaptation(kernel, unnorm_log_posterior, progress_bar=True, saving_path=saving_path).run(
        rng_key, init_params, ids, config.warmup_steps, n_devices)
    return res.state, res.parameters
