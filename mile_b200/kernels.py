"""Kernel registry (mirror of src/training/kernels/__init__.py:1-20): KERNELS['mclmc'] returns a
blackjax-style SamplingAlgorithm(init, step) whose arithmetic runs in libmile_b200.so.

`sampler.step` advances ONE step per call (one kernel launch): it exists for API parity and single-step
tests.  `inference_loop` drives the persistent kernel in large chunks instead."""
from __future__ import annotations

from typing import Callable

import numpy as np

from .probabilistic import unwrap_posterior
from .types import HMCState, IntegratorState, MCLMCInfo, NUTSInfo, SamplingAlgorithm, key_to_seed

__all__ = ['mclmc', 'nuts', 'KERNELS']


class _Bound:
    """Lazily created Ensemble bound to one (log-posterior closure, chain count)."""

    def __init__(self, logdensity_fn):
        self.model, self.x, self.y = unwrap_posterior(logdensity_fn)
        self.spec = self.model.spec
        self.ens = None
        self.step_counter = 0

    def ensure(self, n_chains):
        if self.ens is None or self.ens.n_chains != n_chains:
            if self.ens is not None:
                self.ens.close()
            self.ens = self.model.make_ensemble(n_chains, self.x, self.y)
        return self.ens

    def flat(self, tree):
        th = self.spec.ravel(tree)
        batched = th.ndim == 2
        return th.reshape(-1, self.spec.n_params), batched

    def tree(self, flat, batched):
        return self.spec.unravel(flat if batched else flat[0])


def mclmc(logdensity_fn: Callable, L, step_size, integrator=None, sqrt_diag_cov=1.0) -> SamplingAlgorithm:
    """blackjax.mclmc(logdensity_fn, L, step_size): isokinetic McLachlan integrator + partial refresh."""
    if integrator is not None:
        raise NotImplementedError('only the default isokinetic_mclachlan integrator is implemented')
    if np.any(np.asarray(sqrt_diag_cov) != 1.0):
        raise NotImplementedError('sqrt_diag_cov != 1 (diagonal preconditioning) is not implemented on the CUDA path')
    b = _Bound(logdensity_fn)

    def init(position, rng_key) -> IntegratorState:
        th, batched = b.flat(position)
        ens = b.ensure(th.shape[0])
        ens.init(th, seed=key_to_seed(rng_key))
        b.step_counter = 0
        return _state(b, ens, batched)

    def step(rng_key, state: IntegratorState):
        th, batched = b.flat(state.position)
        ens = b.ensure(th.shape[0])
        u, _ = b.flat(state.momentum)
        g, _ = b.flat(state.logdensity_grad)
        ens.set_state(th, u, np.atleast_1d(np.asarray(state.logdensity, np.float32)), g)
        _, info = ens.sample(1, step_size, L, step_base=b.step_counter, seed=key_to_seed(rng_key), keep=False, info=True)
        b.step_counter += 1
        inf = info[0] if batched else info[0, 0]
        return _state(b, ens, batched), MCLMCInfo(inf[..., 0], inf[..., 1], inf[..., 2])

    return SamplingAlgorithm(init, step)


def _state(b: _Bound, ens, batched) -> IntegratorState:
    th, u, lp, g = ens.get_state()
    return IntegratorState(b.tree(th, batched), b.tree(u, batched), lp if batched else lp[0], b.tree(g, batched))


def nuts(logdensity_fn: Callable, step_size, inverse_mass_matrix, max_num_doublings: int = 10,
         divergence_threshold: float = 1000.0) -> SamplingAlgorithm:
    """blackjax.nuts(logdensity_fn, step_size, inverse_mass_matrix): init(position) / step(rng_key, state) -> (state, info)
    with info = NUTSInfo (blackjax's field names; the fields the reference keeps, sampling.py:200-210, are filled).  Diagonal metric only."""
    b = _Bound(logdensity_fn)

    def _ensure(th):
        fresh = b.ens is None or b.ens.n_chains != th.shape[0]
        ens = b.ensure(th.shape[0])
        if fresh:
            ens.nuts_init(th, max_num_doublings, divergence_threshold)
            ens.set_nuts_params(step_size, inverse_mass_matrix)
        return ens

    def init(position, rng_key=None) -> HMCState:
        th, batched = b.flat(position)
        ens = _ensure(th)
        ens.nuts_init(th, max_num_doublings, divergence_threshold)
        ens.set_nuts_params(step_size, inverse_mass_matrix)
        b.step_counter = 0
        return _hmc_state(b, ens, batched)

    def step(rng_key, state: HMCState):
        th, batched = b.flat(state.position)
        ens = _ensure(th)
        g, _ = b.flat(state.logdensity_grad)
        ens.set_state(th, None, np.atleast_1d(np.asarray(state.logdensity, np.float32)), g)
        _, info = ens.nuts_sample(1, step_base=b.step_counter, seed=key_to_seed(rng_key), keep=False, info=True)
        b.step_counter += 1
        inf = info[0] if batched else info[0, 0]
        f = {k: inf[..., i] for i, k in enumerate(ens.NUTS_INFO_FIELDS)}
        return _hmc_state(b, ens, batched), NUTSInfo(None, f['is_divergent'] > 0.5, f['is_turning'] > 0.5, f['energy'], None, None,
                                                     f['num_trajectory_expansions'].astype(np.int32),
                                                     f['num_integration_steps'].astype(np.int32), f['acceptance_rate'])

    return SamplingAlgorithm(init, step)


def _hmc_state(b: _Bound, ens, batched) -> HMCState:
    th, _, lp, g = ens.get_state()
    return HMCState(b.tree(th, batched), lp if batched else lp[0], b.tree(g, batched))


def _not_on_cuda(name):
    def f(*a, **k):
        raise NotImplementedError(f"'{name}' is outside the accelerated hot path (SURVEY.md section 2, rows 12): "
                                  'use the reference implementation for it')
    return f


KERNELS: dict = {'nuts': nuts, 'hmc': _not_on_cuda('hmc'), 'mclmc': mclmc}
WARMUP_KERNELS: dict = {}
