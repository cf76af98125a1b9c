"""Small value types mirroring the reference / blackjax ones (src/types.py, blackjax.mcmc.integrators,
blackjax.mcmc.mclmc, blackjax.adaptation.mclmc_adaptation, blackjax.base)."""
from __future__ import annotations

from typing import Any, Callable, NamedTuple

import numpy as np

ParamTree = dict  # dict[str, np.ndarray | ParamTree]  (src/types.py:13)


class IntegratorState(NamedTuple):
    """blackjax.mcmc.integrators.IntegratorState: what `sampler.init/step` carry."""
    position: Any
    momentum: Any
    logdensity: Any
    logdensity_grad: Any


class HMCState(NamedTuple):
    """blackjax.mcmc.hmc.HMCState (the state of the NUTS branch)."""
    position: Any
    logdensity: Any
    logdensity_grad: Any


class NUTSInfo(NamedTuple):
    """blackjax.mcmc.nuts.NUTSInfo (field names and order).  The trajectory end states are not exported by the CUDA path."""
    momentum: Any
    is_divergent: Any
    is_turning: Any
    energy: Any
    trajectory_leftmost_state: Any
    trajectory_rightmost_state: Any
    num_trajectory_expansions: Any
    num_integration_steps: Any
    acceptance_rate: Any


class MCLMCInfo(NamedTuple):
    """blackjax.mcmc.mclmc.MCLMCInfo."""
    logdensity: Any
    kinetic_change: Any
    energy_change: Any


class MCLMCAdaptationState(NamedTuple):
    """blackjax.adaptation.mclmc_adaptation.MCLMCAdaptationState (src/training/warmup.py:205,403)."""
    L: Any
    step_size: Any
    sqrt_diag_cov: Any


class SamplingAlgorithm(NamedTuple):
    """blackjax.base.SamplingAlgorithm."""
    init: Callable
    step: Callable


class AdaptationAlgorithm(NamedTuple):
    """blackjax.base.AdaptationAlgorithm."""
    run: Callable


class AdaptationResults(NamedTuple):
    """blackjax.base.AdaptationResults."""
    state: Any
    parameters: Any


# ---- PRNG keys ----------------------------------------------------------------------------------
# JAX threefry key splitting cannot be reproduced without jax, so keys are mapped to
# 64-bit seeds of the in-kernel Philox4x32-10 generator.  Accepted: python ints, uint32[2] arrays
# (the raw form of a jax PRNGKey) or anything with `.tolist()`.
_MASK = (1 << 64) - 1


def key_to_seed(rng_key) -> int:
    if isinstance(rng_key, (int, np.integer)):
        return int(rng_key) & _MASK
    a = np.asarray(rng_key).astype(np.uint64).reshape(-1)
    s = 0
    for w in a.tolist():
        s = ((s << 32) ^ int(w)) & _MASK
    return s


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _MASK
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
    return z ^ (z >> 31)


def split(rng_key, num: int = 2) -> list:
    """Stand-in for jax.random.split: `num` statistically independent child seeds."""
    s = key_to_seed(rng_key)
    return [_splitmix64((s + 0x632BE59BD9B4E019 * (i + 1)) & _MASK) for i in range(num)]
