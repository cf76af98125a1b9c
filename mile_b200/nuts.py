"""NUTS branch of the sampling seam (mirror of src/training/warmup.py:27-152 `custom_window_adaptation` and
src/training/sampling.py:220-262 `warmup_nuts`).  Whole transitions -- momentum draw, trajectory doubling, U-turn
checkpoints, progressive sampling -- and the per-transition adaptation run inside ONE persistent CUDA kernel per chunk
(csrc/mile_nuts.cuh, `mile_nuts_*`), all chains of the wave at once."""
from __future__ import annotations

import logging

import numpy as np

from .engine import Ensemble
from .probabilistic import unwrap_posterior
from .types import AdaptationAlgorithm, AdaptationResults, key_to_seed

logger = logging.getLogger(__name__)

CHUNK = 500   # transitions per launch (a transition is up to 2^max_num_doublings gradient evaluations)


def build_schedule(num_steps: int, initial_buffer_size: int = 75, final_buffer_size: int = 50,
                   first_window_size: int = 25) -> list:
    """blackjax.adaptation.window_adaptation.build_schedule (imported by warmup.py:13, used at warmup.py:133): Stan's
    fast initial buffer, doubling slow windows, fast final buffer -> [(stage, is_middle_window_end)] per step."""
    if num_steps < 20:
        return [(0, False)] * num_steps
    if initial_buffer_size + first_window_size + final_buffer_size > num_steps:
        initial_buffer_size = int(0.15 * num_steps)
        final_buffer_size = int(0.1 * num_steps)
        first_window_size = num_steps - initial_buffer_size - final_buffer_size
    schedule = [(0, False)] * initial_buffer_size
    final_start = num_steps - final_buffer_size
    size, start = first_window_size, initial_buffer_size
    while start < final_start:
        cur_start, cur_size = start, size
        if 3 * cur_size <= final_start - cur_start:
            size = 2 * cur_size
        else:
            cur_size = final_start - cur_start
        start = cur_start + cur_size
        schedule += [(1, False)] * (start - 1 - cur_start) + [(1, True)]
    return schedule + [(0, False)] * (num_steps - final_start)


def run_window_adaptation(ens: Ensemble, theta0: np.ndarray, rng_key, num_steps: int, *, initial_step_size: float = 1.0,
                          target_acceptance_rate: float = 0.8, max_num_doublings: int = 10,
                          divergence_threshold: float = 1000.0, saving_path=None, step_ids=None):
    """custom_window_adaptation(...).run (warmup.py:112-150) for all chains of `ens`.  Returns (step_size [C],
    inverse_mass_matrix [C, d]); the warmed-up chain state and the adapted parameters stay in `ens`.  With `saving_path`
    the position before every warm-up step n is written to saving_path/{step_id}/sample_{n}.npz (warmup.py:102-109)."""
    seed = key_to_seed(rng_key)
    ens.nuts_init(theta0, max_num_doublings, divergence_threshold, target_acceptance_rate, initial_step_size)
    schedule = build_schedule(int(num_steps))
    writer = None
    if saving_path:
        from .callbacks import SampleWriter
        ids = list(range(ens.n_chains)) if step_ids is None else [int(s) for s in np.atleast_1d(np.asarray(step_ids))]
        writer = SampleWriter(ens.spec, saving_path, ids)
    done = 0
    while done < num_steps:
        n = min(CHUNK, num_steps - done)
        if writer is not None:
            pos, _ = ens.nuts_warmup(n, schedule[done:done + n], step_base=done, seed=seed, keep=True)
            writer.submit(pos, range(done, done + n))
        else:
            ens.nuts_warmup(n, schedule[done:done + n], step_base=done, seed=seed)
        done += n
    if writer is not None:
        writer.close()
    ens.nuts_finish_warmup()                                               # adapt_final, warmup.py:142
    return ens.nuts_params()


def custom_window_adaptation(algorithm, logdensity_fn, is_mass_matrix_diagonal: bool = True, initial_step_size: float = 1.0,
                             target_acceptance_rate: float = 0.80, progress_bar: bool = False, saving_path=None,
                             **extra_parameters) -> AdaptationAlgorithm:
    """warmup.py:27-152 (same argument names and defaults).  `algorithm` is accepted for signature parity; the NUTS kernel
    is the library's."""
    if not is_mass_matrix_diagonal:
        raise NotImplementedError('only the diagonal mass matrix (the reference default) is adapted on the CUDA path')
    model, x, y = unwrap_posterior(logdensity_fn)
    spec = model.spec

    def run(rng_key, position, device_id=None, num_steps: int = 1000, n_devices: int = 1):
        from .types import HMCState
        theta0 = spec.ravel(position)
        batched = theta0.ndim == 2
        theta0 = theta0.reshape(-1, spec.n_params)
        ens = model.make_ensemble(theta0.shape[0], x, y)
        try:
            ids = None if device_id is None else np.unique(np.asarray(device_id).reshape(theta0.shape[0], -1)[:, 0])
            eps, imm = run_window_adaptation(ens, theta0, rng_key, num_steps, initial_step_size=initial_step_size,
                                             target_acceptance_rate=target_acceptance_rate, saving_path=saving_path,
                                             step_ids=ids, **extra_parameters)
            th, _, lp, g = ens.get_state()
        finally:
            ens.close()
        un = (lambda a: spec.unravel(a)) if batched else (lambda a: spec.unravel(a[0]))
        state = HMCState(un(th), lp if batched else lp[0], un(g))
        params = {'step_size': eps if batched else eps[0], 'inverse_mass_matrix': imm if batched else imm[0],
                  **extra_parameters}
        return AdaptationResults(state, params)

    return AdaptationAlgorithm(run)
