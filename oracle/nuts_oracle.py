"""CPU oracle for the NUTS branch of the sampling seam (numpy, dtype-generic: fp32 or fp64).

TEST INFRASTRUCTURE ONLY (same rule as mile_oracle.py: only tests/, smoke() and bench.py's CPU legs may import it).

PARITY UNPINNED.  What is restated:
  * the reference's own `custom_window_adaptation` loop (/root/reference/src/training/warmup.py:27-152: nuts kernel step ->
    `adapt_step(adaptation_state, stage, position, info.acceptance_rate)` per transition, `adapt_final` at the end) and its
    call site `warmup_nuts` + the NUTS branch of `inference_loop` (src/training/sampling.py:70-81,200-210,220-262);
  * from the published algorithms of the un-vendored blackjax 1.2.2 (pyproject.toml:12), from memory:
    `blackjax.mcmc.nuts` (iterative_nuts_proposal), `blackjax.mcmc.trajectory` (dynamic_progressive_integration,
    dynamic_multiplicative_expansion), `blackjax.mcmc.termination.iterative_uturn_numpyro`, `blackjax.mcmc.proposal`
    (progressive_uniform_sampling / progressive_biased_sampling), `blackjax.mcmc.metrics.gaussian_euclidean` (diagonal),
    `blackjax.mcmc.integrators.velocity_verlet`, `blackjax.adaptation.window_adaptation.{base,build_schedule}`,
    `blackjax.adaptation.step_size.dual_averaging_adaptation`, `blackjax.adaptation.mass_matrix` (Welford, diagonal).
It is pinned only by independent checks (tests/test_oracle_nuts.py): reversibility and energy error of the integrator,
detailed-balance statistics on a Gaussian target, the schedule's window structure, dual averaging reaching the target.

Randomness is an INPUT: every transition takes `z` (d standard normals: the momentum draw) and `uni` (uniforms in [0,1),
layout below), so the CPU and the GPU consume identical numbers.  JAX's key splitting is not reproduced.

uni layout for max_num_doublings = D:  uni[0:D]      direction of expansion j   (u < 0.5 -> +1, like bernoulli(0.5) == True)
                                       uni[D:2D]     biased progressive acceptance after expansion j
                                       uni[2D + n]   uniform progressive acceptance of the n-th leapfrog state of the
                                                     transition (n counts every integrator call of the transition from 0)
"""
from __future__ import annotations

from typing import Callable, NamedTuple

import numpy as np


class NutsInfo(NamedTuple):
    num_integration_steps: int
    acceptance_rate: float
    num_trajectory_expansions: int
    is_divergent: bool
    energy: float
    is_turning: bool


def uni_len(max_num_doublings: int) -> int:
    return 2 * max_num_doublings + 2 ** max_num_doublings


def _logaddexp(a, b, dt):
    return dt(np.logaddexp(dt(a), dt(b)))


def kinetic_energy(p, imm):
    """metrics.gaussian_euclidean (diagonal): 0.5 * p . (M^-1 p)."""
    dt = p.dtype.type
    return dt(0.5) * np.dot(p, imm * p).astype(p.dtype)


def is_turning(imm, p_left, p_right, p_sum):
    """metrics.gaussian_euclidean.is_turning (generalised U-turn criterion, Betancourt 2013)."""
    dt = p_sum.dtype.type
    rho = p_sum - (p_right + p_left) / dt(2)
    return bool(np.dot(imm * p_left, rho) <= 0) or bool(np.dot(imm * p_right, rho) <= 0)


def velocity_verlet(logdensity_and_grad, theta, p, g, step, imm):
    """integrators.velocity_verlet: p += step/2 g; theta += step M^-1 p; p += step/2 g(theta')."""
    dt = theta.dtype.type
    half = dt(0.5) * dt(step)
    p = p + half * g
    theta = theta + dt(step) * (imm * p)
    lp, g = logdensity_and_grad(theta)
    p = p + half * g
    return theta, p, dt(lp), g


def leaf_idx_to_ckpt_idxs(n: int):
    """termination._leaf_idx_to_ckpt_idxs: (number of set bits of n >> 1, that minus the trailing ones of n plus 1)."""
    idx_max = bin(n >> 1).count('1')
    t, num_subtrees = n, 0
    while t & 1:
        t >>= 1
        num_subtrees += 1
    return idx_max - num_subtrees + 1, idx_max


def nuts_step(logdensity_and_grad: Callable, theta, lp, g, step_size, imm, z, uni, max_num_doublings: int = 10,
              divergence_threshold: float = 1000.0):
    """One transition of blackjax.mcmc.nuts.build_kernel()(...) with the randomness supplied.

    Returns (theta', lp', g', NutsInfo)."""
    dt = theta.dtype.type
    D = max_num_doublings
    imm = imm.astype(theta.dtype)
    p0 = z.astype(theta.dtype) / np.sqrt(imm)                    # momentum_generator: sqrt(M) * N(0, I)
    e0 = -dt(lp) + kinetic_energy(p0, imm)
    # proposal = (theta, lp, g, energy, weight, sum_log_p_accept); trajectory = (left, right, p_sum, n)
    prop = (theta, dt(lp), g, e0, dt(0), dt(-np.inf))
    left = right = (theta, p0, dt(lp), g)
    p_sum, n_states = p0.copy(), 0
    ck_p, ck_s = np.zeros((D, theta.shape[0]), theta.dtype), np.zeros((D, theta.shape[0]), theta.dtype)
    n_leap = 0
    step, diverging, turning = 0, False, False
    eps = dt(step_size)
    while step < D and not diverging and not turning:
        direction = 1 if uni[step] < 0.5 else -1
        start = right if direction > 0 else left
        # ---- trajectory.dynamic_progressive_integration: up to 2^step leapfrog states, extended "to the right" ----
        cur = start
        s_prop, s_first, s_sum, s_n = None, None, None, 0
        s_div, s_term = False, False
        k = 0
        while k < 2 ** step and not s_term and not s_div:
            th, p, l, gg = velocity_verlet(logdensity_and_grad, cur[0], cur[1], cur[3], dt(direction) * eps, imm)
            cur = (th, p, l, gg)
            new_e = -l + kinetic_energy(p, imm)
            delta = e0 - new_e
            if np.isnan(delta):
                delta = dt(-np.inf)
            s_div = bool(abs(delta) > divergence_threshold)
            new_prop = (th, l, gg, new_e, dt(delta), dt(min(delta, dt(0))))
            if k == 0:
                s_first, s_sum, s_n, s_prop = cur, p.copy(), 1, new_prop
            else:
                s_sum, s_n = s_sum + p, s_n + 1
                # proposal.progressive_uniform_sampling
                with np.errstate(over='ignore', invalid='ignore'):
                    p_acc = dt(1) / (dt(1) + np.exp(-(new_prop[4] - s_prop[4])))
                w = _logaddexp(s_prop[4], new_prop[4], dt)
                sl = _logaddexp(s_prop[5], new_prop[5], dt)
                keep = new_prop if uni[2 * D + n_leap] < p_acc else s_prop
                s_prop = (keep[0], keep[1], keep[2], keep[3], w, sl)
            # termination.iterative_uturn_numpyro
            idx_min, idx_max = leaf_idx_to_ckpt_idxs(k)
            if k % 2 == 0:
                ck_p[idx_max], ck_s[idx_max] = p, s_sum
            else:
                i = idx_max
                while i >= idx_min and not s_term:
                    s_term = is_turning(imm, ck_p[i], p, s_sum - ck_s[i] + ck_p[i])
                    i -= 1
            k += 1
            n_leap += 1
        # ---- trajectory.dynamic_multiplicative_expansion: merge, biased progressive sampling, whole-trajectory U-turn ----
        if direction > 0:
            left, right = left, cur
        else:
            left, right = cur, right
        p_sum, n_states = p_sum + s_sum, n_states + s_n
        if s_div or s_term:
            prop = (prop[0], prop[1], prop[2], prop[3], prop[4], _logaddexp(prop[5], s_prop[5], dt))
        else:
            with np.errstate(over='ignore'):
                p_acc = min(dt(1), np.exp(s_prop[4] - prop[4]))
            w = _logaddexp(prop[4], s_prop[4], dt)
            sl = _logaddexp(prop[5], s_prop[5], dt)
            keep = s_prop if uni[D + step] < p_acc else prop
            prop = (keep[0], keep[1], keep[2], keep[3], w, sl)
        turning_whole = is_turning(imm, left[1], right[1], p_sum)
        diverging, turning = s_div, (s_term or turning_whole)
        step += 1
    acc = dt(np.exp(prop[5])) / dt(max(n_states, 1)) if n_states > 0 else dt(0)
    info = NutsInfo(n_states, float(acc), step, bool(diverging), float(prop[3]), bool(turning))
    return prop[0], prop[1], prop[2], info


# --------------------------------------------------------------------------------------
# Window adaptation (Stan's scheme as blackjax 1.2.2 implements it)
# --------------------------------------------------------------------------------------

def build_schedule(num_steps: int, initial_buffer_size: int = 75, final_buffer_size: int = 50, first_window_size: int = 25):
    """window_adaptation.build_schedule -> [(stage, is_middle_window_end)] * num_steps; stage 0 = fast, 1 = slow."""
    schedule = []
    if num_steps < 20:
        return [(0, False)] * num_steps
    if initial_buffer_size + first_window_size + final_buffer_size > num_steps:
        initial_buffer_size = int(0.15 * num_steps)
        final_buffer_size = int(0.1 * num_steps)
        first_window_size = num_steps - initial_buffer_size - final_buffer_size
    schedule += [(0, False)] * initial_buffer_size
    final_buffer_start = num_steps - final_buffer_size
    next_window_size, next_window_start = first_window_size, initial_buffer_size
    while next_window_start < final_buffer_start:
        current_start, current_size = next_window_start, next_window_size
        if 3 * current_size <= final_buffer_start - current_start:
            next_window_size = 2 * current_size
        else:
            current_size = final_buffer_start - current_start
        next_window_start = current_start + current_size
        schedule += [(1, False)] * (next_window_start - 1 - current_start)
        schedule.append((1, True))
    schedule += [(0, False)] * (num_steps - final_buffer_start)
    return schedule


class AdaptState(NamedTuple):
    """WindowAdaptationState flattened: dual averaging (log_x, log_x_avg, step, avg_error, mu), Welford (mean, m2, count),
    and the parameters the kernel uses next (step_size, inverse_mass_matrix)."""
    log_x: float
    log_x_avg: float
    da_step: int
    avg_error: float
    mu: float
    mean: np.ndarray
    m2: np.ndarray
    count: int
    step_size: float
    imm: np.ndarray


def _da_init(step_size, dt):
    return dt(np.log(dt(step_size))), dt(0), 1, dt(0), dt(np.log(dt(10) * dt(step_size)))


def adapt_init(d: int, initial_step_size: float = 1.0, dt=np.float64) -> AdaptState:
    lx, lxa, st, ae, mu = _da_init(initial_step_size, dt)
    return AdaptState(lx, lxa, st, ae, mu, np.zeros(d, dt), np.zeros(d, dt), 0, dt(initial_step_size), np.ones(d, dt))


def _da_update(s: AdaptState, acceptance_rate, target, dt, t0=10, gamma=0.05, kappa=0.75):
    """optimizers.dual_averaging update with gradient = target - acceptance_rate (step_size.dual_averaging_adaptation)."""
    grad = dt(target) - dt(acceptance_rate)
    reg_step = dt(s.da_step + t0)
    eta = dt(s.da_step) ** dt(-kappa)
    avg_error = (dt(1) - dt(1) / reg_step) * s.avg_error + grad / reg_step
    log_x = s.mu - (np.sqrt(dt(s.da_step)) / dt(gamma)) * avg_error
    log_x_avg = eta * log_x + (dt(1) - eta) * s.log_x_avg
    return dt(log_x), dt(log_x_avg), s.da_step + 1, dt(avg_error)


def adapt_step(s: AdaptState, stage, position, acceptance_rate, target: float = 0.8) -> AdaptState:
    """window_adaptation.base.update: fast (stage 0) or slow (stage 1) update, then slow_final at a window end."""
    dt = s.mean.dtype.type
    st, window_end = stage
    mean, m2, count = s.mean, s.m2, s.count
    if st == 1:     # mass_matrix.welford update (diagonal)
        count = count + 1
        delta = position.astype(mean.dtype) - mean
        mean = mean + delta / dt(count)
        m2 = m2 + delta * (position.astype(mean.dtype) - mean)
    lx, lxa, dst, ae = _da_update(s, acceptance_rate, target, dt)
    s = AdaptState(lx, lxa, dst, ae, s.mu, mean, m2, count, dt(np.exp(lx)), s.imm)
    if window_end:  # slow_final: new metric from the window, dual averaging restarted at the averaged step size
        cov = s.m2 / dt(s.count - 1)
        scaled = (dt(s.count) / dt(s.count + 5)) * cov
        shrink = dt(1e-3) * (dt(5) / dt(s.count + 5))
        imm = scaled + shrink
        lx, lxa, dst, ae, mu = _da_init(np.exp(s.log_x_avg), dt)
        s = AdaptState(lx, lxa, dst, ae, mu, np.zeros_like(s.mean), np.zeros_like(s.m2), 0, dt(np.exp(lx)), imm)
    return s


def adapt_final(s: AdaptState):
    return s.mean.dtype.type(np.exp(s.log_x_avg)), s.imm


def run_window_adaptation(logdensity_and_grad, theta, z_steps, uni_steps, initial_step_size: float = 1.0,
                          target: float = 0.8, max_num_doublings: int = 10):
    """custom_window_adaptation(...).run (src/training/warmup.py:112-150) for one chain with the randomness supplied.

    Returns (theta, lp, g, step_size, inverse_mass_matrix, infos)."""
    dt = theta.dtype.type
    n = len(z_steps)
    lp, g = logdensity_and_grad(theta)
    st = adapt_init(theta.shape[0], initial_step_size, dt)
    schedule = build_schedule(n)
    infos = []
    for k in range(n):
        theta, lp, g, info = nuts_step(logdensity_and_grad, theta, dt(lp), g, st.step_size, st.imm, z_steps[k], uni_steps[k],
                                       max_num_doublings)
        st = adapt_step(st, schedule[k], theta, info.acceptance_rate, target)
        infos.append(info)
    step_size, imm = adapt_final(st)
    return theta, lp, g, step_size, imm, infos


def run_nuts(logdensity_and_grad, theta, lp, g, step_size, imm, z_steps, uni_steps, max_num_doublings: int = 10):
    """scan(sampler.step) of the NUTS branch (src/training/sampling.py:107-177): positions AFTER every transition."""
    out, infos = [], []
    for k in range(len(z_steps)):
        theta, lp, g, info = nuts_step(logdensity_and_grad, theta, lp, g, step_size, imm, z_steps[k], uni_steps[k],
                                       max_num_doublings)
        out.append(theta.copy())
        infos.append(info)
    return theta, lp, g, np.stack(out), infos
