"""CPU oracle for the MILE MCLMC ensemble sampling hot path (numpy, fp32 + fp64 twin).

TEST INFRASTRUCTURE ONLY.  Nothing under ``mile_b200/`` may import this module; it is
the checker used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.

PARITY UNPINNED: the reference (zhiyuan-yang/MILE) ships no tests, golden vectors or
fixtures, its sampler arithmetic lives in the un-vendored
third-party package ``blackjax==1.2.2`` (pyproject.toml:12, poetry.lock:358-359) and
JAX/BlackJAX are not installable in this image, so this restatement cannot be checked
against reference outputs.  It is pinned instead by independent implementations and
closed-form identities (tests/test_oracle_*.py): torch.autograd, scipy.stats, central
finite differences in fp64, ESH invariants, Gaussian-target statistics.

Every function cites the reference file:line (relative to /root/reference) or the
blackjax 1.2.2 module it follows.  Noise is always an INPUT (host-supplied ``z``
arrays), never generated inside, so CPU and GPU consume identical bits.

Flat parameter layout (``jax.flatten_util.ravel_pytree`` order = sorted dict keys,
SURVEY.md section 5): for every layer in lexicographic order of ``layer{i}``:
``bias (out)`` then ``kernel (in, out) row-major``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import NamedTuple, Sequence

import numpy as np

# --------------------------------------------------------------------------------------
# Model description
# --------------------------------------------------------------------------------------

ACTIVATIONS = ('identity', 'relu', 'sigmoid', 'tanh', 'gelu', 'leaky_relu')
# src/config/models/base.py:24-37 lists sigmoid, relu, gelu, tanh, softmax, leaky_relu
# (flax.linen.<name>).  'softmax' as a hidden activation is not used by any shipped
# YAML and is not restated here.


@dataclass(frozen=True)
class ModelSpec:
    """FCN + task + prior: what `log_unnormalized_posterior` closes over.

    src/models/tabular/fcn.py:11-28 (hidden_structure INCLUDES the output width),
    src/training/probabilistic.py:20-47, src/training/priors.py:58-90.
    """

    n_features: int
    widths: tuple  # hidden_structure, last entry = output width K
    activation: str = 'relu'
    task: str = 'regr'  # 'regr' | 'class'
    prior: str = 'normal'  # 'normal' | 'laplace'
    prior_loc: float = 0.0
    prior_scale: float = 1.0
    n_batches: int = 1  # probabilistic.py:136 multiplies the likelihood by n_batches
    layer_order: tuple = field(default=None)  # order of layers inside the flat vector

    def __post_init__(self):
        object.__setattr__(self, 'widths', tuple(int(w) for w in self.widths))
        assert self.activation in ACTIVATIONS, self.activation
        assert self.task in ('regr', 'class')
        if self.layer_order is None:
            # JAX flattens dicts in sorted key order: 'layer10' < 'layer2'.
            order = sorted(range(len(self.widths)), key=lambda i: f'layer{i}')
            object.__setattr__(self, 'layer_order', tuple(order))

    @property
    def n_layers(self) -> int:
        return len(self.widths)

    @property
    def dims(self) -> tuple:
        return (self.n_features,) + self.widths

    @property
    def n_params(self) -> int:
        d = self.dims
        return sum(d[i] * d[i + 1] + d[i + 1] for i in range(self.n_layers))

    def offsets(self):
        """Return (bias_off[l], kernel_off[l]) into the flat vector."""
        d = self.dims
        bias_off = [0] * self.n_layers
        kern_off = [0] * self.n_layers
        off = 0
        for l in self.layer_order:
            bias_off[l] = off
            off += d[l + 1]
            kern_off[l] = off
            off += d[l] * d[l + 1]
        return bias_off, kern_off


def unravel(spec: ModelSpec, theta: np.ndarray):
    """Flat vector -> list of (kernel[in,out], bias[out]) per layer (views)."""
    bias_off, kern_off = spec.offsets()
    d = spec.dims
    layers = []
    for l in range(spec.n_layers):
        b = theta[..., bias_off[l]:bias_off[l] + d[l + 1]]
        w = theta[..., kern_off[l]:kern_off[l] + d[l] * d[l + 1]]
        w = w.reshape(theta.shape[:-1] + (d[l], d[l + 1]))
        layers.append((w, b))
    return layers


def ravel_tree(spec: ModelSpec, tree: dict) -> np.ndarray:
    """{'fcn': {'layer0': {'kernel','bias'}, ...}} -> flat vector (ravel_pytree order)."""
    inner = tree['fcn'] if 'fcn' in tree else tree
    parts = []
    for l in spec.layer_order:
        lay = inner[f'layer{l}']
        parts.append(np.asarray(lay['bias']).reshape(-1))
        parts.append(np.asarray(lay['kernel']).reshape(-1))
    return np.concatenate(parts)


def unravel_tree(spec: ModelSpec, theta: np.ndarray) -> dict:
    layers = unravel(spec, theta)
    return {'fcn': {f'layer{l}': {'bias': np.array(b), 'kernel': np.array(w)}
                    for l, (w, b) in enumerate(layers)}}


# --------------------------------------------------------------------------------------
# Activations (flax.linen.<name>; jax.nn)
# --------------------------------------------------------------------------------------

def _act(name: str, z: np.ndarray):
    """Return (a, da/dz).  relu'(0)=0 (jax.nn.relu custom jvp), leaky slope 0.01,
    gelu = tanh approximation (flax.linen.gelu default approximate=True)."""
    dt = z.dtype.type
    if name == 'identity':
        return z, np.ones_like(z)
    if name == 'relu':
        return np.maximum(z, dt(0)), (z > 0).astype(z.dtype)
    if name == 'leaky_relu':
        return np.where(z >= 0, z, dt(0.01) * z), np.where(z >= 0, dt(1), dt(0.01)).astype(z.dtype)
    if name == 'sigmoid':
        s = dt(1) / (dt(1) + np.exp(-z))
        return s, s * (dt(1) - s)
    if name == 'tanh':
        t = np.tanh(z)
        return t, dt(1) - t * t
    if name == 'gelu':
        c = dt(math.sqrt(2.0 / math.pi))
        k = dt(0.044715)
        inner = c * (z + k * z * z * z)
        t = np.tanh(inner)
        a = dt(0.5) * z * (dt(1) + t)
        dinner = c * (dt(1) + dt(3) * k * z * z)
        da = dt(0.5) * (dt(1) + t) + dt(0.5) * z * (dt(1) - t * t) * dinner
        return a, da
    raise NotImplementedError(name)


# --------------------------------------------------------------------------------------
# a5: FCN forward  (src/flax_building_blocks/basic.py:41-61, src/models/tabular/fcn.py:26-28)
# --------------------------------------------------------------------------------------

def forward(spec: ModelSpec, theta: np.ndarray, X: np.ndarray, keep: bool = False):
    """x <- act(x @ kernel_i + bias_i) for all but the last layer; last layer linear."""
    layers = unravel(spec, theta)
    a = X
    acts, dacts = [X], []
    for l, (w, b) in enumerate(layers):
        z = a @ w + b
        if l < spec.n_layers - 1:
            a, da = _act(spec.activation, z)
        else:
            a, da = z, None
        acts.append(a)
        dacts.append(da)
    return (a, acts, dacts) if keep else a


# --------------------------------------------------------------------------------------
# a1-a4: log posterior value and gradient
# (src/training/probabilistic.py:92-138, src/training/priors.py:101-128)
# --------------------------------------------------------------------------------------

_LOG_2PI = math.log(2.0 * math.pi)


def log_prior(spec: ModelSpec, theta: np.ndarray):
    """priors.py:101-108 (Normal: jax.scipy.stats.norm.logpdf summed) / :121-128 (Laplace)."""
    dt = theta.dtype.type
    loc, s = dt(spec.prior_loc), dt(spec.prior_scale)
    if spec.prior == 'normal':
        s2 = s * s
        val = np.sum((np.log(dt(2.0 * math.pi) * s2) + np.square(theta - loc) / s2) / dt(-2))
        grad = -(theta - loc) / s2
    elif spec.prior == 'laplace':
        val = np.sum(-np.log(dt(2) * s) - np.abs(theta - loc) / s)
        grad = -np.sign(theta - loc) / s
    else:
        raise NotImplementedError(spec.prior)
    return val, grad


def pointwise_loglik(spec: ModelSpec, out: np.ndarray, y: np.ndarray):
    """Per-row log-likelihood and d/d(out).

    Regression (probabilistic.py:93-100): norm.logpdf(y; loc=out[...,0],
    scale=clip(exp(out[...,1]),1e-6,1e6)) with jax's formula
    -(log(2 pi s^2) + (y-mu)^2/s^2)/2.  clip passes gradient strictly inside only.
    Classification (probabilistic.py:101-109): out[n,y_n] - logsumexp_k out[n,k].
    """
    dt = out.dtype.type
    if spec.task == 'regr':
        mu, s = out[..., 0], out[..., 1]
        e = np.exp(s)
        sigma = np.clip(e, dt(1e-6), dt(1e6))
        inside = ((e > dt(1e-6)) & (e < dt(1e6))).astype(out.dtype)
        s2 = sigma * sigma
        r = y - mu
        ll = (np.log(dt(2.0 * math.pi) * s2) + r * r / s2) / dt(-2)
        dmu = r / s2
        ds = (r * r / s2 - dt(1)) * inside
        dout = np.stack([dmu, ds], axis=-1)
    else:
        m = out.max(axis=-1, keepdims=True)
        ex = np.exp(out - m)
        se = ex.sum(axis=-1, keepdims=True)
        lse = m + np.log(se)
        yi = y.astype(np.int64)
        ll = np.take_along_axis(out, yi[..., None], axis=-1)[..., 0] - lse[..., 0]
        dout = -(ex / se)
        np.put_along_axis(dout, yi[..., None],
                          np.take_along_axis(dout, yi[..., None], axis=-1) + dt(1), axis=-1)
    return ll, dout


def logpost_value_and_grad(spec: ModelSpec, theta: np.ndarray, X: np.ndarray, y: np.ndarray):
    """value_and_grad of ProbabilisticModel.log_unnormalized_posterior (probabilistic.py:115-138)
    for one chain.  jnp.nansum: NaN rows contribute 0 to the value; their cotangent is 0 and
    this restatement (like the CUDA path) drops their gradient contribution.
    """
    dt = theta.dtype
    X = X.astype(dt, copy=False)
    out, acts, dacts = forward(spec, theta, X, keep=True)
    ll, dout = pointwise_loglik(spec, out, y if spec.task == 'class' else y.astype(dt, copy=False))
    nan = np.isnan(ll)
    ll = np.where(nan, dt.type(0), ll)
    dout = np.where(nan[..., None], dt.type(0), dout)
    nb = dt.type(spec.n_batches)
    lik = ll.sum(dtype=dt) * nb
    layers = unravel(spec, theta)
    bias_off, kern_off = spec.offsets()
    grad = np.zeros_like(theta)
    delta = dout * nb
    for l in reversed(range(spec.n_layers)):
        w, _ = layers[l]
        a_in = acts[l]
        gw = a_in.T @ delta
        gb = delta.sum(axis=0, dtype=dt)
        grad[kern_off[l]:kern_off[l] + gw.size] = gw.reshape(-1)
        grad[bias_off[l]:bias_off[l] + gb.size] = gb
        if l > 0:
            delta = (delta @ w.T) * dacts[l - 1]
    pv, pg = log_prior(spec, theta)
    return (pv + lik).astype(dt), grad + pg


def logpost_batch(spec: ModelSpec, thetas: np.ndarray, X, y):
    """All chains: thetas [C,d] -> (lp[C], grad[C,d])."""
    lps, gs = [], []
    for c in range(thetas.shape[0]):
        lp, g = logpost_value_and_grad(spec, thetas[c], X, y)
        lps.append(lp)
        gs.append(g)
    return np.array(lps, dtype=thetas.dtype), np.stack(gs)


# --------------------------------------------------------------------------------------
# a14: blackjax 1.2.2 MCLMC arithmetic (third-party, restated)
# --------------------------------------------------------------------------------------

class IntegratorState(NamedTuple):
    """blackjax.mcmc.integrators.IntegratorState."""
    position: np.ndarray
    momentum: np.ndarray
    logdensity: np.floating
    logdensity_grad: np.ndarray


class MCLMCInfo(NamedTuple):
    """blackjax.mcmc.mclmc.MCLMCInfo."""
    logdensity: np.floating
    kinetic_change: np.floating
    energy_change: np.floating


MCLACHLAN_B1 = 0.1931833275037836
MCLACHLAN_COEFFS = (MCLACHLAN_B1, 0.5, 1.0 - 2.0 * MCLACHLAN_B1, 0.5, MCLACHLAN_B1)


def _norm(x):
    return np.sqrt(np.sum(x * x, dtype=x.dtype))


def _normalized(x, tol=1e-13):
    """blackjax.mcmc.integrators._normalized_flatten_array."""
    n = _norm(x)
    return (x / n if n > tol else x), n


def generate_unit_vector(z: np.ndarray):
    """blackjax.util.generate_unit_vector with the normal draw z supplied by the host."""
    return z / _norm(z)


def mclmc_init(logdensity_and_grad, position: np.ndarray, z0: np.ndarray) -> IntegratorState:
    """blackjax.mcmc.mclmc.init (call site src/training/warmup.py:539-541)."""
    if position.shape[-1] < 2:
        raise ValueError('The target distribution must have more than 1 dimension for MCLMC.')
    lp, g = logdensity_and_grad(position)
    return IntegratorState(position, generate_unit_vector(z0.astype(position.dtype)), lp, g)


def esh_momentum_update(u, g, step_size, coef, prev_dk, sqrt_diag_cov=1.0):
    """blackjax.mcmc.integrators.esh_dynamics_momentum_update_one_step (B-step)."""
    dt = u.dtype.type
    g = g * dt(sqrt_diag_cov) if np.isscalar(sqrt_diag_cov) else g * sqrt_diag_cov.astype(u.dtype)
    d = u.shape[0]
    e, gn = _normalized(g)
    p = np.dot(u, e).astype(u.dtype)
    delta = dt(step_size) * dt(coef) * gn / dt(d - 1)
    zeta = np.exp(-delta)
    raw = e * (dt(1) - zeta) * (dt(1) + zeta + p * (dt(1) - zeta)) + dt(2) * zeta * u
    new_u, _ = _normalized(raw)
    dk = (delta - dt(math.log(2.0)) + np.log(dt(1) + p + (dt(1) - p) * zeta * zeta)) * dt(d - 1)
    return new_u, (dk + prev_dk).astype(u.dtype)


def isokinetic_mclachlan(logdensity_and_grad, state: IntegratorState, step_size, sqrt_diag_cov=1.0):
    """blackjax.mcmc.integrators.isokinetic_mclachlan: B(b1) A(.5) B(1-2b1) A(.5) B(b1)."""
    dt = state.position.dtype.type
    theta, u, lp, g = state
    dk = dt(0)
    eps = dt(step_size)
    b1, a1, b2, _, _ = MCLACHLAN_COEFFS
    u, dk = esh_momentum_update(u, g, eps, b1, dk, sqrt_diag_cov)
    theta = theta + eps * dt(a1) * (u * sqrt_diag_cov if not np.isscalar(sqrt_diag_cov) else u * dt(sqrt_diag_cov))
    lp, g = logdensity_and_grad(theta)
    u, dk = esh_momentum_update(u, g, eps, b2, dk, sqrt_diag_cov)
    theta = theta + eps * dt(a1) * (u * sqrt_diag_cov if not np.isscalar(sqrt_diag_cov) else u * dt(sqrt_diag_cov))
    lp, g = logdensity_and_grad(theta)
    u, dk = esh_momentum_update(u, g, eps, b1, dk, sqrt_diag_cov)
    return IntegratorState(theta, u, lp, g), dk


def partially_refresh_momentum(u, z, step_size, L):
    """blackjax.mcmc.integrators.partially_refresh_momentum with host-supplied z ~ N(0,I)."""
    dt = u.dtype.type
    d = u.shape[0]
    if np.isinf(L):
        return u
    nu = np.sqrt((np.exp(dt(2) * dt(step_size) / dt(L)) - dt(1)) / dt(d))
    v = u + nu * z.astype(u.dtype)
    return v / _norm(v)


def mclmc_step(logdensity_and_grad, state: IntegratorState, step_size, L, z,
               sqrt_diag_cov=1.0, refresh='post'):
    """blackjax.mcmc.mclmc.build_kernel(...).kernel(rng_key, state, L, step_size).

    refresh='post': one full-step refresh after the deterministic step (blackjax 1.2.2,
    the pinned version).  refresh='maruyama': half-step refreshes around the integrator
    (later 1.2.x `with_isokinetic_maruyama`); z must then be [2,d].  SURVEY.md Appendix A.
    """
    dt = state.position.dtype.type
    if refresh == 'post':
        new, dk = isokinetic_mclachlan(logdensity_and_grad, state, step_size, sqrt_diag_cov)
        u = partially_refresh_momentum(new.momentum, z, step_size, L)
    else:
        u0 = partially_refresh_momentum(state.momentum, z[0], dt(step_size) * dt(0.5), L)
        new, dk = isokinetic_mclachlan(logdensity_and_grad, state._replace(momentum=u0),
                                       step_size, sqrt_diag_cov)
        u = partially_refresh_momentum(new.momentum, z[1], dt(step_size) * dt(0.5), L)
    new = new._replace(momentum=u)
    info = MCLMCInfo(new.logdensity, dk, (dk - new.logdensity + state.logdensity))
    return new, info


# --------------------------------------------------------------------------------------
# a9-a11: step-size / L tuning  (src/training/warmup.py:155-483)
# --------------------------------------------------------------------------------------

@dataclass
class TuneConfig:
    """Arguments of custom_mclmc_warmup (warmup.py:486-495) + phase lengths (l.543,555-557)."""
    tune1: int
    tune2: int
    tune3: int
    desired_energy_var_start: float = 5e-4
    desired_energy_var_end: float = 5e-4
    trust_in_estimate: float = 1.5
    num_effective_samples: int = 100
    step_size_init: float = 0.005
    diagonal_preconditioning: bool = False

    @classmethod
    def from_warmup_steps(cls, num_steps: int, **kw):
        return cls(int(num_steps * 0.8), int(num_steps * 0.1), int(num_steps * 0.1), **kw)


def desired_energy_var(cfg: TuneConfig, step: int, dt=np.float32):
    """warmup.py:249-274 (linear schedule; exponential if start > 2)."""
    total = cfg.tune1 + cfg.tune2 + 1
    s, e = dt(cfg.desired_energy_var_start), dt(cfg.desired_energy_var_end)
    if cfg.desired_energy_var_start > 2.0:
        tau = dt(total / 4)
        ex = np.exp(-dt(step) / tau)
        return s * ex + e * (dt(1) - ex)
    progress = min(dt(step) / dt(total), dt(1))
    return s - (s - e) * progress


def handle_nans(prev: IntegratorState, new: IntegratorState, step_size, step_size_max, energy_change):
    """warmup.py:468-483."""
    dt = prev.position.dtype.type
    nonans = bool(np.all(np.isfinite(new.position)))
    if nonans:
        state = IntegratorState(*[np.nan_to_num(np.asarray(v)) for v in new])
        return True, state, dt(np.nan_to_num(step_size_max)), dt(np.nan_to_num(energy_change))
    return False, prev, dt(step_size) * dt(0.8), dt(0.0)


class TuneState(NamedTuple):
    time: np.floating
    x_average: np.floating
    step_size_max: np.floating
    step_size: np.floating
    L: np.floating
    w_total: np.floating  # streaming_avg weight
    avg_x: np.ndarray  # E[x]
    avg_x2: np.ndarray  # E[x^2]


def tune_init(cfg: TuneConfig, d: int, dt=np.float32) -> TuneState:
    """warmup.py:204-209,358-363."""
    return TuneState(dt(0), dt(0), dt(np.inf), dt(cfg.step_size_init),
                     dt(max(math.sqrt(d), 15.0)), dt(0), np.zeros(d, dt), np.zeros(d, dt))


def tune_update(cfg: TuneConfig, ts: TuneState, x: np.ndarray, de, success: bool, step_number: int):
    """The adaptive part of one HOT LOOP A iteration, after the dynamics and handle_nans:
    energy-variance step-size predictor (warmup.py:302-322) and the streaming average of
    (x, x^2) (warmup.py:341-348, blackjax.util.streaming_average_update).  `ts.step_size_max`
    must already hold the value handle_nans returned."""
    dt = x.dtype.type
    d = x.shape[0]
    decay = dt((cfg.num_effective_samples - 1.0) / (cfg.num_effective_samples + 1.0))
    target = desired_energy_var(cfg, step_number, dt)
    step_size_max = ts.step_size_max
    with np.errstate(divide='ignore', over='ignore', invalid='ignore'):
        xi = np.square(dt(de)) / (dt(d) * target) + dt(1e-8)
        weight = np.exp(dt(-0.5) * np.square(np.log(xi) / dt(6.0 * cfg.trust_in_estimate)))
        x_average = decay * ts.x_average + weight * (xi / np.power(ts.step_size, dt(6.0)))
        time = decay * ts.time + weight
        step_size = np.power(x_average / time, dt(-1.0 / 6.0))
    # warmup.py:318-320 -- yields 0 when step_size == step_size_max (kept bit-for-bit)
    step_size = dt(step_size < step_size_max) * step_size + dt(step_size > step_size_max) * step_size_max
    mask = 1.0 if step_number < cfg.tune1 else 0.0  # warmup.py:376
    w = dt((1.0 - mask) * float(success)) * step_size
    denom = ts.w_total + w + dt(mask)
    with np.errstate(divide='ignore', invalid='ignore'):
        avg_x = (ts.w_total * ts.avg_x + w * x) / denom
        avg_x2 = (ts.w_total * ts.avg_x2 + w * np.square(x)) / denom
    return TuneState(dt(time), dt(x_average), dt(step_size_max), dt(step_size), ts.L,
                     dt(ts.w_total + w), avg_x.astype(x.dtype), avg_x2.astype(x.dtype))


def tune_step(logdensity_and_grad, cfg: TuneConfig, state: IntegratorState, ts: TuneState,
              z, step_number: int, refresh='post', sqrt_diag_cov=1.0):
    """One iteration of HOT LOOP A: warmup.py:276-352 (`predictor` + `step`)."""
    new, info = mclmc_step(logdensity_and_grad, state, ts.step_size, ts.L, z, sqrt_diag_cov=sqrt_diag_cov, refresh=refresh)
    success, state, step_size_max, de = handle_nans(state, new, ts.step_size, ts.step_size_max,
                                                    info.energy_change)
    ts = tune_update(cfg, ts._replace(step_size_max=step_size_max), state.position, de, success, step_number)
    return state, ts, info._replace(energy_change=de), success


def tune_finish_phase2(cfg: TuneConfig, ts: TuneState) -> TuneState:
    """warmup.py:383-390: L = sqrt(sum(E[x^2]-E[x]^2)) when tune2 != 0 (no diag precond)."""
    if cfg.tune2 != 0:
        var = ts.avg_x2 - np.square(ts.avg_x)
        return ts._replace(L=np.sqrt(np.sum(var, dtype=var.dtype)))
    return ts


def next_fast_len(n: int) -> int:
    """scipy.fft.next_fast_len for real transforms of 5-smooth... blackjax uses
    `scipy.fftpack.next_fast_len` (2,3,5-smooth)."""
    while True:
        m = n
        for p in (2, 3, 5):
            while m % p == 0:
                m //= p
        if m == 1:
            return n
        n += 1


def effective_sample_size(x: np.ndarray) -> np.ndarray:
    """blackjax.diagnostics.effective_sample_size (== numpyro's): x [chains, samples, dim].

    FFT autocovariance + Geyer initial positive / initial monotone sequence.
    """
    x = np.asarray(x)
    dt = x.dtype
    n_chains, n = x.shape[0], x.shape[1]
    assert n > 1
    mean = x.mean(axis=1, keepdims=True)
    cen = x - mean
    m = next_fast_len(2 * n)
    f = np.fft.rfft(cen, n=m, axis=1)
    f = f * np.conjugate(f)
    acov = np.fft.irfft(f, n=m, axis=1)[:, :n].astype(dt) / dt.type(n)
    mean_acov = acov.mean(axis=0, keepdims=True)
    mean_var0 = mean_acov[:, :1] * dt.type(n) / dt.type(n - 1.0)
    weighted_var = mean_var0 * dt.type(n - 1.0) / dt.type(n)
    if n_chains > 1:
        weighted_var = weighted_var + mean.var(axis=0, ddof=1, keepdims=True)
    n_even = n - n % 2
    tp1 = mean_acov[:, 1:n_even]
    rho = np.concatenate([np.ones_like(mean_var0), dt.type(1) - (mean_var0 - tp1) / weighted_var],
                         axis=1)
    rho = np.moveaxis(rho, 1, 0)  # [lag, 1, dim]
    rho_even, rho_odd = rho[0::2].copy(), rho[1::2].copy()
    mask0 = (rho_even + rho_odd) > 0
    T = mask0.shape[0]
    mask = np.logical_and.accumulate(mask0, axis=0)
    # max_t = last t where the running mask is still true (0 if none)
    idx = np.arange(T).reshape((T,) + (1,) * (mask.ndim - 1))
    max_t = np.where(mask, idx, 0).max(axis=0)
    rho_odd = np.where(mask, rho_odd, 0)
    sel = max_t + 1  # "improve estimation": keep even term at max_t+1 if positive
    mask_even = mask.copy()
    it = np.ndindex(*max_t.shape)
    for ix in it:
        t = int(sel[ix])
        if t < T:  # JAX drops out-of-bounds scatter updates
            mask_even[(t,) + ix] = rho_even[(t,) + ix] > 0
    rho_even = np.where(mask_even, rho_even, 0)
    rho_sum = rho_even + rho_odd
    upd_mask = np.zeros_like(mask)
    upd_val = np.zeros_like(rho_sum)
    prev = rho_sum[0]
    for t in range(T):
        um = rho_sum[t] > prev
        nxt = np.where(um, prev, rho_sum[t])
        upd_mask[t], upd_val[t] = um, nxt
        prev = nxt
    rho_even_f = np.where(upd_mask, upd_val / dt.type(2), rho_even)
    rho_odd_f = np.where(upd_mask, upd_val / dt.type(2), rho_odd)
    ess_raw = n_chains * n
    sel_c = np.minimum(sel, T - 1)  # JAX clamps out-of-bounds gathers
    last = np.take_along_axis(rho_even_f, sel_c[None], axis=0)[0]
    tau = dt.type(-1) + dt.type(2) * np.sum(rho_even_f + rho_odd_f, axis=0) - last
    tau = np.maximum(tau, dt.type(1 / np.log10(ess_raw)))
    return (dt.type(ess_raw) / tau).squeeze()


def effective_sample_size_direct(x: np.ndarray) -> float:
    """The same estimator for ONE series [n], written the way csrc/mile_ess.cuh computes it: direct autocovariance lag by lag,
    stopping at the first non-positive pair sum (Geyer's initial positive sequence never reads beyond it), then the
    monotone-sequence pass over the kept pairs.  A second opinion for `effective_sample_size` (FFT, all lags) and the
    statement of why the kernel may stop early."""
    x = np.asarray(x)
    dt = x.dtype.type
    n = x.shape[0]
    n_even = n - (n & 1)
    T = n_even // 2
    cen = x - x.mean(dtype=x.dtype)
    fn = dt(n)
    ac0 = np.dot(cen, cen) / fn
    var0 = ac0 * fn / (fn - dt(1))
    wvar = var0 * (fn - dt(1)) / fn
    rho_cache = {0: dt(1)}

    def rho(t):
        if t not in rho_cache:
            rho_cache[t] = dt(1) - (var0 - np.dot(cen[:n - t], cen[t:]) / fn) / wvar
        return rho_cache[t]
    n_pos = T
    for k in range(T):
        if not (rho(2 * k) + rho(2 * k + 1) > 0):
            n_pos = k
            break
    sel = (n_pos - 1 if n_pos > 0 else 0) + 1
    last_k = sel if sel < T else T - 1
    run_min, total, last_even = dt(0), dt(0), dt(0)
    for k in range(last_k + 1):
        if k < n_pos:
            re, ro = rho(2 * k), rho(2 * k + 1)
        elif k == sel:
            r = rho(2 * k)
            re, ro = (r if r > 0 else dt(0)), dt(0)
        else:
            re, ro = dt(0), dt(0)
        sm = re + ro
        run_min = sm if k == 0 else min(run_min, sm)
        if sm > run_min:
            re = ro = run_min / dt(2)
        total += re + ro
        if k == last_k:
            last_even = re
    tau = max(dt(-1) + dt(2) * total - last_even, dt(1) / dt(np.log10(fn)))
    return float(fn / tau), len(rho_cache)


def adaptation_L(step_size, samples: np.ndarray, Lfactor=0.4):
    """warmup.py:442-463: L = Lfactor * eps * mean(n / ESS) over the tune3 positions
    ([tune3, d]; the >2000-parameter / >10000-sample subsampling is applied by the caller)."""
    dt = samples.dtype.type
    n = samples.shape[0]
    ess = effective_sample_size(samples[None])
    return dt(Lfactor) * dt(step_size) * np.mean(dt(n) / ess, dtype=samples.dtype)


def subsample_for_fft(samples: np.ndarray, perm: np.ndarray | None,
                      fft_params_limit=2000, fft_samples_limit=10000):
    """warmup.py:442-456.  `perm` replaces jax.random.permutation(key, arange(d))."""
    if samples.shape[1] > fft_params_limit:
        samples = samples[:, perm[:fft_params_limit]]
    if samples.shape[0] > fft_samples_limit:
        idx = np.linspace(0, samples.shape[0] - 1, fft_samples_limit).astype(np.int32)
        samples = samples[idx]
    return samples


def run_warmup(logdensity_and_grad, cfg: TuneConfig, position, z0, z_steps, refresh='post',
               diagonal_preconditioning=False):
    """custom_mclmc_warmup(...).run for one chain (warmup.py:533-566).

    z_steps: [tune1+tune2 (+ tune2//3 with diagonal_preconditioning) + tune3, d] normal draws (host supplied).
    Returns (state, step_size, L, TuneState) and, with diagonal_preconditioning, sqrt_diag_cov as a fifth value.
    """
    state = mclmc_init(logdensity_and_grad, position, z0)
    d = position.shape[0]
    dt = position.dtype.type
    ts = tune_init(cfg, d, dt)
    k = 0
    for i in range(cfg.tune1 + cfg.tune2):
        state, ts, _, _ = tune_step(logdensity_and_grad, cfg, state, ts, z_steps[k], i, refresh)
        k += 1
    ts = tune_finish_phase2(cfg, ts)
    sdc = 1.0
    if diagonal_preconditioning and cfg.tune2 != 0:                      # warmup.py:391-401
        var = ts.avg_x2 - np.square(ts.avg_x)
        sdc = np.sqrt(np.maximum(var, 0)).astype(position.dtype)
        ts2 = tune_init(cfg, d, dt)._replace(step_size=ts.step_size, L=dt(np.sqrt(dt(d))))
        for i in range(cfg.tune2 // 3):                                  # run_steps restarts adaptive state and counter
            state, ts2, _, _ = tune_step(logdensity_and_grad, cfg, state, ts2, z_steps[k], i, refresh, sqrt_diag_cov=sdc)
            k += 1
        ts = ts2
    L, eps = ts.L, ts.step_size
    if cfg.tune3 != 0:
        pos = np.empty((cfg.tune3, d), position.dtype)
        for i in range(cfg.tune3):
            state, _ = mclmc_step(logdensity_and_grad, state, eps, L, z_steps[k], sqrt_diag_cov=sdc, refresh=refresh)
            pos[i] = state.position
            k += 1
        L = adaptation_L(eps, pos)
    if diagonal_preconditioning:
        return state, eps, L, ts, sdc
    return state, eps, L, ts


# --------------------------------------------------------------------------------------
# a6: sampling loop structure (src/training/sampling.py:134-177)
# --------------------------------------------------------------------------------------

def run_sampling(logdensity_and_grad, state, step_size, L, z_steps, n_thinning=1, refresh='post'):
    """scan(sampler.step) over n_samples with the thinning predicate idx % n_thinning == 0
    (sampling.py:150-164).  Returns final state, kept positions [n_kept,d], kept indices,
    per-step energy changes."""
    kept, idxs, des = [], [], []
    for i in range(z_steps.shape[0]):
        state, info = mclmc_step(logdensity_and_grad, state, step_size, L, z_steps[i], refresh=refresh)
        des.append(info.energy_change)
        if i % n_thinning == 0:
            kept.append(state.position.copy())
            idxs.append(i)
    return state, np.stack(kept) if kept else np.empty((0,) + state.position.shape), idxs, np.array(des)


# --------------------------------------------------------------------------------------
# a15: posterior-predictive LPPD (src/inference/metrics.py:247-312,428-446)
# --------------------------------------------------------------------------------------

def pointwise_lppd(spec: ModelSpec, lvals: np.ndarray, y: np.ndarray):
    """metrics.py:247-293: Normal(mu, clip(exp(s))).log_prob(y) / Categorical(logits).log_prob(y).
    lvals [..., n_obs, K] -> [..., n_obs].  (numpyro Normal.log_prob:
    -((y-mu)^2)/(2 s^2) - log(s) - log(sqrt(2 pi)).)"""
    dt = lvals.dtype.type
    if spec.task == 'regr':
        mu = lvals[..., 0]
        sigma = np.clip(np.exp(lvals[..., 1]), dt(1e-6), dt(1e6))
        return (-np.square(y.astype(lvals.dtype) - mu) / (dt(2) * sigma * sigma)
                - np.log(sigma) - dt(math.log(math.sqrt(2 * math.pi))))
    m = lvals.max(axis=-1, keepdims=True)
    lse = m[..., 0] + np.log(np.exp(lvals - m).sum(axis=-1))
    yi = np.broadcast_to(y.astype(np.int64), lvals.shape[:-1])
    return np.take_along_axis(lvals, yi[..., None], axis=-1)[..., 0] - lse


def lppd(lp_pointwise: np.ndarray):
    """metrics.py:296-312: mean_n logsumexp_{c,s}(lp, b=1/(C*S))."""
    dt = lp_pointwise.dtype.type
    flat = lp_pointwise.reshape(-1, lp_pointwise.shape[-1])
    m = flat.max(axis=0)
    m = np.where(np.isfinite(m), m, dt(0))
    s = np.exp(flat - m).sum(axis=0) / dt(flat.shape[0])
    return (m + np.log(s)).mean()


def running_lppd(lp_pointwise: np.ndarray):
    """metrics.py:428-446: log(running_mean(exp(lp), axis=-2)).mean(-1).mean(0)."""
    e = np.exp(lp_pointwise)
    cs = np.cumsum(e, axis=-2)
    cnt = np.arange(1, e.shape[-2] + 1).reshape((1, -1, 1))
    return np.log(cs / cnt).mean(axis=-1).mean(axis=0)


def online_logsumexp_update(m: np.ndarray, s: np.ndarray, lp: np.ndarray):
    """Streaming (max, sum exp) state used by the fused LPPD accumulation: the CUDA path
    keeps (m_n, s_n) per chain and test point and merges them at the end."""
    new_m = np.maximum(m, lp)
    safe = np.where(np.isfinite(new_m), new_m, 0)
    s = s * np.exp(np.where(np.isfinite(m), m, -np.inf) - safe) + np.exp(lp - safe)
    return new_m, s


def lppd_from_state(m: np.ndarray, s: np.ndarray, count: int):
    """Merge per-chain (m, s) [C, Nt] -> LPPD scalar; count = C*S."""
    M = m.max(axis=0)
    safe = np.where(np.isfinite(M), M, 0)
    tot = (s * np.exp(m - safe)).sum(axis=0)
    return (safe + np.log(tot / count)).mean()


# --------------------------------------------------------------------------------------
# Diagnostics used as parity metrics (src/inference/metrics.py:226-244,354-405,449-523)
# --------------------------------------------------------------------------------------

def rank_normalize_array(samples: np.ndarray):
    """metrics.py:226-244: overall ranks (average ties) -> (r-0.375)/(n+0.25) -> norm.ppf."""
    from scipy.stats import norm, rankdata
    n = samples.size
    ranks = rankdata(samples, axis=None).reshape(samples.shape)
    return norm.ppf((ranks - 0.375) / (n + 0.25))


def _rank_normalize_per_param(x: np.ndarray):
    """`jnp.apply_along_axis(rank_normalize_array, 0, x.reshape(-1, *x.shape[2:]))`."""
    flat = x.reshape((-1,) + x.shape[2:])
    return np.apply_along_axis(rank_normalize_array, 0, flat).reshape(x.shape)


def between_chain_var(x: np.ndarray):
    """metrics.py:354-367."""
    return x.mean(axis=1).var(axis=0, ddof=1)


def within_chain_var(x: np.ndarray):
    """metrics.py:370-383."""
    return x.var(axis=1, ddof=1).mean(axis=0)


def gelman_split_r_hat(samples: np.ndarray, n_splits: int, rank_normalize: bool = True):
    """metrics.py:449-497."""
    n_chains = samples.shape[0]
    n_samples = samples.shape[1] / n_splits
    if (n_samples % 1) != 0:
        raise ValueError('Number of samples must be divisible by n_splits')
    if rank_normalize:
        samples = _rank_normalize_per_param(samples)
    splits = samples.reshape((n_chains * n_splits, -1) + samples.shape[2:])
    wcv = within_chain_var(splits)
    bcv = between_chain_var(splits)
    return np.sqrt((((n_samples - 1) / n_samples) * wcv + bcv) / wcv)


def split_chain_r_hat(samples: np.ndarray, n_splits: int = 4, rank_normalize: bool = True):
    """metrics.py:500-523: gelman_split_r_hat per chain -> [C, ...]."""
    return np.stack([gelman_split_r_hat(chain[None], n_splits, rank_normalize) for chain in samples])


def ess_rank_normalized(x: np.ndarray, rank_normalize: bool = True):
    """metrics.py:386-405: per-chain numpyro ESS on (optionally) rank-normalised samples."""
    if rank_normalize:
        x = _rank_normalize_per_param(x)
    return np.stack([effective_sample_size(c[None]) for c in x])


# --------------------------------------------------------------------------------------
# Synthetic inputs of the BASELINE shapes
# --------------------------------------------------------------------------------------

# Synthetic workloads: one definition for bench.py, smoke() and the tests lives in mile_b200/synthetic.py (numpy only);
# the oracle re-exports it under its historical names.
from mile_b200.synthetic import CONFIGS, synthetic_data  # noqa: E402,F401
from mile_b200 import synthetic as _syn  # noqa: E402


def make_spec(name: str) -> ModelSpec:
    _, _, F, widths, act, task = CONFIGS[name]
    return ModelSpec(F, widths, act, task)


def synthetic_theta0(spec: ModelSpec, n_chains: int, scale: float = 0.3, seed0: int = 1000):
    """Warm-start stand-in: theta0_c ~ N(0, 0.3^2) with chain seed 1000+c."""
    return _syn.synthetic_theta0(spec.n_params, n_chains, scale, seed0)
