"""Pins for the blackjax-1.2.2 MCLMC restatement:
ESH invariants, energy bookkeeping, reversibility, Var[dE] ~ eps^6, Gaussian-target stats."""
import math

import numpy as np
import pytest

from oracle import mile_oracle as o


def gauss(theta):
    """d-dim standard normal target."""
    return (-0.5 * np.sum(theta * theta)).astype(theta.dtype), -theta


def _unit(rng, d, dt=np.float64):
    z = rng.standard_normal(d).astype(dt)
    return z / np.linalg.norm(z)


def test_esh_unit_norm_and_alignment_limits():
    rng = np.random.default_rng(0)
    d = 50
    g = rng.standard_normal(d) * 3
    gn = np.linalg.norm(g)
    e = g / gn
    u = _unit(rng, d)
    for eps in (1e-3, 0.1, 2.0):
        u2, dk = o.esh_momentum_update(u, g, eps, 0.5, 0.0)
        assert abs(np.linalg.norm(u2) - 1) < 1e-12
    # p = +1: u' = u, dK = +eps*coef*|g|
    u2, dk = o.esh_momentum_update(e, g, 0.3, 0.5, 0.0)
    np.testing.assert_allclose(u2, e, atol=1e-12)
    assert abs(dk - 0.3 * 0.5 * gn) < 1e-9
    # p = -1: u' = u, dK = -eps*coef*|g|   (ln(2 zeta^2) = ln 2 - 2 delta)
    u2, dk = o.esh_momentum_update(-e, g, 0.3, 0.5, 0.0)
    np.testing.assert_allclose(u2, -e, atol=1e-9)
    assert abs(dk + 0.3 * 0.5 * gn) < 1e-9
    # delta -> infinity: u' -> e
    u2, _ = o.esh_momentum_update(u, g, 1e4, 0.5, 0.0)
    np.testing.assert_allclose(u2, e, atol=1e-9)


def test_esh_first_order():
    rng = np.random.default_rng(1)
    d = 30
    g = rng.standard_normal(d)
    u = _unit(rng, d)
    eps, coef = 1e-5, 0.5
    gn = np.linalg.norm(g)
    e = g / gn
    p = u @ e
    delta = eps * coef * gn / (d - 1)
    u2, dk = o.esh_momentum_update(u, g, eps, coef, 0.0)
    np.testing.assert_allclose(u2, u + delta * (e - p * u), atol=10 * delta ** 2)
    assert abs(dk - eps * coef * (u @ g)) < 1e-3 * abs(eps * coef * (u @ g)) + 1e-12


def test_energy_change_bookkeeping_and_order():
    """energy_change = dK - l' + l; local error O(eps^3) for McLachlan."""
    rng = np.random.default_rng(2)
    d = 100
    theta = rng.standard_normal(d)
    st = o.mclmc_init(gauss, theta, rng.standard_normal(d))
    errs = []
    for eps in (0.4, 0.2, 0.1):
        new, info = o.mclmc_step(gauss, st, eps, np.inf, np.zeros(d))
        assert info.energy_change == info.kinetic_change - new.logdensity + st.logdensity
        errs.append(abs(info.energy_change))
    assert errs[0] / errs[1] > 5 and errs[1] / errs[2] > 5  # ~2^3 per halving


def test_deterministic_part_is_time_reversible():
    rng = np.random.default_rng(3)
    d = 40
    st = o.mclmc_init(gauss, rng.standard_normal(d), rng.standard_normal(d))
    fwd, _ = o.isokinetic_mclachlan(gauss, st, 0.3)
    back, _ = o.isokinetic_mclachlan(gauss, fwd._replace(momentum=-fwd.momentum), 0.3)
    np.testing.assert_allclose(back.position, st.position, atol=1e-10)
    np.testing.assert_allclose(-back.momentum, st.momentum, atol=1e-10)


def test_energy_variance_scales_as_eps6():
    """Var[dE]/d ~ eps^6 -- the relation the tuner relies on (warmup.py:315-317)."""
    rng = np.random.default_rng(4)
    d = 100

    def var_at(eps):
        st = o.mclmc_init(gauss, rng.standard_normal(d), rng.standard_normal(d))
        des = []
        for _ in range(400):
            st, info = o.mclmc_step(gauss, st, eps, 10.0, rng.standard_normal(d))
            des.append(info.energy_change)
        return np.var(des[50:])

    r = var_at(0.4) / var_at(0.2)
    assert 2 ** 4.5 < r < 2 ** 7.5


def test_refresh_is_unit_and_identity_for_infinite_L():
    rng = np.random.default_rng(5)
    u = _unit(rng, 20)
    z = rng.standard_normal(20)
    v = o.partially_refresh_momentum(u, z, 0.1, 3.0)
    assert abs(np.linalg.norm(v) - 1) < 1e-12
    nu = math.sqrt((math.exp(2 * 0.1 / 3.0) - 1) / 20)
    np.testing.assert_allclose(v, (u + nu * z) / np.linalg.norm(u + nu * z))
    assert o.partially_refresh_momentum(u, z, 0.1, np.inf) is u


def test_debug_ipynb_cell8_gaussian_2d():
    """debug.ipynb cell 8: 2-D standard normal, L=0.5... uses L=0.5, step_size=0.1, 1000 steps:
    (visual check in the reference) -> sample mean ~ 0, var ~ 1 within MC error."""
    rng = np.random.default_rng(6)
    st = o.mclmc_init(gauss, np.array([1.0, 1.0]), rng.standard_normal(2))
    xs = []
    for _ in range(20000):
        st, _ = o.mclmc_step(gauss, st, 0.1, 0.5, rng.standard_normal(2))
        xs.append(st.position)
    xs = np.array(xs)[1000:]
    assert np.all(np.abs(xs.mean(0)) < 0.1)
    assert np.all(np.abs(xs.var(0) - 1) < 0.15)


def test_maruyama_variant_switch():
    rng = np.random.default_rng(7)
    d = 10
    st = o.mclmc_init(gauss, rng.standard_normal(d), rng.standard_normal(d))
    z = rng.standard_normal((2, d))
    a, _ = o.mclmc_step(gauss, st, 0.1, 2.0, z, refresh='maruyama')
    b, _ = o.mclmc_step(gauss, st, 0.1, 2.0, z[1], refresh='post')
    assert abs(np.linalg.norm(a.momentum) - 1) < 1e-12
    assert not np.allclose(a.position, b.position)


def test_init_requires_two_dims_and_unit_momentum():
    with pytest.raises(ValueError):
        o.mclmc_init(gauss, np.zeros(1), np.ones(1))
    st = o.mclmc_init(gauss, np.ones(3), np.array([3.0, 0.0, 4.0]))
    np.testing.assert_allclose(st.momentum, [0.6, 0.0, 0.8])
    assert st.logdensity == -1.5


def test_fp32_step_tracks_fp64_twin():
    """Single-step fp32 vs fp64 twin on the airfoil log-posterior: the north-star 1e-5 band."""
    spec = o.make_spec('airfoil_3x16')
    X, y, _, _ = o.synthetic_data('airfoil_3x16')
    th = o.synthetic_theta0(spec, 1)[0]
    rng = np.random.default_rng(8)
    z0, z = rng.standard_normal(spec.n_params), rng.standard_normal(spec.n_params)
    f32 = lambda t: o.logpost_value_and_grad(spec, t, X, y)
    f64 = lambda t: o.logpost_value_and_grad(spec, t, X.astype(np.float64), y.astype(np.float64))
    s32 = o.mclmc_init(f32, th, z0.astype(np.float32))
    s64 = o.mclmc_init(f64, th.astype(np.float64), z0)
    n32, i32 = o.mclmc_step(f32, s32, 0.01, 26.0, z.astype(np.float32))
    n64, i64 = o.mclmc_step(f64, s64, 0.01, 26.0, z)
    assert n32.position.dtype == np.float32 and n32.momentum.dtype == np.float32
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert rel(n32.position, n64.position) < 1e-6
    assert rel(n32.momentum, n64.momentum) < 1e-5
    assert abs(n32.logdensity - n64.logdensity) < 1e-5 * abs(n64.logdensity)
