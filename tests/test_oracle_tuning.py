"""Pins for the tuning restatement (src/training/warmup.py:155-483) and ESS."""
import math

import numpy as np

from oracle import mile_oracle as o
from tests.test_oracle_mclmc import gauss


def test_handle_nans_paths():
    st = o.IntegratorState(np.ones(3), np.ones(3), np.float64(1.0), np.ones(3))
    bad = o.IntegratorState(np.array([1.0, np.nan, 2.0]), np.ones(3), np.float64(2.0), np.ones(3))
    ok, s, smax, de = o.handle_nans(st, bad, 0.5, np.inf, 3.0)
    assert not ok and s is st and smax == 0.4 and de == 0.0
    good = o.IntegratorState(np.full(3, 2.0), np.ones(3), np.float64(np.nan), np.ones(3))
    ok, s, smax, de = o.handle_nans(st, good, 0.5, np.float64(0.7), 3.0)
    assert ok and smax == 0.7 and de == 3.0 and s.logdensity == 0.0  # nan_to_num(new)


def test_desired_energy_var_schedules():
    cfg = o.TuneConfig(80, 10, 10, 0.5, 0.1)
    assert o.desired_energy_var(cfg, 0, np.float64) == 0.5
    assert abs(o.desired_energy_var(cfg, 91, np.float64) - 0.1) < 1e-12
    mid = o.desired_energy_var(cfg, 45, np.float64)
    assert abs(mid - (0.5 - 0.4 * 45 / 91)) < 1e-12
    cfg2 = o.TuneConfig(80, 10, 10, 4.0, 0.1)
    v = o.desired_energy_var(cfg2, 10, np.float64)
    tau = 91 / 4
    assert abs(v - (4.0 * math.exp(-10 / tau) + 0.1 * (1 - math.exp(-10 / tau)))) < 1e-12


def test_step_size_predictor_hand_computed():
    """Scripted dE -> hand-computed eps trajectory (warmup.py:303-320)."""
    cfg = o.TuneConfig(5, 0, 0, 0.5, 0.5, trust_in_estimate=1.5, num_effective_samples=100,
                       step_size_init=0.1)
    d = 10
    ts = o.tune_init(cfg, d, np.float64)
    gamma = 99.0 / 101.0
    time = xavg = 0.0
    eps = 0.1
    rng = np.random.default_rng(0)
    st = o.mclmc_init(gauss, rng.standard_normal(d), rng.standard_normal(d))
    for i in range(5):
        z = rng.standard_normal(d)
        st_new, ts, info, ok = o.tune_step(gauss, cfg, st, ts, z, i)
        de = info.energy_change
        xi = de ** 2 / (d * 0.5) + 1e-8
        w = math.exp(-0.5 * (math.log(xi) / (6 * 1.5)) ** 2)
        xavg = gamma * xavg + w * xi / eps ** 6
        time = gamma * time + w
        eps = (xavg / time) ** (-1 / 6)
        assert abs(ts.step_size - eps) < 1e-12 * eps
        st = st_new
    assert ts.w_total == 0  # phase 1: mask = 1 -> weight 0


def test_streaming_average_and_L_from_variance():
    """Phase 2 accumulates eps-weighted E[x], E[x^2]; L = sqrt(sum var) (warmup.py:341-390)."""
    cfg = o.TuneConfig(0, 300, 0, 0.5, 0.5, step_size_init=0.3)
    d = 20
    rng = np.random.default_rng(1)
    st = o.mclmc_init(gauss, rng.standard_normal(d), rng.standard_normal(d))
    ts = o.tune_init(cfg, d, np.float64)
    xs, ws = [], []
    for i in range(300):
        st, ts, _, ok = o.tune_step(gauss, cfg, st, ts, rng.standard_normal(d), i)
        xs.append(st.position)
        ws.append(ts.step_size)
    xs, ws = np.array(xs), np.array(ws)
    np.testing.assert_allclose(ts.avg_x, (ws[:, None] * xs).sum(0) / ws.sum(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(ts.avg_x2, (ws[:, None] * xs ** 2).sum(0) / ws.sum(), rtol=1e-9)
    ts2 = o.tune_finish_phase2(cfg, ts)
    assert abs(ts2.L - math.sqrt((ts.avg_x2 - ts.avg_x ** 2).sum())) < 1e-12
    assert 2.0 < ts2.L < 7.0  # ~ sqrt(d) for a unit Gaussian


def test_nan_injection_shrinks_step_size_max():
    cfg = o.TuneConfig(3, 0, 0, 0.5, 0.5, step_size_init=0.2)
    d = 4
    calls = {'n': 0}

    def bad(theta):
        calls['n'] += 1
        lp, g = gauss(theta)
        if calls['n'] in (2, 3):  # poison the first tuning step's gradients
            g = g * np.nan
        return lp, g

    rng = np.random.default_rng(2)
    st = o.mclmc_init(bad, rng.standard_normal(d), rng.standard_normal(d))
    ts = o.tune_init(cfg, d, np.float64)
    st2, ts2, info, ok = o.tune_step(bad, cfg, st, ts, rng.standard_normal(d), 0)
    assert not ok
    assert st2 is st
    assert ts2.step_size_max == 0.2 * 0.8
    assert ts2.step_size <= ts2.step_size_max


def test_ess_iid_and_ar1():
    rng = np.random.default_rng(3)
    n = 4000
    iid = rng.standard_normal((1, n, 3))
    ess = o.effective_sample_size(iid)
    assert np.all(ess > 0.7 * n) and np.all(ess < 1.4 * n)
    phi = 0.9
    x = np.zeros((1, n, 2))
    e = rng.standard_normal((n, 2))
    for t in range(1, n):
        x[0, t] = phi * x[0, t - 1] + e[t]
    ess = o.effective_sample_size(x)
    expect = n * (1 - phi) / (1 + phi)
    assert np.all(ess > 0.5 * expect) and np.all(ess < 2.0 * expect)


def test_adaptation_L_formula():
    rng = np.random.default_rng(4)
    pos = rng.standard_normal((500, 6))
    L = o.adaptation_L(0.05, pos)
    ess = o.effective_sample_size(pos[None])
    assert abs(L - 0.4 * 0.05 * np.mean(500 / ess)) < 1e-12


def test_run_warmup_gaussian_end_to_end():
    """Full warmup on a 50-d Gaussian: the tuned eps reproduces the desired energy variance
    (Var[dE]/d ~ target) -- the self-consistency the predictor is built on."""
    d = 50
    cfg = o.TuneConfig.from_warmup_steps(1000, desired_energy_var_start=5e-4, desired_energy_var_end=5e-4,
                                         step_size_init=0.01)
    assert (cfg.tune1, cfg.tune2, cfg.tune3) == (800, 100, 100)
    rng = np.random.default_rng(5)
    st, eps, L, ts = o.run_warmup(gauss, cfg, rng.standard_normal(d), rng.standard_normal(d),
                                  rng.standard_normal((1000, d)))
    assert np.isfinite(eps) and np.isfinite(L) and 0.1 < L < 100
    des = []
    for _ in range(1000):
        st, info = o.mclmc_step(gauss, st, eps, L, rng.standard_normal(d))
        des.append(info.energy_change)
    assert 0.5 * 5e-4 < np.var(des) / d < 2 * 5e-4


def test_lazy_direct_ess_equals_the_fft_estimator():
    """The estimator as the device kernel computes it (csrc/mile_ess.cuh: direct autocovariance, lags produced only up to the
    first non-positive pair) against the restatement of blackjax.diagnostics.effective_sample_size (FFT, all lags), fp64 and
    fp32, even / odd lengths, strongly and negatively correlated series; and the early stop really is early."""
    rng = np.random.default_rng(0)
    for n in (600, 601, 1500):
        for phi in (0.0, 0.3, 0.9, 0.97, 0.995, -0.5):
            x = np.zeros(n)
            e = rng.standard_normal(n)
            for t in range(1, n):
                x[t] = phi * x[t - 1] + e[t]
            want = float(o.effective_sample_size(x[None, :, None]))
            got, lags = o.effective_sample_size_direct(x)
            assert abs(got - want) <= 1e-9 * want, (n, phi, got, want)
            got32, _ = o.effective_sample_size_direct(x.astype(np.float32))
            assert abs(got32 - want) <= 1e-4 * want, (n, phi, got32, want)
            if abs(phi) <= 0.9:
                assert lags < n // 2, (n, phi, lags)      # far fewer than the n lags an FFT produces
