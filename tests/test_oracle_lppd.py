"""Pins for LPPD / diagnostics restatement (src/inference/metrics.py)."""
import numpy as np
from scipy import special, stats

from oracle import mile_oracle as o


def test_pointwise_lppd_vs_scipy():
    rng = np.random.default_rng(0)
    spec = o.ModelSpec(3, (4, 2), 'relu', 'regr')
    lv = rng.standard_normal((2, 5, 7, 2))
    y = rng.standard_normal(7)
    ref = stats.norm.logpdf(y, lv[..., 0], np.clip(np.exp(lv[..., 1]), 1e-6, 1e6))
    np.testing.assert_allclose(o.pointwise_lppd(spec, lv, y), ref, rtol=1e-12)
    specc = o.ModelSpec(3, (4, 5), 'relu', 'class')
    lc = rng.standard_normal((2, 5, 7, 5))
    yc = rng.integers(0, 5, 7)
    refc = special.log_softmax(lc, axis=-1)[..., np.arange(7), yc]
    np.testing.assert_allclose(o.pointwise_lppd(specc, lc, yc), refc, rtol=1e-12)


def test_lppd_bruteforce_and_streaming():
    rng = np.random.default_rng(1)
    lp = rng.standard_normal((3, 40, 11)) * 3
    brute = np.log(np.exp(lp).mean(axis=(0, 1))).mean()
    assert abs(o.lppd(lp) - brute) < 1e-12
    m = np.full((3, 11), -np.inf)
    s = np.zeros((3, 11))
    for t in range(40):
        m, s = o.online_logsumexp_update(m, s, lp[:, t])
    assert abs(o.lppd_from_state(m, s, 3 * 40) - brute) < 1e-12
    run = o.running_lppd(lp)
    assert run.shape == (40,)
    assert abs(run[-1] - np.log(np.exp(lp).mean(axis=1)).mean(-1).mean(0)) < 1e-12


def test_split_rhat_and_ess_diagnostics():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((3, 400, 2))
    r = o.split_chain_r_hat(x, 4)
    assert r.shape == (3, 2) and np.all(np.abs(r - 1) < 0.05)
    x[1, 200:] += 3.0  # a chain that moved: its split R-hat must flag it
    r = o.split_chain_r_hat(x, 4)
    assert np.all(r[1] > 1.3) and np.all(np.abs(r[0] - 1) < 0.05)
    ess = o.ess_rank_normalized(rng.standard_normal((2, 500, 2)))
    assert ess.shape == (2, 2) and np.all(ess > 250)
    b, w = o.between_chain_var(x), o.within_chain_var(x)
    assert b.shape == (2,) and w.shape == (2,)
