"""Posterior-predictive evaluation mirror (src/inference/evaluation.py:46-137,334-544).
CPU part: the host-side reductions.  GPU part: predict_bde / evaluate_bde against the oracle's forward pass."""
import numpy as np
import pytest

from oracle import mile_oracle as o


def test_coverage_helpers():
    from mile_b200 import evaluation as ev
    assert np.allclose(ev.get_quantiles(0.9), [0.05, 0.95])
    w = ev.coverage_weighting([0.5, 0.9], kappa=2.0)
    assert np.isclose(w.sum(), 1.0) and w[1] > w[0]
    assert np.isclose(ev.calibration_error([0.5, 0.9], [0.4, 0.9]), np.sqrt(0.01 / 2))
    rng = np.random.default_rng(0)
    preds = rng.standard_normal((4, 500, 64))           # (chains, samples, N) predictive draws ~ N(0,1)
    cov = ev.calculate_coverage([0.5, 0.9], np.zeros(64), preds)
    assert np.all(cov == 1.0)                           # 0 lies inside every central interval
    y = rng.standard_normal(4000)
    cov = ev.calculate_coverage([0.5, 0.9], y, rng.standard_normal((2, 400, 4000)))
    assert abs(cov[0] - 0.5) < 0.05 and abs(cov[1] - 0.9) < 0.03


def test_sample_from_predictions_distribution():
    from mile_b200 import evaluation as ev
    lv = np.zeros((2, 2000, 3, 2), np.float32)
    lv[..., 0] = 1.5
    lv[..., 1] = np.log(0.5)
    s = ev.sample_from_predictions(lv, 'regr', rng_key=7)
    assert s.shape == (2, 2000, 3) and abs(s.mean() - 1.5) < 0.03 and abs(s.std() - 0.5) < 0.03
    logits = np.log(np.array([0.7, 0.2, 0.1], np.float32)) * np.ones((1, 20000, 1, 3), np.float32)
    c = ev.sample_from_predictions(logits, 'class', rng_key=np.array([0, 7], np.uint32))
    freq = np.bincount(c.ravel(), minlength=3) / c.size
    assert np.allclose(freq, [0.7, 0.2, 0.1], atol=0.02)
    assert np.array_equal(ev._mode(np.array([[0, 2], [2, 1], [2, 1]]), 3, (0,)), [2, 1])


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['airfoil_3x16', 'covertype_ref'])
def test_evaluate_bde_matches_oracle(name):
    from mile_b200 import FCN, evaluation as ev, metrics
    from mile_b200.engine import FCNSpec
    ospec = o.make_spec(name)
    _, _, Xt, yt = o.synthetic_data(name, n_test=300)
    C, S, d = 3, 5, ospec.n_params
    rng = np.random.default_rng(5)
    theta = (0.3 * rng.standard_normal((C, S, d))).astype(np.float32)
    spec = FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task)
    params = spec.unravel(theta)
    module = FCN(ospec.widths, ospec.activation)
    lv = ev.predict_bde(params, module, Xt, yt, None, verbose=False, task=ospec.task, chunk=4)
    ref = np.stack([[o.forward(ospec, theta[c, s].astype(np.float64), Xt.astype(np.float64)) for s in range(S)] for c in range(C)])
    assert lv.shape == ref.shape
    assert np.max(np.abs(lv - ref)) <= 1e-5 * max(1.0, np.abs(ref).max())
    one = ev.predict_from_samples(module, spec.unravel(theta[0]), Xt, task=ospec.task)
    assert np.array_equal(one, lv[0])
    _, md = ev.evaluate_bde(params, module, Xt, yt, ospec.task, None, verbose=False, nominal_coverages=[0.5, 0.9])
    pw = o.pointwise_lppd(ospec, ref, yt)
    assert abs(md['lppd'] - o.lppd(pw)) <= 1e-5 * max(1.0, abs(o.lppd(pw)))
    assert abs(md['nll'] + pw.mean()) <= 1e-5 * max(1.0, abs(pw.mean()))
    assert ('rmse' in md and 'coverage_0.9' in md) if ospec.task == 'regr' else ('acc' in md and 0.0 <= md['acc'] <= 1.0)
