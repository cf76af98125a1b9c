"""The C restatement (oracle/mile_oracle.c, the CPU baseline) against the numpy oracle."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import mile_oracle as o


def rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize('name', ['airfoil_3x16', 'bikesharing_2x16', 'covertype_ref'])
def test_c_logpost_matches_numpy(name):
    spec = o.make_spec(name)
    X, y, _, _ = o.synthetic_data(name, n_train=700)
    th = o.synthetic_theta0(spec, 3)
    lp, g = c_oracle.logpost_batch(spec, th, X, y, threads=2)
    lp64, g64 = o.logpost_batch(spec, th.astype(np.float64), X.astype(np.float64), y)
    for c in range(3):
        assert abs(lp[c] - lp64[c]) <= 2e-5 * abs(lp64[c])
        assert rel(g[c], g64[c]) <= 2e-5


@pytest.mark.parametrize('act,task,prior', [('tanh', 'regr', 'laplace'), ('gelu', 'class', 'normal'),
                                            ('leaky_relu', 'regr', 'normal'), ('sigmoid', 'class', 'laplace')])
def test_c_logpost_variants(act, task, prior):
    K = 2 if task == 'regr' else 4
    spec = o.ModelSpec(6, (9, 5, K), act, task, prior, prior_loc=0.2, prior_scale=1.3)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((150, 6)).astype(np.float32)
    y = rng.standard_normal(150).astype(np.float32) if task == 'regr' else rng.integers(0, K, 150).astype(np.int32)
    th = (rng.standard_normal((2, spec.n_params)) * 0.4).astype(np.float32)
    lp, g = c_oracle.logpost_batch(spec, th, X, y)
    lp64, g64 = o.logpost_batch(spec, th.astype(np.float64), X.astype(np.float64), y)
    for c in range(2):
        assert abs(lp[c] - lp64[c]) <= 2e-5 * abs(lp64[c])
        assert rel(g[c], g64[c]) <= 2e-5


def test_c_steps_match_numpy_with_host_noise():
    name, C, n = 'airfoil_2x16', 2, 5
    spec = o.make_spec(name)
    X, y, _, _ = o.synthetic_data(name, n_train=400)
    th0 = o.synthetic_theta0(spec, C)
    rng = np.random.default_rng(1)
    z0 = rng.standard_normal((C, spec.n_params)).astype(np.float32)
    z = rng.standard_normal((n, C, spec.n_params)).astype(np.float32)
    ch = c_oracle.Chains(spec, X, y, th0, z0=z0, threads=2)
    samples, info = ch.sample(n, 0.02, 20.0, thin=2, z=z, info=True)
    assert samples.shape == (3, C, spec.n_params)
    f64 = lambda t: o.logpost_value_and_grad(spec, t, X.astype(np.float64), y)
    for c in range(C):
        st = o.mclmc_init(f64, th0[c].astype(np.float64), z0[c].astype(np.float64))
        st, kept, idxs, des = o.run_sampling(f64, st, 0.02, 20.0, z[:, c].astype(np.float64), n_thinning=2)
        assert idxs == [0, 2, 4]
        assert rel(ch.theta[c], st.position) <= 2e-5
        assert rel(ch.u[c], st.momentum) <= 2e-4
        for k in range(3):
            assert rel(samples[k, c], kept[k]) <= 2e-5
        assert np.max(np.abs(info[:, c, 2] - des)) <= 1e-4 * abs(st.logdensity)


def test_c_philox_is_standard_normal():
    lib = c_oracle.load()
    xs = np.array([lib.mo_philox_normal(7, 1, s, 1, e) for s in range(50) for e in range(200)])
    assert abs(xs.mean()) < 0.03 and abs(xs.var() - 1) < 0.05
