"""Pins for the oracle's log-posterior: torch.autograd (fp64+fp32),
scipy.stats closed forms, central finite differences in fp64."""
import math

import numpy as np
import pytest
import torch
from scipy import special, stats

from oracle import mile_oracle as o

SMALL = ['airfoil_3x16', 'airfoil_2x16', 'bikesharing_2x16', 'protein_2x16', 'covertype_ref']


def _torch_logpost(spec, theta, X, y, dtype):
    th = torch.tensor(theta, dtype=dtype, requires_grad=True)
    Xt = torch.tensor(X, dtype=dtype)
    bias_off, kern_off = spec.offsets()
    d = spec.dims
    a = Xt
    for l in range(spec.n_layers):
        b = th[bias_off[l]:bias_off[l] + d[l + 1]]
        w = th[kern_off[l]:kern_off[l] + d[l] * d[l + 1]].reshape(d[l], d[l + 1])
        a = a @ w + b
        if l < spec.n_layers - 1:
            a = {'relu': torch.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh,
                 'gelu': lambda v: torch.nn.functional.gelu(v, approximate='tanh'),
                 'leaky_relu': lambda v: torch.nn.functional.leaky_relu(v, 0.01),
                 'identity': lambda v: v}[spec.activation](a)
    if spec.task == 'regr':
        yt = torch.tensor(y, dtype=dtype)
        sigma = torch.exp(a[:, 1]).clamp(1e-6, 1e6)
        ll = torch.distributions.Normal(a[:, 0], sigma).log_prob(yt).sum()
    else:
        yt = torch.tensor(y, dtype=torch.long)
        ll = -torch.nn.functional.cross_entropy(a, yt, reduction='sum')
    if spec.prior == 'normal':
        lp = torch.distributions.Normal(torch.tensor(spec.prior_loc, dtype=dtype),
                                        torch.tensor(spec.prior_scale, dtype=dtype)).log_prob(th).sum()
    else:
        lp = torch.distributions.Laplace(torch.tensor(spec.prior_loc, dtype=dtype),
                                         torch.tensor(spec.prior_scale, dtype=dtype)).log_prob(th).sum()
    val = lp + ll * spec.n_batches
    val.backward()
    return val.item(), th.grad.numpy()


@pytest.mark.parametrize('name', SMALL)
def test_value_and_grad_vs_torch_fp64(name):
    spec = o.make_spec(name)
    X, y, _, _ = o.synthetic_data(name, n_train=257)
    theta = o.synthetic_theta0(spec, 1)[0].astype(np.float64)
    lp, g = o.logpost_value_and_grad(spec, theta, X.astype(np.float64), y)
    lp_t, g_t = _torch_logpost(spec, theta, X, y, torch.float64)
    assert abs(lp - lp_t) <= 1e-10 * abs(lp_t)
    assert np.linalg.norm(g - g_t) <= 1e-10 * np.linalg.norm(g_t)


@pytest.mark.parametrize('name', SMALL)
def test_value_and_grad_fp32_twin(name):
    """fp32 oracle vs its fp64 twin: separates 'our error' from fp32 rounding."""
    spec = o.make_spec(name)
    X, y, _, _ = o.synthetic_data(name, n_train=1052)
    theta = o.synthetic_theta0(spec, 1)[0]
    lp32, g32 = o.logpost_value_and_grad(spec, theta, X, y)
    lp64, g64 = o.logpost_value_and_grad(spec, theta.astype(np.float64), X.astype(np.float64), y)
    assert lp32.dtype == np.float32 and g32.dtype == np.float32
    assert abs(lp32 - lp64) <= 2e-6 * abs(lp64)
    assert np.linalg.norm(g32 - g64) <= 5e-6 * np.linalg.norm(g64)


@pytest.mark.parametrize('act', ['relu', 'sigmoid', 'tanh', 'gelu', 'leaky_relu', 'identity'])
@pytest.mark.parametrize('task', ['regr', 'class'])
@pytest.mark.parametrize('prior', ['normal', 'laplace'])
def test_activations_tasks_priors_vs_torch(act, task, prior):
    K = 2 if task == 'regr' else 5
    spec = o.ModelSpec(7, (9, 6, K), act, task, prior, prior_loc=0.1, prior_scale=1.7)
    rng = np.random.default_rng(3)
    X = rng.standard_normal((40, 7))
    y = rng.standard_normal(40) if task == 'regr' else rng.integers(0, K, 40).astype(np.int32)
    theta = rng.standard_normal(spec.n_params) * 0.4
    lp, g = o.logpost_value_and_grad(spec, theta, X, y)
    lp_t, g_t = _torch_logpost(spec, theta, X, y, torch.float64)
    assert abs(lp - lp_t) <= 1e-10 * abs(lp_t)
    assert np.linalg.norm(g - g_t) <= 1e-9 * np.linalg.norm(g_t)


def test_grad_vs_finite_differences_fp64():
    spec = o.ModelSpec(5, (8, 8, 2), 'tanh', 'regr')
    rng = np.random.default_rng(0)
    X = rng.standard_normal((64, 5))
    y = rng.standard_normal(64)
    theta = rng.standard_normal(spec.n_params) * 0.3
    _, g = o.logpost_value_and_grad(spec, theta, X, y)
    h = 1e-6
    for i in rng.choice(spec.n_params, 25, replace=False):
        e = np.zeros_like(theta)
        e[i] = h
        fd = (o.logpost_value_and_grad(spec, theta + e, X, y)[0]
              - o.logpost_value_and_grad(spec, theta - e, X, y)[0]) / (2 * h)
        assert abs(fd - g[i]) <= 1e-6 * max(1.0, abs(g[i]))


def test_logdensity_vs_scipy():
    """probabilistic.py:93-109 and priors.py:101-108 against scipy closed forms."""
    spec = o.ModelSpec(4, (6, 2), 'relu', 'regr')
    rng = np.random.default_rng(1)
    X = rng.standard_normal((30, 4))
    y = rng.standard_normal(30)
    theta = rng.standard_normal(spec.n_params) * 0.5
    out = o.forward(spec, theta, X)
    ll = stats.norm.logpdf(y, loc=out[:, 0], scale=np.clip(np.exp(out[:, 1]), 1e-6, 1e6)).sum()
    lp = stats.norm.logpdf(theta).sum()
    val, _ = o.logpost_value_and_grad(spec, theta, X, y)
    assert abs(val - (ll + lp)) < 1e-10 * abs(ll + lp)
    specc = o.ModelSpec(4, (6, 3), 'sigmoid', 'class')
    yc = rng.integers(0, 3, 30).astype(np.int32)
    thetac = rng.standard_normal(specc.n_params) * 0.5
    outc = o.forward(specc, thetac, X)
    llc = special.log_softmax(outc, axis=1)[np.arange(30), yc].sum()
    valc, _ = o.logpost_value_and_grad(specc, thetac, X, yc)
    assert abs(valc - (llc + stats.norm.logpdf(thetac).sum())) < 1e-10 * abs(valc)


def test_clip_and_nansum_semantics():
    """clip passes no gradient outside [1e-6,1e6] (out[...,1] outside +-13.8155);
    NaN rows contribute 0 (jnp.nansum)."""
    spec = o.ModelSpec(1, (2,), 'relu', 'regr')
    # theta = [b0, b1, w00, w01]; out = [b0 + x w00, b1 + x w01]
    theta = np.array([0.0, 20.0, 0.0, 0.0])
    X = np.ones((3, 1))
    y = np.array([0.5, -0.5, np.nan])
    val, g = o.logpost_value_and_grad(spec, theta, X, y)
    sigma = 1e6
    expect = 2 * (-0.5 * math.log(2 * math.pi * sigma ** 2)) - 0.5 * (0.25 + 0.25) / sigma ** 2
    prior, pg = o.log_prior(spec, theta)
    assert abs(val - (expect + prior)) < 1e-9
    # d/db1 of likelihood is exactly 0 (clipped), so g[1] is the prior gradient only
    assert g[1] == pg[1]


def test_flat_layout_is_ravel_pytree_order():
    """per layer bias then kernel, layers in lexicographic order."""
    spec = o.ModelSpec(3, (4, 2), 'relu', 'regr')
    theta = np.arange(spec.n_params, dtype=np.float64)
    tree = o.unravel_tree(spec, theta)
    assert tree['fcn']['layer0']['bias'].tolist() == [0, 1, 2, 3]
    assert tree['fcn']['layer0']['kernel'].shape == (3, 4)
    assert tree['fcn']['layer0']['kernel'][0].tolist() == [4, 5, 6, 7]
    assert tree['fcn']['layer1']['bias'].tolist() == [16, 17]
    np.testing.assert_array_equal(o.ravel_tree(spec, tree), theta)
    deep = o.ModelSpec(2, (2,) * 11 + (2,), 'relu', 'regr')
    assert deep.layer_order[:4] == (0, 1, 10, 11)  # 'layer10' < 'layer2'
