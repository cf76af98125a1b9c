"""The chain-batched torch-CPU restatement (oracle/torch_batched.py, the second CPU-baseline mode of bench.py) against
the numpy oracle: value_and_grad for the named shapes, and one MCLMC step with the same noise."""
import numpy as np
import pytest
import torch

from oracle import mile_oracle as o
from oracle import torch_batched as tb


def rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize('name', ['airfoil_3x16', 'bikesharing_2x16', 'covertype_ref'])
def test_value_and_grad_matches_numpy_oracle(name):
    spec = o.make_spec(name)
    X, y, _, _ = o.synthetic_data(name, n_train=700)
    th = o.synthetic_theta0(spec, 3)
    lp64, g64 = o.logpost_batch(spec, th.astype(np.float64), X.astype(np.float64), y)
    yt = torch.from_numpy(y).to(torch.float32 if spec.task.startswith('regr') else torch.int64)
    lp, g = tb.value_and_grad(spec, torch.from_numpy(th), torch.from_numpy(X), yt)
    for c in range(3):
        assert abs(lp[c].item() - lp64[c]) <= 2e-5 * abs(lp64[c])
        assert rel(g[c].numpy(), g64[c]) <= 5e-5


def test_step_matches_numpy_oracle():
    name = 'airfoil_2x16'
    spec = o.make_spec(name)
    X, y, _, _ = o.synthetic_data(name, n_train=400)
    C, d = 2, spec.n_params
    th = o.synthetic_theta0(spec, C)
    rng = np.random.default_rng(0)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    z = rng.standard_normal((C, d)).astype(np.float32)
    Xt, yt = torch.from_numpy(X), torch.from_numpy(y)
    f = lambda t: tb.value_and_grad(spec, t, Xt, yt)
    theta = torch.from_numpy(th)
    lp, g = f(theta)
    u = torch.from_numpy(z0) / torch.from_numpy(z0).norm(dim=1, keepdim=True)
    eps, L = torch.full((C, 1), 0.01), torch.full((C, 1), 20.0)
    theta, u, lp, g, info = tb.mclmc_step(f, theta, u, lp, g, eps, L, torch.from_numpy(z))
    f64 = lambda t: o.logpost_value_and_grad(spec, t, X.astype(np.float64), y)
    for c in range(C):
        st = o.mclmc_init(f64, th[c].astype(np.float64), z0[c].astype(np.float64))
        st, inf = o.mclmc_step(f64, st, 0.01, 20.0, z[c].astype(np.float64))
        assert rel(theta[c].numpy(), st.position) <= 1e-5
        assert rel(u[c].numpy(), st.momentum) <= 1e-3      # literal fp32 blackjax formula: cancellation in 1 - exp(-delta)
        assert abs(lp[c].item() - st.logdensity) <= 1e-5 * abs(st.logdensity)
