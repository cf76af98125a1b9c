"""Chain diagnostics mirror (src/inference/metrics.py:226-244,354-425,449-523) against the
oracle's numpy/scipy restatement.  Device-agnostic torch code: checked here on CPU in fp64, and on the GPU when present."""
import numpy as np
import pytest
import torch

from oracle import mile_oracle as o


def ar1(rng, C, S, D, phi):
    x = np.zeros((C, S, D))
    e = rng.standard_normal((C, S, D))
    for s in range(1, S):
        x[:, s] = phi * x[:, s - 1] + e[:, s]
    return x + rng.standard_normal((C, 1, D)) * 0.3


def check(dev):
    from mile_b200 import diagnostics as dg
    rng = np.random.default_rng(3)
    x = ar1(rng, 4, 400, 5, 0.6)
    t = torch.from_numpy(x).to(dev)
    assert np.allclose(dg.rank_normalize_array(t.reshape(-1, 5)).cpu().numpy(),
                       np.apply_along_axis(o.rank_normalize_array, 0, x.reshape(-1, 5)), atol=1e-10)
    assert np.allclose(dg.between_chain_var(t).cpu().numpy(), o.between_chain_var(x))
    assert np.allclose(dg.within_chain_var(t).cpu().numpy(), o.within_chain_var(x))
    for rn in (True, False):
        assert np.allclose(dg.gelman_split_r_hat(t, 4, rn).cpu().numpy(), o.gelman_split_r_hat(x, 4, rn), rtol=1e-9)
        assert np.allclose(dg.split_chain_r_hat(t, 4, rn).cpu().numpy(), o.split_chain_r_hat(x, 4, rn), rtol=1e-9)
        assert np.allclose(dg.chain_effective_sample_size(t, rn).cpu().numpy(), o.ess_rank_normalized(x, rn), rtol=1e-8)
    rm = dg.running_mean(t, 1).cpu().numpy()
    assert np.allclose(rm[:, -1], x.mean(axis=1)) and np.allclose(rm[:, 0], x[:, 0])
    # ties are ranked by their average, like scipy.stats.rankdata
    xi = np.round(x[:, :50, :2] * 2) / 2
    assert np.allclose(dg.rank_normalize_array(torch.from_numpy(xi).to(dev).reshape(-1, 2)).cpu().numpy(),
                       np.apply_along_axis(o.rank_normalize_array, 0, xi.reshape(-1, 2)), atol=1e-10)
    with pytest.raises(ValueError):
        dg.gelman_split_r_hat(t, 3)


def test_diagnostics_cpu():
    check('cpu')


@pytest.mark.gpu
def test_diagnostics_gpu():
    check('cuda:0')


def test_metrics_mirror_names():
    from mile_b200 import metrics
    rng = np.random.default_rng(5)
    x = ar1(rng, 3, 200, 4, 0.4)
    assert np.allclose(metrics.split_chain_r_hat(x, 4), o.split_chain_r_hat(x, 4), rtol=1e-8)
    assert np.allclose(metrics.effective_sample_size(x), o.ess_rank_normalized(x), rtol=1e-7)
    assert np.allclose(metrics.rank_normalize_array(x), o.rank_normalize_array(x), atol=1e-9)
    assert np.allclose(metrics.between_chain_var(x), o.between_chain_var(x))
