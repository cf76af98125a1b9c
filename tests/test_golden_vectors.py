"""Regression pin: tests/golden/oracle_vectors.npz (frozen fp64 oracle trajectories, made by
tests/golden/make_oracle_vectors.py).  CPU: the numpy oracle and the C oracle still reproduce them.  GPU: the CUDA path
matches them at the parity bar (1e-5 relative in fp32; energies relative to |log-density|)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import mile_oracle as o

G = np.load(Path(__file__).resolve().parent / 'golden' / 'oracle_vectors.npz')
NAMES = ('airfoil_3x16', 'bikesharing_2x16', 'covertype_ref')


def rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize('name', NAMES)
def test_numpy_oracle_reproduces_golden(name):
    spec = o.make_spec(name)
    X, y = G[f'{name}.X'].astype(np.float64), G[f'{name}.y']
    f = lambda t: o.logpost_value_and_grad(spec, t, X, y)
    for c in range(2):
        st = o.mclmc_init(f, G[f'{name}.theta0'][c].astype(np.float64), G[f'{name}.z0'][c].astype(np.float64))
        assert rel(st.logdensity_grad, G[f'{name}.grad0'][c]) <= 1e-12
        for s in range(4):
            st, info = o.mclmc_step(f, st, 0.02, 15.0, G[f'{name}.z'][s, c].astype(np.float64))
            assert abs(info.energy_change - G[f'{name}.energy_change'][c, s]) <= 1e-9 * abs(st.logdensity)
        assert rel(st.position, G[f'{name}.position'][c]) <= 1e-12 and rel(st.momentum, G[f'{name}.momentum'][c]) <= 1e-10
        assert abs(st.logdensity - G[f'{name}.logdensity'][c]) <= 1e-11 * abs(st.logdensity)


def test_c_oracle_reproduces_golden_gradient():
    from oracle import c_oracle
    if not c_oracle.available():
        pytest.skip('C oracle not built')
    for name in ('airfoil_3x16', 'bikesharing_2x16'):
        spec = o.make_spec(name)
        lp, g = c_oracle.logpost_batch(spec, G[f'{name}.theta0'], G[f'{name}.X'], G[f'{name}.y'])
        for c in range(2):
            assert rel(g[c], G[f'{name}.grad0'][c]) <= 2e-5      # fp32 C restatement against the fp64 vector


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_cuda_path_reproduces_golden(name):
    from mile_b200 import Ensemble, FCNSpec
    spec = o.make_spec(name)
    ens = Ensemble(FCNSpec(spec.n_features, spec.widths, spec.activation, spec.task), 2)
    ens.set_data(G[f'{name}.X'], G[f'{name}.y'])
    ens.set_test(G[f'{name}.Xt'], G[f'{name}.yt'])
    lp0, g0 = ens.value_and_grad(G[f'{name}.theta0'])
    for c in range(2):
        assert rel(g0[c], G[f'{name}.grad0'][c]) <= 1e-5
    ens.init(G[f'{name}.theta0'], G[f'{name}.z0'])
    _, info = ens.sample(4, 0.02, 15.0, z=G[f'{name}.z'], info=True)
    th, u, lp, _ = ens.get_state()
    for c in range(2):
        assert rel(th[c], G[f'{name}.position'][c]) <= 1e-5
        assert rel(u[c], G[f'{name}.momentum'][c]) <= 5e-5
        assert abs(lp[c] - G[f'{name}.logdensity'][c]) <= 1e-5 * abs(G[f'{name}.logdensity'][c])
        assert np.max(np.abs(info[:, c, 2] - G[f'{name}.energy_change'][c])) <= 2e-5 * abs(G[f'{name}.logdensity'][c])
    ens.lppd_reset()
    ens.lppd_accumulate(th)
    m, s, cnt = ens.lppd_state()
    from mile_b200 import lppd_from_state
    assert abs(lppd_from_state(m, s, cnt * 2) - float(G[f'{name}.lppd'])) <= 1e-4 * max(1.0, abs(float(G[f'{name}.lppd'])))
    ens.close()
