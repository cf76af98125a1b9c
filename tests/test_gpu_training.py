"""Deep-ensemble warm-start training on the device (csrc/mile_train.cuh through the C ABI) against the training oracle
(oracle/train_oracle.py, itself pinned against torch autograd + torch.optim in tests/test_oracle_training.py) with the
SAME minibatch order: per-step losses, metrics, parameters and optimizer moments."""
import numpy as np
import pytest

from oracle import mile_oracle as o
from oracle import train_oracle as t

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize('name,kind,B', [('airfoil_3x16', 'adamw', 32), ('airfoil_2x16', 'adam', 50), ('airfoil_2x16', 'sgd', 32),
                                         ('covertype_ref', 'adamw', 64)])
def test_train_epochs_match_oracle(name, kind, B):
    from mile_b200 import Ensemble, FCNSpec
    from mile_b200.training import init_params, make_batches
    ospec = o.make_spec(name)
    X, y, Xt, yt = o.synthetic_data(name, n_train=640, n_test=200)
    C = 3
    spec = FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task)
    rng = np.random.default_rng(5)
    th0 = init_params(spec, rng, C)
    ens = Ensemble(spec, C)
    ens.set_data(X, y)
    ens.set_test(Xt, yt)
    opt = dict(learning_rate=2e-3, b1=0.9, b2=0.999, eps=1e-8)
    cfg = ens.opt_cfg(kind, **opt)
    ens.train_init(th0)
    batches = [make_batches(640, B, rng) for _ in range(2)]
    mets = np.concatenate([ens.train_epoch(b, cfg) for b in batches])
    th, m, v, tc = ens.train_state()
    ev = ens.eval_metrics(None, 'test')
    nsteps = sum(len(b) for b in batches)
    assert mets.shape == (nsteps, C, 2) and np.all(tc == nsteps)        # TrainState.step counts every applied update
    for c in range(C):
        st = t.OptState(spec.n_params)
        tho = th0[c].astype(np.float64)
        ms = []
        for b in batches:
            tho, mm = t.train_epoch(ospec, tho, st, X.astype(np.float64), y, b, kind=kind, lr=2e-3, b1=0.9, b2=0.999, eps=1e-8,
                                    wd=1e-4)
            ms.append(mm)
        ms = np.concatenate(ms)
        # the loss of every step is computed BEFORE its update, so it tracks the whole parameter trajectory
        np.testing.assert_allclose(mets[:, c, 0], ms[:, 0], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(mets[:, c, 1], ms[:, 1], rtol=2e-4, atol=2e-3)
        assert rel(th[c], tho) <= 1e-4, rel(th[c], tho)
        if kind != 'sgd':
            assert rel(m[c], st.m) <= 1e-3 and rel(v[c], st.v) <= 1e-3
        lo, au = t.eval_metrics(ospec, tho, Xt.astype(np.float64), yt)
        assert abs(ev[c, 0] - lo) <= 1e-4 * max(1.0, abs(lo)) and abs(ev[c, 1] - au) <= 2e-3
    ens.close()


def test_stopped_members_are_untouched_and_report_nan():
    from mile_b200 import Ensemble, FCNSpec
    from mile_b200.training import init_params, make_batches
    ospec = o.make_spec('airfoil_2x16')
    X, y, _, _ = o.synthetic_data('airfoil_2x16', n_train=320)
    spec = FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task)
    rng = np.random.default_rng(1)
    th0 = init_params(spec, rng, 3)
    ens = Ensemble(spec, 3)
    ens.set_data(X, y)
    ens.train_init(th0)
    cfg = ens.opt_cfg('adamw', learning_rate=1e-3)
    mets = ens.train_epoch(make_batches(320, 32, rng), cfg, stopped=[False, True, False])
    th, _, _, tc = ens.train_state()
    np.testing.assert_array_equal(th[1], th0[1])
    assert tc.tolist() == [10, 0, 10]
    assert np.all(np.isnan(mets[:, 1])) and np.all(np.isfinite(mets[:, 0])) and np.all(np.isfinite(mets[:, 2]))
    assert not np.array_equal(th[0], th0[0])
    ens.close()


def test_train_warmstart_writes_reference_layout_and_learns(tmp_path):
    """BDETrainer.train_warmstart (trainer.py:330-364): members are trained with early stopping and saved as
    warmstart/params_{i}.npz, which load_params_batch stacks for start_sampling (trainer.py:556-575)."""
    from mile_b200 import FCNSpec
    from mile_b200.training import train_warmstart
    from mile_b200.utils import load_params_batch
    ospec = o.make_spec('airfoil_2x16')
    X, y, Xt, yt = o.synthetic_data('airfoil_2x16', n_train=800, n_test=300)
    spec = FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task)
    params, metrics = train_warmstart(spec, (X[:650], y[:650]), (X[650:], y[650:]), (Xt, yt), tmp_path, [0, 1, 2, 3],
                                      optimizer=dict(name='adamw', learning_rate=5e-3), max_epochs=60, batch_size=32,
                                      patience=5, seed=3)
    assert params.shape == (4, spec.n_params) and np.all(np.isfinite(params))
    first, last = metrics['valid'][0, :, 0], np.nanmin(metrics['valid'][:, :, 0], axis=0)
    assert np.all(last < first - 0.2), (first, last)          # the Gaussian NLL on held-out rows drops clearly
    assert np.all(metrics['test'][:, 0] < 1.3) and metrics['epochs'] <= 60
    tree = load_params_batch([tmp_path / 'warmstart' / f'params_{i}.npz' for i in range(4)])
    np.testing.assert_array_equal(spec.ravel(tree), params)
    assert (tmp_path / 'tree.json').exists()
