"""Partition sampling (src/training/partition_sampling.py, trainer.py:613-659): first + last layer sampled, hidden layers
frozen.  The CUDA path (frozen-parameter mask) against the numpy oracle run on the REDUCED parameter vector, and the
reference-facing `partition_inference_loop`."""
import functools

import numpy as np
import pytest

from oracle import mile_oracle as o

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(b), 1e-30)


def reduced_logpost(ospec, X, y, base, active):
    """log prior of the sampled parameters + log-likelihood of the merged network, as a function of the sampled
    parameters only (StandardNormal prior)."""
    frozen = np.ones(base.shape[0], bool)
    frozen[active] = False

    def f(ta):
        th = base.astype(ta.dtype).copy()
        th[active] = ta
        lp, g = o.logpost_value_and_grad(ospec, th, X.astype(ta.dtype), y)
        lp_frozen_prior = np.sum(-0.5 * np.log(2 * np.pi) - 0.5 * th[frozen] ** 2)
        return lp - lp_frozen_prior, g[active]
    return f


@pytest.mark.parametrize('name,opts', [('airfoil_3x16', {}), ('airfoil_3x16', {'fast': 0, 'cluster_size': 2}), ('covertype_ref', {'_deep': True})])
def test_partition_step_and_tuning_match_reduced_oracle(name, opts):
    from mile_b200 import Ensemble, FCNSpec
    opts = dict(opts)
    deep = opts.pop('_deep', False)
    ospec = o.make_spec(name)
    if deep:      # covertype_ref has no hidden layer between first and last: use a 54-24-16-7 sigmoid classifier
        ospec = o.ModelSpec(ospec.n_features, (24, 16, 7), 'sigmoid', 'class')
    X, y, _, _ = o.synthetic_data(name, n_train=700)
    C, n = 2, 2
    spec = FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task)
    frozen = spec.hidden_layer_mask()
    assert frozen.any() and not frozen.all()
    active = np.flatnonzero(~frozen)
    d, da = ospec.n_params, active.size
    ens = Ensemble(spec, C, **opts)
    ens.set_data(X, y)
    ens.set_frozen_mask(frozen)
    assert ens.get_option('d_eff') == da
    rng = np.random.default_rng(21)
    th0 = o.synthetic_theta0(ospec, C)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    z = rng.standard_normal((n, C, d)).astype(np.float32)
    eps, L = 0.01, float(np.sqrt(da))
    ens.init(th0, z0)
    th_i, u_i, lp_i, g_i = ens.get_state()
    _, info = ens.sample(n, eps, L, z=z, keep=False, info=True)
    th, u, lp, g = ens.get_state()
    for c in range(C):
        f64 = reduced_logpost(ospec, X, y, th0[c].astype(np.float64), active)
        st = o.mclmc_init(f64, th0[c, active].astype(np.float64), z0[c, active].astype(np.float64))
        assert rel(u_i[c, active], st.momentum) <= 1e-6 and np.all(u_i[c, frozen] == 0)
        assert abs(lp_i[c] - st.logdensity) <= 1e-5 * abs(st.logdensity)
        assert rel(g_i[c, active], st.logdensity_grad) <= 1e-5 and np.all(g_i[c, frozen] == 0)
        for s in range(n):
            st, inf = o.mclmc_step(f64, st, eps, L, z[s, c, active].astype(np.float64))
        scale = abs(st.logdensity)
        assert rel(th[c, active], st.position) <= 1e-5
        np.testing.assert_array_equal(th[c, frozen], th0[c, frozen])          # frozen parameters: bit for bit
        assert rel(u[c, active], st.momentum) <= 2e-5 and np.all(u[c, frozen] == 0)
        assert abs(lp[c] - st.logdensity) <= 1e-5 * scale
        assert abs(info[n - 1, c, 1] - inf.kinetic_change) <= 1e-5 * scale
        assert abs(info[n - 1, c, 2] - inf.energy_change) <= 2e-5 * scale
    # tuning: L_0 = max(sqrt(d_eff), 15), xi uses d_eff; predictor driven by the GPU's own energy changes
    ens.init(th0, z0)
    ens.tune_reset(0.01)
    _, L0, _ = ens.get_tuning()
    assert np.allclose(L0, max(np.sqrt(da), 15.0), rtol=1e-6)
    tc = ens.tune_cfg(4, 2, 0.5, 0.1, 1.5, 100)
    zt = rng.standard_normal((6, C, d)).astype(np.float32)
    ti = ens.tune(6, 0, tc, z=zt, info=True)
    cfg = o.TuneConfig(4, 2, 0, 0.5, 0.1, 1.5, 100, 0.01)
    for c in range(C):
        ts = o.tune_init(cfg, da, np.float64)
        for i in range(6):
            assert ti[i, c, 3] == 1.0
            ts = o.tune_update(cfg, ts._replace(step_size_max=np.float64(ti[i, c, 2])), np.zeros(da), np.float64(ti[i, c, 0]), True, i)
            assert abs(ti[i, c, 1] - ts.step_size) <= 2e-5 * ts.step_size
    np.testing.assert_array_equal(ens.get_state()[0][:, frozen], th0[:, frozen])
    ens.close()


def test_partition_inference_loop_layout(tmp_path):
    """partition_inference_loop: warmup_params.txt, merged samples in the reference layout; the hidden layers of every saved
    sample equal the warm-start values, the first and last layers move."""
    from mile_b200 import FCN, PriorDist, ProbabilisticModel, partition_inference_loop, partition_params
    from mile_b200.config import SamplerConfig
    name = 'airfoil_3x16'
    ospec = o.make_spec(name)
    X, y, _, _ = o.synthetic_data(name, n_train=300)
    module = FCN(ospec.widths, ospec.activation)
    rng = np.random.default_rng(3)
    pm = ProbabilisticModel(module, module.init(rng, ospec.n_features), PriorDist.StandardNormal.get_prior(), 'regr')
    C = 2
    pos = [module.init(rng, ospec.n_features, scale=0.5) for _ in range(C)]
    tree = {'fcn': {k: {kk: np.stack([p['fcn'][k][kk] for p in pos]) for kk in v} for k, v in pos[0]['fcn'].items()}}
    io, hidden = partition_params(tree)
    assert set(io['fcn']) == {'layer0', 'layer3'} and set(hidden['fcn']) == {'layer1', 'layer2'}
    # the partition log-density: prior over the sampled layers only
    lp_part = pm.log_unnormalized_posterior_partition(io, hidden, X, y)
    lp_full = pm.log_unnormalized_posterior(tree, X, y)
    hid_prior = sum(np.sum(-0.5 * np.log(2 * np.pi) - 0.5 * np.asarray(v, np.float64) ** 2, axis=tuple(range(1, np.asarray(v).ndim)))
                    for lay in hidden['fcn'].values() for v in lay.values())
    assert np.allclose(lp_part, lp_full - hid_prior, rtol=1e-5)
    log_post = functools.partial(pm.log_unnormalized_posterior_partition, x=X, y=y)
    cfg = SamplerConfig(name='mclmc', warmup_steps=200, n_samples=40, n_thinning=10, desired_energy_var_start=0.5,
                        desired_energy_var_end=0.1, step_size_init=0.01, partition_sampling=True)
    out = tmp_path / 'exp' / 'samples'
    partition_inference_loop(log_post, cfg, 5, tree, np.arange(C), out)
    assert (tmp_path / 'exp' / 'warmup_params.txt').exists()
    for c in range(C):
        files = sorted((out / str(c)).glob('sample_*.npz'))
        assert [f.name for f in files] == [f'sample_{n}.npz' for n in (0, 10, 20, 30)]
        for f in files:
            z = np.load(f)
            for lay in ('layer1', 'layer2'):
                for part in ('kernel', 'bias'):
                    np.testing.assert_array_equal(z[f'fcn.{lay}.{part}'], tree['fcn'][lay][part][c])
            assert not np.array_equal(z['fcn.layer0.kernel'], tree['fcn']['layer0']['kernel'][c])
            assert not np.array_equal(z['fcn.layer3.kernel'], tree['fcn']['layer3']['kernel'][c])
