"""Diagonal preconditioning (src/training/warmup.py:385-401, blackjax `sqrt_diag_cov`): the preconditioned MCLMC step and
the tuning loop on the CUDA path against the numpy oracle, and the warmup mirror with `diagonal_preconditioning=True`
(the reference's default)."""
import numpy as np
import pytest

from oracle import mile_oracle as o

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(b), 1e-30)


def make(name, C, n_train=None, **opts):
    from mile_b200 import Ensemble, FCNSpec
    ospec = o.make_spec(name)
    X, y, _, _ = o.synthetic_data(name, n_train=n_train)
    ens = Ensemble(FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task), C, **opts)
    ens.set_data(X, y)
    return ospec, ens, X, y


@pytest.mark.parametrize('name,opts', [('airfoil_3x16', {}), ('airfoil_3x16', {'cluster_size': 1}), ('airfoil_3x16', {'fast': 1}),
                                       ('bikesharing_2x16', {'cluster_size': 14}), ('covertype_ref', {}), ('covertype_ref', {'cluster_size': 2})])
@pytest.mark.parametrize('refresh_mode', [0, 1])
def test_preconditioned_step_matches_oracle(name, opts, refresh_mode):
    """Two MCLMC steps with a non-trivial sqrt_diag_cov: (theta', u', l', dK, dE) within 1e-5 of the fp64 oracle
    (mclmc_step(sqrt_diag_cov=m): B-steps see m .* g, A-steps move by eps * m .* u)."""
    C, n = 3, 2
    ospec, ens, X, y = make(name, C, n_train=1500, **opts)
    ens.set_option('refresh_mode', refresh_mode)
    d = ospec.n_params
    rng = np.random.default_rng(5)
    th0 = o.synthetic_theta0(ospec, C)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    nslot = 2 if refresh_mode else 1
    z = rng.standard_normal((n, nslot, C, d)).astype(np.float32) if refresh_mode else rng.standard_normal((n, C, d)).astype(np.float32)
    sdc = np.exp(rng.uniform(np.log(0.3), np.log(3.0), size=(C, d))).astype(np.float32)
    eps, L = 0.01, float(np.sqrt(d))
    ens.init(th0, z0)
    ens.set_sqrt_diag_cov(sdc)
    assert np.array_equal(ens.get_sqrt_diag_cov(), sdc)
    _, info = ens.sample(n, eps, L, z=z, keep=False, info=True)
    th, u, lp, g = ens.get_state()
    f64 = lambda t: o.logpost_value_and_grad(ospec, t, X.astype(np.float64), y)
    for c in range(C):
        st = o.mclmc_init(f64, th0[c].astype(np.float64), z0[c].astype(np.float64))
        for s in range(n):
            zz = z[s, :, c].astype(np.float64) if refresh_mode else z[s, c].astype(np.float64)
            st, inf = o.mclmc_step(f64, st, eps, L, zz, sqrt_diag_cov=sdc[c].astype(np.float64),
                                   refresh='maruyama' if refresh_mode else 'post')
        scale = abs(st.logdensity)
        assert rel(th[c], st.position) <= 1e-5
        assert rel(u[c], st.momentum) <= 2e-5
        assert abs(lp[c] - st.logdensity) <= 1e-5 * scale
        assert rel(g[c], st.logdensity_grad) <= 1e-4
        assert abs(info[n - 1, c, 1] - inf.kinetic_change) <= 1e-5 * scale
        assert abs(info[n - 1, c, 2] - inf.energy_change) <= 2e-5 * scale
    # clearing it restores the plain dynamics
    ens.set_sqrt_diag_cov(None)
    assert np.all(ens.get_sqrt_diag_cov() == 1.0)
    ens.init(th0, z0)
    ens.sample(1, eps, L, z=z[:1], keep=False)
    th1 = ens.get_state()[0]
    for c in range(C):
        st = o.mclmc_init(f64, th0[c].astype(np.float64), z0[c].astype(np.float64))
        zz = z[0, :, c].astype(np.float64) if refresh_mode else z[0, c].astype(np.float64)
        st, _ = o.mclmc_step(f64, st, eps, L, zz, refresh='maruyama' if refresh_mode else 'post')
        assert rel(th1[c], st.position) <= 1e-5
    ens.close()


def test_preconditioned_step_wide_path():
    """The integrator kernel of the wide / sharded path applies the same preconditioned dynamics."""
    from mile_b200 import Ensemble, FCNSpec
    F, N, C, widths = 12, 600, 2, (128, 128, 2)
    ospec = o.ModelSpec(F, widths, 'relu', 'regr')
    rng = np.random.default_rng(2)
    X = rng.standard_normal((N, F)).astype(np.float32)
    y = rng.standard_normal(N).astype(np.float32)
    d = ospec.n_params
    th0 = (rng.standard_normal((C, d)) * (0.5 / np.sqrt(128))).astype(np.float32)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    z = rng.standard_normal((1, C, d)).astype(np.float32)
    sdc = np.exp(rng.uniform(np.log(0.5), np.log(2.0), size=(C, d))).astype(np.float32)
    ens = Ensemble(FCNSpec(F, widths, 'relu', 'regr'), C)
    ens.set_data(X, y)
    assert ens.get_option('wide') == 1
    ens.init(th0, z0)
    ens.set_sqrt_diag_cov(sdc)
    _, info = ens.sample(1, 0.01, float(np.sqrt(d)), z=z, keep=False, info=True)
    th, u, lp, _ = ens.get_state()
    f64 = lambda t: o.logpost_value_and_grad(ospec, t, X.astype(np.float64), y)
    for c in range(C):
        st = o.mclmc_init(f64, th0[c].astype(np.float64), z0[c].astype(np.float64))
        st, inf = o.mclmc_step(f64, st, 0.01, float(np.sqrt(d)), z[0, c].astype(np.float64), sqrt_diag_cov=sdc[c].astype(np.float64))
        assert rel(th[c], st.position) <= 1e-5
        assert rel(u[c], st.momentum) <= 2e-5
        assert abs(lp[c] - st.logdensity) <= 1e-5 * abs(st.logdensity)
        assert abs(info[0, c, 2] - inf.energy_change) <= 2e-5 * abs(st.logdensity)
    ens.close()


def test_precondition_from_moments_and_readjust_match_oracle():
    """warmup.py:385-401 on the device: after phases 1+2, sqrt_diag_cov = sqrt(E[x^2]-E[x]^2) and L = sqrt(d); the
    following tune2//3 re-adjustment steps run the preconditioned kernel with a restarted adaptive state.  The oracle's
    predictor, driven by the GPU's own energy changes, must reproduce the step-size trajectory (same criterion as
    test_tuning_matches_oracle), and the preconditioner must equal the one computed from the GPU's moments."""
    name, C = 'airfoil_2x16', 2
    ospec, ens, X, y = make(name, C, n_train=500)
    d = ospec.n_params
    rng = np.random.default_rng(9)
    th0 = o.synthetic_theta0(ospec, C)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    t1, t2 = 12, 9
    zt = rng.standard_normal((t1 + t2, C, d)).astype(np.float32)
    zr = rng.standard_normal((t2 // 3, C, d)).astype(np.float32)
    ens.init(th0, z0)
    ens.tune_reset(0.01)
    tc = ens.tune_cfg(t1, t2, 0.5, 0.1, 1.5, 100)
    ens.tune(t1 + t2, 0, tc, z=zt)
    ens.tune_finish_phase2()
    eps0, L0, _, mx, mx2 = ens.get_tuning(moments=True)
    ens.precondition_from_moments()
    sdc = ens.get_sqrt_diag_cov()
    var = mx.astype(np.float64) ** 2
    assert rel(sdc, np.sqrt(np.maximum(mx2.astype(np.float64) - var, 0))) <= 2e-3     # (fp32 cancellation in E[x^2] - E[x]^2)
    _, Ld, _ = ens.get_tuning()
    assert np.allclose(Ld, np.sqrt(d), rtol=1e-6)
    start = ens.get_state()
    ens.tune_reset(0.0)
    ens.set_tuning(step_size=eps0, L=Ld)
    ti = ens.tune(t2 // 3, 0, tc, z=zr, info=True)
    cfg = o.TuneConfig(t1, t2, 0, 0.5, 0.1, 1.5, 100, 0.01)
    f64 = lambda t: o.logpost_value_and_grad(ospec, t, X.astype(np.float64), y)
    th_end = ens.get_state()[0]
    for c in range(C):
        # first re-adjustment step against the full oracle step (identical inputs)
        st = o.IntegratorState(start[0][c].astype(np.float64), start[1][c].astype(np.float64), np.float64(start[2][c]),
                               start[3][c].astype(np.float64))
        _, inf = o.mclmc_step(f64, st, float(eps0[c]), float(Ld[c]), zr[0, c].astype(np.float64), sqrt_diag_cov=sdc[c].astype(np.float64))
        assert abs(ti[0, c, 0] - inf.energy_change) <= 2e-5 * abs(start[2][c])
        # predictor arithmetic with the restarted adaptive state
        ts = o.tune_init(cfg, d, np.float64)._replace(step_size=np.float64(eps0[c]))
        for i in range(t2 // 3):
            assert ti[i, c, 3] == 1.0
            ts = o.tune_update(cfg, ts._replace(step_size_max=np.float64(ti[i, c, 2])), np.zeros(d), np.float64(ti[i, c, 0]), True, i)
            assert abs(ti[i, c, 1] - ts.step_size) <= 2e-5 * ts.step_size
    assert np.all(np.isfinite(th_end))
    ens.close()


def test_custom_mclmc_warmup_default_diagonal_preconditioning():
    """The reference's default call `custom_mclmc_warmup(logdensity_fn)` (diagonal_preconditioning=True) runs on the CUDA
    path: finite positive step size / L, a non-trivial preconditioner in MCLMCAdaptationState.sqrt_diag_cov, and
    warmup_mclmc accepts `diagonal_preconditioning: true` from the sampler config."""
    import functools
    from mile_b200 import FCN, PriorDist, ProbabilisticModel, custom_mclmc_warmup
    from mile_b200.config import SamplerConfig
    from mile_b200.sampling import warmup_mclmc
    name = 'airfoil_2x16'
    ospec = o.make_spec(name)
    X, y, _, _ = o.synthetic_data(name, n_train=400)
    module = FCN(ospec.widths, ospec.activation)
    rng = np.random.default_rng(2)
    pm = ProbabilisticModel(module, module.init(rng, ospec.n_features), PriorDist.StandardNormal.get_prior(), 'regr')
    log_post = functools.partial(pm.log_unnormalized_posterior, x=X, y=y)
    C = 3
    pos = [module.init(rng, ospec.n_features, scale=0.5) for _ in range(C)]
    tree = {'fcn': {k: {kk: np.stack([p['fcn'][k][kk] for p in pos]) for kk in v} for k, v in pos[0]['fcn'].items()}}
    res = custom_mclmc_warmup(log_post, desired_energy_var_start=0.5, desired_energy_var_end=0.1, step_size_init=0.01).run(0, tree, 600)
    eps, L, sdc = res.parameters.step_size, res.parameters.L, res.parameters.sqrt_diag_cov
    assert eps.shape == (C,) and np.all(np.isfinite(eps)) and np.all(eps > 0)
    assert np.all(np.isfinite(L)) and np.all(L > 0)
    assert sdc.shape == (C, ospec.n_params) and np.all(np.isfinite(sdc)) and np.all(sdc >= 0) and sdc.std() > 0
    cfg = SamplerConfig(name='mclmc', warmup_steps=300, n_samples=20, n_thinning=2, diagonal_preconditioning=True,
                        desired_energy_var_start=0.5, desired_energy_var_end=0.1, step_size_init=0.01)
    _, p = warmup_mclmc(cfg, 1, tree, log_post, C)
    assert np.all(np.isfinite(p['step_size'])) and np.all(np.asarray(p['L']) > 0)
