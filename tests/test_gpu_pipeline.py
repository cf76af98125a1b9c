"""GPU tests of the reference-facing seams: KERNELS['mclmc'], custom_mclmc_warmup, inference_loop with the
reference's on-disk layout, and statistical parity of a full run against the CPU restatement."""
import functools
import pickle

import numpy as np
import pytest

from oracle import mile_oracle as o

pytestmark = pytest.mark.gpu


def setup_problem(name='airfoil_2x16', n_train=600, n_test=150, seed=0):
    from mile_b200 import FCN, PriorDist, ProbabilisticModel
    ospec = o.make_spec(name)
    X, y, Xt, yt = o.synthetic_data(name, n_train=n_train, n_test=n_test)
    module = FCN(ospec.widths, ospec.activation)
    params = module.init(np.random.default_rng(seed), ospec.n_features)
    pm = ProbabilisticModel(module, params, PriorDist.StandardNormal.get_prior(), 'regr')
    return ospec, module, pm, X, y, Xt, yt


def test_kernel_registry_init_and_step_api():
    """kernel(logdensity_fn, step_size=, L=) -> .init(position, rng_key), .step(rng_key, state) -> (state, info)
    (src/training/kernels/__init__.py:14-18, sampling.py:133,150)."""
    from mile_b200 import KERNELS
    ospec, module, pm, X, y, _, _ = setup_problem()
    log_post = functools.partial(pm.log_unnormalized_posterior, x=X, y=y)
    pos = module.init(np.random.default_rng(1), ospec.n_features, scale=0.5)
    lp_direct = log_post(pos)
    want, gwant = o.logpost_value_and_grad(ospec, o.ravel_tree(ospec, pos).astype(np.float64), X.astype(np.float64), y)
    assert abs(lp_direct - want) <= 1e-5 * abs(want)
    sampler = KERNELS['mclmc'](log_post, step_size=0.01, L=20.0)
    st = sampler.init(pos, 7)
    assert abs(st.logdensity - want) <= 1e-5 * abs(want)
    u = o.ravel_tree(ospec, st.momentum)
    assert abs(np.linalg.norm(u) - 1) < 1e-5
    g = o.ravel_tree(ospec, st.logdensity_grad)
    assert np.linalg.norm(g - gwant) <= 1e-5 * np.linalg.norm(gwant)
    st2, info = sampler.step(8, st)
    assert st2.position['fcn']['layer0']['kernel'].shape == (ospec.n_features, 16)
    assert abs(info.energy_change - (info.kinetic_change - st2.logdensity + st.logdensity)) <= 1e-4 * abs(want)
    lp2, _ = o.logpost_value_and_grad(ospec, o.ravel_tree(ospec, st2.position).astype(np.float64), X.astype(np.float64), y)
    assert abs(st2.logdensity - lp2) <= 1e-5 * abs(lp2)
    # batched positions (leading chain axis) go through the same objects
    posb = {'fcn': {k: {kk: np.stack([vv, vv * 0.9]) for kk, vv in v.items()} for k, v in pos['fcn'].items()}}
    stb = sampler.init(posb, 7)
    assert stb.logdensity.shape == (2,) and abs(stb.logdensity[0] - want) <= 1e-5 * abs(want)


def test_custom_mclmc_warmup_reaches_desired_energy_variance():
    from mile_b200 import KERNELS, custom_mclmc_warmup
    ospec, module, pm, X, y, _, _ = setup_problem()
    log_post = functools.partial(pm.log_unnormalized_posterior, x=X, y=y)
    rng = np.random.default_rng(2)
    C = 4
    pos = [module.init(rng, ospec.n_features, scale=0.5) for _ in range(C)]
    posb = {'fcn': {k: {kk: np.stack([p['fcn'][k][kk] for p in pos]) for kk in v} for k, v in pos[0]['fcn'].items()}}
    algo = custom_mclmc_warmup(log_post, diagonal_preconditioning=False, desired_energy_var_start=0.5,
                               desired_energy_var_end=0.1, trust_in_estimate=1.5, num_effective_samples=100,
                               step_size_init=0.01)
    state, params = algo.run(11, posb, 3000)
    assert params.step_size.shape == (C,) and params.L.shape == (C,)
    assert np.all(np.isfinite(params.step_size)) and np.all(params.step_size > 0)
    assert np.all(np.isfinite(params.L)) and np.all(params.L > 0)
    assert params.sqrt_diag_cov.shape == (C, ospec.n_params) and np.all(params.sqrt_diag_cov == 1.0)   # (preconditioning: tests/test_gpu_preconditioning.py)
    # Var[dE]/d with the tuned step size is of the order of the target (the predictor's premise)
    from mile_b200 import Ensemble, FCNSpec
    ens = pm.make_ensemble(C, X, y)
    th = pm.spec.ravel(state.position)
    ens.set_state(th, pm.spec.ravel(state.momentum), state.logdensity, pm.spec.ravel(state.logdensity_grad))
    _, info = ens.sample(600, params.step_size, params.L, seed=5, keep=False, info=True)
    v = info[100:, :, 2].var(axis=0) / ospec.n_params
    # (a chain started from a random init can still be in its transient after 3000 warmup steps; the
    #  reference warm-starts from trained ensemble members and uses 50000 -- judge the typical chain)
    assert 0.1 / 6 < np.median(v) < 0.1 * 6, v
    ens.close()


def test_inference_loop_outputs_and_layout(tmp_path):
    """sampling.py:32-216 side effects: warmup_params.txt (2 lines), samples/{id}/sample_{n}.npz with
    n % n_thinning == 0, samples/info.pkl == {} ; fused LPPD equals the post-hoc LPPD of the written samples."""
    from mile_b200 import SamplerConfig, inference_loop
    from mile_b200.utils import load_samples_from_dir
    ospec, module, pm, X, y, Xt, yt = setup_problem()
    pm.attach_test_split(Xt, yt)
    C = 3
    rng = np.random.default_rng(3)
    pos = [module.init(rng, ospec.n_features, scale=0.5) for _ in range(C)]
    posb = {'fcn': {k: {kk: np.stack([p['fcn'][k][kk] for p in pos]) for kk in v} for k, v in pos[0]['fcn'].items()}}
    cfg = SamplerConfig.from_dict({'name': 'mclmc', 'warmup_steps': 400, 'n_chains': C, 'n_samples': 250, 'n_thinning': 10,
                                   'diagonal_preconditioning': False, 'desired_energy_var_start': 0.5,
                                   'desired_energy_var_end': 0.1, 'step_size_init': 0.01})
    log_post = functools.partial(pm.log_unnormalized_posterior, x=X, y=y)
    exp = tmp_path / 'exp'
    info = inference_loop(log_post, cfg, 42, posb, np.array([4, 5, 6]), exp / 'samples')
    lines = (exp / 'warmup_params.txt').read_text().strip().split('\n')
    assert len(lines) == 2 and all(len(l.split(',')) == C for l in lines)
    assert all(float(v) > 0 for l in lines for v in l.split(','))
    with open(exp / 'samples' / 'info.pkl', 'rb') as f:
        assert pickle.load(f) == {}
    for cid in (4, 5, 6):
        files = sorted(p.name for p in (exp / 'samples' / str(cid)).iterdir())
        assert len(files) == 25 and 'sample_0.npz' in files and 'sample_240.npz' in files and 'sample_5.npz' not in files
    samples = load_samples_from_dir(exp / 'samples')
    assert samples['fcn']['layer0']['kernel'].shape == (C, 25, ospec.n_features, 16)
    flat = pm.spec.ravel(samples)                                      # [C, S, d]
    lv = np.stack([[o.forward(ospec, flat[c, s].astype(np.float64), Xt.astype(np.float64)) for s in range(25)]
                   for c in range(C)])
    want = o.lppd(o.pointwise_lppd(ospec, lv, yt.astype(np.float64)))
    assert abs(info['lppd'] - want) <= 1e-4 * abs(want)
    # same seed -> same files (determinism of the whole pipeline)
    info2 = inference_loop(log_post, cfg, 42, posb, np.array([4, 5, 6]), tmp_path / 'exp2' / 'samples')
    s2 = load_samples_from_dir(tmp_path / 'exp2' / 'samples')
    np.testing.assert_array_equal(s2['fcn']['layer1']['kernel'], samples['fcn']['layer1']['kernel'])


def test_statistical_parity_with_cpu_restatement():
    """Full-run parity (north star): same data / split / warm-start / (eps, L); GPU (Philox noise) vs the C
    restatement (its own noise), THREE seeds (the reference replicates over rng: [1, 2, 3],
    experiments/replicate_uci/repl_uci_search.yaml:1-4).  Stated tolerances: LPPD on the test split |mean difference| <=
    3 SE (SE from the chain-to-chain spread, pooled over the seeds) + 0.01; RMSE within 0.05; mean log-density within
    3 SE + 2; effective sample size of the log-density trace within a factor 2; parameter-space split-R-hat (rank
    normalised, 4 splits, median over the parameters) within 0.15."""
    from oracle import c_oracle
    name, C, n, thin = 'airfoil_2x16', 8, 3000, 10
    ospec, module, pm, X, y, Xt, yt = setup_problem(name, n_train=400, n_test=200)
    pm.attach_test_split(Xt, yt)

    def metrics(samples):
        S = samples.shape[0]
        lv = np.stack([[o.forward(ospec, samples[s, c].astype(np.float64), Xt.astype(np.float64)) for s in range(S // 2, S)]
                       for c in range(C)])                               # second half of the run
        pw = o.pointwise_lppd(ospec, lv, yt.astype(np.float64))          # [C, S/2, Nt]
        per_chain = np.array([o.lppd(pw[c:c + 1]) for c in range(C)])
        rmse = np.sqrt(np.mean((lv[..., 0].mean(axis=(0, 1)) - yt) ** 2))
        return o.lppd(pw), per_chain, rmse

    d_lppd, v_lppd, d_lp, v_lp = [], [], [], []
    for seed in (0, 1, 2):
        th0 = o.synthetic_theta0(ospec, C, seed0=1000 + 50 * seed)
        ens = pm.make_ensemble(C, X, y)
        ens.init(th0, seed=1 + 10 * seed)
        ens.tune_reset(0.01)
        tc = ens.tune_cfg(1600, 200, 0.5, 0.1, 1.5, 100)
        ens.tune(1800, 0, tc, seed=2 + 10 * seed)
        ens.tune_finish_phase2()
        eps, L, _ = ens.get_tuning()
        assert np.all(np.isfinite(eps)) and np.all(np.isfinite(L))
        start = ens.get_state()
        g_samples, g_info = ens.sample(n, eps, L, n_thinning=thin, seed=3 + 10 * seed, info=True)
        ens.close()
        ch = c_oracle.Chains(ospec, X, y, start[0], threads=8)
        ch.u[:], ch.lp[:], ch.g[:] = start[1], start[2], start[3]
        c_samples, c_info = ch.sample(n, eps, L, thin=thin, seed=4 + 10 * seed, info=True)
        gl, gpc, grmse = metrics(g_samples)
        cl, cpc, crmse = metrics(c_samples)
        d_lppd.append(gl - cl); v_lppd.append(gpc.var(ddof=1) / C + cpc.var(ddof=1) / C)
        assert abs(grmse - crmse) <= 0.05, (seed, grmse, crmse)
        glp, clp = g_info[n // 2:, :, 0].mean(axis=0), c_info[n // 2:, :, 0].mean(axis=0)
        d_lp.append(glp.mean() - clp.mean()); v_lp.append(glp.var(ddof=1) / C + clp.var(ddof=1) / C)
        # effective sample size of the log-density trace (all chains), same estimator for both
        g_ess = o.ess_rank_normalized(g_info[n // 2:, :, 0].T[:, :, None], rank_normalize=False)
        c_ess = o.ess_rank_normalized(c_info[n // 2:, :, 0].T[:, :, None], rank_normalize=False)
        assert 0.5 <= float(np.mean(g_ess)) / float(np.mean(c_ess)) <= 2.0, (seed, np.mean(g_ess), np.mean(c_ess))
        # parameter-space split-R-hat over the kept samples of the second half: [C, S/2, d]
        half = g_samples.shape[0] // 2 // 4 * 4          # second half, trimmed to a multiple of the 4 splits
        gs, cs = np.transpose(g_samples[-half:], (1, 0, 2)), np.transpose(c_samples[-half:], (1, 0, 2))
        gr, cr = o.split_chain_r_hat(gs, 4), o.split_chain_r_hat(cs, 4)
        assert abs(np.median(gr) - np.median(cr)) <= 0.15, (seed, np.median(gr), np.median(cr))
    k = len(d_lppd)
    se = np.sqrt(np.sum(v_lppd)) / k
    assert abs(np.mean(d_lppd)) <= 3 * se + 0.01, (d_lppd, se)
    se_lp = np.sqrt(np.sum(v_lp)) / k
    assert abs(np.mean(d_lp)) <= 3 * se_lp + 2.0, (d_lp, se_lp)
