"""GPU parity at the BASELINE.json shapes and on the sampler branches round 1 left uncovered:
full-size covertype (232 404 rows), wide 4x256 at 12 165 rows x 8 chains on all three GEMM cores, the 1024-chain
layout, the `with_isokinetic_maruyama` refresh placement (refresh_mode=1), a deterministic handle_nans failure, and
the phase-3 ESS-based L.  Same bar as tests/test_gpu_parity.py: 1e-5 relative against the fp64 oracle."""
import numpy as np
import pytest

from oracle import mile_oracle as o

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(b), 1e-30)


def make(name, C, **opts):
    from mile_b200 import Ensemble, FCNSpec
    ospec = o.make_spec(name)
    X, y, Xt, yt = o.synthetic_data(name)
    ens = Ensemble(FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task), C, **opts)
    ens.set_data(X, y)
    return ospec, ens, X, y, Xt, yt


@pytest.mark.parametrize('G', [12, 1])
def test_covertype_full_value_and_grad_and_step(G):
    """BASELINE configs[2] at its full synthetic size: 232 404 rows x 54 features, 7 classes, [32, 7] sigmoid."""
    C = 3
    ospec, ens, X, y, _, _ = make('covertype_full', C, cluster_size=G)
    assert X.shape == (232404, 54)
    d = ospec.n_params
    th0 = o.synthetic_theta0(ospec, C)
    lp, g = ens.value_and_grad(th0)
    X64 = X.astype(np.float64)
    lp64, g64 = o.logpost_batch(ospec, th0.astype(np.float64), X64, y)
    for c in range(C):
        assert abs(lp[c] - lp64[c]) <= 1e-5 * abs(lp64[c])
        assert rel(g[c], g64[c]) <= 1e-5
    rng = np.random.default_rng(21)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    z = rng.standard_normal((1, C, d)).astype(np.float32)
    eps, L = 2e-4, float(np.sqrt(d))
    ens.init(th0, z0)
    _, info = ens.sample(1, eps, L, z=z, keep=False, info=True)
    th, u, lpn, gn = ens.get_state()
    f64 = lambda t: o.logpost_value_and_grad(ospec, t, X64, y)
    for c in range(C):
        st = o.mclmc_init(f64, th0[c].astype(np.float64), z0[c].astype(np.float64))
        st, inf = o.mclmc_step(f64, st, eps, L, z[0, c].astype(np.float64))
        scale = abs(st.logdensity)
        assert rel(th[c], st.position) <= 1e-5
        assert rel(u[c], st.momentum) <= 1e-5
        assert abs(lpn[c] - st.logdensity) <= 1e-5 * scale
        assert rel(gn[c], st.logdensity_grad) <= 1e-5
        assert abs(info[0, c, 2] - inf.energy_change) <= 1e-5 * scale
    ens.close()


def test_wide_4x256_full_shape_all_tensor_cores():
    """BASELINE configs[3]: [256,256,256,256,2] relu on the bikesharing shape, 12 165 rows x 8 chains; SIMT (0),
    register-staged tcgen05 (1), TMA-fed tcgen05 (2) and TMA-fed CTA-pair tcgen05 (3, the default) cores against the fp64 oracle."""
    from mile_b200 import Ensemble, FCNSpec
    C = 8
    ospec = o.make_spec('wide_4x256')
    X, y, _, _ = o.synthetic_data('wide_4x256')
    assert X.shape[0] == 12165
    d = ospec.n_params
    th = o.synthetic_theta0(ospec, C, scale=0.05)
    lp64, g64 = o.logpost_batch(ospec, th.astype(np.float64), X.astype(np.float64), y)
    _, g32 = o.logpost_batch(ospec, th, X, y)

    worst = {}
    for tensor in (3, 2, 1, 0):
        ens = Ensemble(FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task), C, tensor=tensor)
        ens.set_data(X, y)
        assert ens.get_option('wide') == 1
        lp, g = ens.value_and_grad(th)
        worst[tensor] = max(rel(g[c], g64[c]) for c in range(C))
        for c in range(C):
            assert abs(lp[c] - lp64[c]) <= 1e-5 * abs(lp64[c]), (tensor, c)
            # At 12 165 rows x 256 units some pre-activations sit within fp32 rounding of the ReLU kink: the literal fp32
            # restatement itself lands up to 2e-5 from its fp64 twin there (chain 0), so that distance is allowed on top.
            # The SIMT core holds the 2e-5 bar.  The tensor cores accumulate in fp32 with round-toward-zero: ~100 MMAs per
            # accumulator (K = 256 x 3 passes, or one 330-row split-K slice) leave a SYSTEMATIC ~6e-6 relative error on
            # every delta / dW element, which the bias gradients (column sums of delta over 12 165 rows with heavy
            # cancellation) amplify; measured 2.0e-4 norm-wise, bounded here at 3e-4.
            #
            # ReLU kink: every chain has 2-12 (row, unit) pairs whose pre-activation lies within 1e-7 relative of zero;
            # which side an fp32 implementation takes there is decided by its summation order, and ONE flip moves the
            # gradient blocks of the layers below it by about a row's share of the sum (measured 1.2e-4 on those blocks,
            # 6e-5 on the whole gradient; tools/wide_grad_blocks.py -- chains without a flip sit at 1e-7 on the SIMT core).
            # A flip is chain-specific, an arithmetic error is not: the MEDIAN over the chains must hold the tight bar,
            # every chain the bar with 2e-4 of kink allowance.
            tol = (2e-5 if tensor == 0 else 3e-4) + 2 * rel(g32[c], g64[c]) + 2e-4
            assert rel(g[c], g64[c]) <= tol, (tensor, c, rel(g[c], g64[c]), rel(g32[c], g64[c]))
        errs = sorted(rel(g[c], g64[c]) for c in range(C))
        assert errs[C // 2] <= (2e-5 if tensor == 0 else 3e-4), (tensor, errs)
        if tensor == 3:
            # what the sampler consumes: one MCLMC step at the full shape holds the 1e-5 bar on position and log-density
            rng = np.random.default_rng(3)
            z0 = rng.standard_normal((C, d)).astype(np.float32)
            z = rng.standard_normal((1, C, d)).astype(np.float32)
            ens.init(th, z0)
            _, info = ens.sample(1, 0.01, float(np.sqrt(d)), z=z, keep=False, info=True)
            thg, ug, lpg, _ = ens.get_state()
            f64 = lambda t: o.logpost_value_and_grad(ospec, t, X.astype(np.float64), y)
            for c in (0, 1):
                st = o.mclmc_init(f64, th[c].astype(np.float64), z0[c].astype(np.float64))
                st, inf = o.mclmc_step(f64, st, 0.01, float(np.sqrt(d)), z[0, c].astype(np.float64))
                assert rel(thg[c], st.position) <= 1e-5
                assert rel(ug[c], st.momentum) <= 1e-4
                assert abs(lpg[c] - st.logdensity) <= 1e-5 * abs(st.logdensity)
                assert abs(info[0, c, 2] - inf.energy_change) <= 2e-5 * abs(st.logdensity)
        ens.close()
    print('wide 4x256 full shape: worst gradient rel error per core', worst)


def test_1024_chains_single_step():
    """BASELINE configs[4]: 1024 airfoil chains, one CTA per chain (G = 1): every chain finite, a spread of chains
    checked field by field against the fp64 oracle."""
    C = 1024
    ospec, ens, X, y, _, _ = make('airfoil_3x16', C)
    assert ens.get_option('cluster_size') == 1
    d = ospec.n_params
    th0 = o.synthetic_theta0(ospec, C)
    rng = np.random.default_rng(31)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    z = rng.standard_normal((1, C, d)).astype(np.float32)
    eps, L = 0.01, float(np.sqrt(d))
    ens.init(th0, z0)
    _, info = ens.sample(1, eps, L, z=z, keep=False, info=True)
    th, u, lp, g = ens.get_state()
    assert np.all(np.isfinite(th)) and np.all(np.isfinite(u)) and np.all(np.isfinite(g)) and np.all(np.isfinite(lp))
    f64 = lambda t: o.logpost_value_and_grad(ospec, t, X.astype(np.float64), y)
    for c in (0, 1, 63, 64, 147, 148, 511, 512, 1000, 1023):
        st = o.mclmc_init(f64, th0[c].astype(np.float64), z0[c].astype(np.float64))
        st, inf = o.mclmc_step(f64, st, eps, L, z[0, c].astype(np.float64))
        scale = abs(st.logdensity)
        assert rel(th[c], st.position) <= 1e-5, c
        assert rel(u[c], st.momentum) <= 1e-5, c
        assert abs(lp[c] - st.logdensity) <= 1e-5 * scale, c
        assert abs(info[0, c, 2] - inf.energy_change) <= 1e-5 * scale, c
    ens.close()


@pytest.mark.parametrize('name,G,fast', [('airfoil_3x16', 1, 2), ('airfoil_3x16', 8, 2), ('airfoil_3x16', 8, 1), ('airfoil_3x16', 8, 0),
                                         ('bikesharing_2x16', 8, 2), ('bikesharing_2x16', 14, 1), ('covertype_ref', 4, 0),
                                         ('protein_2x16', 8, 2), ('airfoil_2x16', 2, 1), ('covertype_ref', 11, 0),
                                         ('bikesharing_2x16', 0, 2)])
def test_single_step_parity_maruyama_refresh(name, G, fast):
    """refresh_mode=1: half-step partial refreshes around the integrator (blackjax `with_isokinetic_maruyama`), the
    alternative reading of SURVEY.md Appendix A.  Two host-supplied noise vectors per step: z [n_steps, 2, C, d]."""
    C = 3
    ospec, ens, X, y, _, _ = make(name, C, cluster_size=G, fast=fast, refresh_mode=1)
    d = ospec.n_params
    th0 = o.synthetic_theta0(ospec, C)
    rng = np.random.default_rng(13)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    nstep = 2
    z = rng.standard_normal((nstep, 2, C, d)).astype(np.float32)
    eps, L = 0.01, float(np.sqrt(d))
    ens.init(th0, z0)
    _, info = ens.sample(nstep, eps, L, z=z, keep=False, info=True)
    th, u, lp, g = ens.get_state()
    f64 = lambda t: o.logpost_value_and_grad(ospec, t, X.astype(np.float64), y)
    for c in range(C):
        st = o.mclmc_init(f64, th0[c].astype(np.float64), z0[c].astype(np.float64))
        for s in range(nstep):
            st, inf = o.mclmc_step(f64, st, eps, L, z[s, :, c].astype(np.float64), refresh='maruyama')
            scale = abs(st.logdensity)
            assert abs(info[s, c, 0] - st.logdensity) <= 1e-5 * scale
            assert abs(info[s, c, 1] - inf.kinetic_change) <= 1e-5 * scale
            assert abs(info[s, c, 2] - inf.energy_change) <= 1e-5 * scale
        assert rel(th[c], st.position) <= 1e-5
        assert rel(u[c], st.momentum) <= 2e-5
        assert abs(lp[c] - st.logdensity) <= 1e-5 * scale
        # the post-step refresh of the same noise gives a DIFFERENT momentum: the switch is observable
        st_post = o.mclmc_init(f64, th0[c].astype(np.float64), z0[c].astype(np.float64))
        st_post, _ = o.mclmc_step(f64, st_post, eps, L, z[0, 1, c].astype(np.float64), refresh='post')
        assert rel(st_post.momentum, st.momentum) > 1e-3
    ens.close()


@pytest.mark.parametrize('fast', [2, 1, 0])
def test_handle_nans_failure_matches_oracle(fast):
    """warmup.py:468-483 with a deterministic failure: chain 1's cached gradient is poisoned (inf), so its first
    B-step yields a NaN momentum and a non-finite position.  Every field is compared with o.tune_step."""
    name, C = 'airfoil_2x16', 2
    ospec, ens, X, y, _, _ = make(name, C, fast=fast)
    d = ospec.n_params
    th0 = o.synthetic_theta0(ospec, C)
    rng = np.random.default_rng(4)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    z = rng.standard_normal((2, C, d)).astype(np.float32)
    cfg = o.TuneConfig(2, 0, 0, 0.5, 0.1, 1.5, 100, 0.01)
    tc = ens.tune_cfg(2, 0, 0.5, 0.1, 1.5, 100)
    ens.init(th0, z0)
    ens.tune_reset(0.01)
    th_b, u_b, lp_b, g_b = ens.get_state()
    g_bad = g_b.copy()
    g_bad[1, 0] = np.inf
    ens.set_state(grad=g_bad)
    info = ens.tune(1, 0, tc, z=z[:1], info=True)
    th_a, u_a, lp_a, g_a = ens.get_state()
    eps, L, emax = ens.get_tuning()
    f64 = lambda t: o.logpost_value_and_grad(ospec, t, X.astype(np.float64), y)
    with np.errstate(all='ignore'):
        for c in range(C):
            st = o.IntegratorState(th_b[c].astype(np.float64), u_b[c].astype(np.float64), np.float64(lp_b[c]),
                                   g_bad[c].astype(np.float64))
            st2, ts, inf, ok = o.tune_step(f64, cfg, st, o.tune_init(cfg, d, np.float64), z[0, c].astype(np.float64), 0)
            assert info[0, c, 3] == (1.0 if ok else 0.0)
            assert ok == (c == 0)
            if ok:
                assert rel(th_a[c], st2.position) <= 1e-5
                assert rel(u_a[c], st2.momentum) <= 1e-5
                assert abs(lp_a[c] - st2.logdensity) <= 1e-5 * abs(st2.logdensity)
                assert abs(info[0, c, 0] - inf.energy_change) <= 1e-5 * abs(st2.logdensity)
                # jnp.nan_to_num(step_size_max = inf) -> largest finite float (warmup.py:480)
                assert emax[c] == np.finfo(np.float32).max and ts.step_size_max == np.finfo(np.float64).max
                assert info[0, c, 2] == np.finfo(np.float32).max
            else:
                # previous state kept bit for bit (including the poisoned gradient), eps_max = 0.8 eps, dE = 0
                np.testing.assert_array_equal(th_a[c], th_b[c])
                np.testing.assert_array_equal(u_a[c], u_b[c])
                np.testing.assert_array_equal(g_a[c], g_bad[c])
                assert lp_a[c] == lp_b[c]
                assert info[0, c, 0] == 0.0 and inf.energy_change == 0.0
                assert abs(emax[c] - ts.step_size_max) <= 1e-6 * ts.step_size_max
                assert abs(emax[c] - 0.8 * 0.01) <= 1e-6
                assert abs(info[0, c, 2] - ts.step_size_max) <= 1e-6 * ts.step_size_max
            if ok:
                # dE at eps = 0.01 is of the size of the fp32 rounding of the log-density, so the predictor is checked
                # by driving the oracle's update with the GPU's own energy change (as test_tuning_matches_oracle does)
                ts = o.tune_update(cfg, o.tune_init(cfg, d, np.float64)._replace(step_size_max=ts.step_size_max),
                                   th_a[c].astype(np.float64), np.float64(info[0, c, 0]), True, 0)
            assert abs(eps[c] - ts.step_size) <= 2e-5 * ts.step_size, (c, eps[c], ts.step_size)
            assert abs(info[0, c, 1] - ts.step_size) <= 2e-5 * ts.step_size
    ens.close()


def test_phase3_L_matches_oracle_on_same_positions():
    """warmup.py:408-465: L = 0.4 eps mean(n / ESS) from the tune3 positions.  The positions come from the GPU
    sampler; the FFT-based ESS of the host mirror (torch.fft on the device) is checked against the oracle's numpy
    restatement of blackjax.diagnostics.effective_sample_size on those same positions."""
    from mile_b200.warmup import phase3_L
    name, C, n3 = 'airfoil_2x16', 3, 600
    ospec, ens, X, y, _, _ = make(name, C)
    th0 = o.synthetic_theta0(ospec, C)
    ens.init(th0, seed=5)
    eps = np.asarray([0.02, 0.03, 0.015], np.float32)
    pos, _ = ens.sample(n3, eps, 8.0, n_thinning=1, seed=9)
    assert pos.shape == (n3, C, ospec.n_params)
    L = phase3_L(pos, eps, device=ens.device)
    for c in range(C):
        want = o.adaptation_L(np.float64(eps[c]), pos[:, c].astype(np.float64))
        assert abs(L[c] - want) <= 1e-3 * want, (c, L[c], want)
    ens.close()


def test_n_batches_scales_value_and_gradient_consistently():
    """probabilistic.py:136: log-posterior = prior + n_batches * log-likelihood, in the value as well as the gradient
    (the reference always passes n_batches = 1)."""
    from mile_b200 import Ensemble, FCNSpec
    rng = np.random.default_rng(8)
    for widths, fast in (((16, 16, 2), 2), ((16, 16, 2), 1), ((16, 16, 2), 0), ((9, 6, 2), 0), ((272, 264, 2), 0)):
        F, N, C = 7, 333, 2
        ospec = o.ModelSpec(F, widths, 'relu', 'regr', n_batches=3.0)
        X = rng.standard_normal((N, F)).astype(np.float32)
        y = rng.standard_normal(N).astype(np.float32)
        th = (rng.standard_normal((C, ospec.n_params)) * (0.4 if max(widths) < 64 else 0.03)).astype(np.float32)
        ens = Ensemble(FCNSpec(F, widths, 'relu', 'regr', n_batches=3.0), C, fast=fast)
        ens.set_data(X, y)
        lp, g = ens.value_and_grad(th)
        lp64, g64 = o.logpost_batch(ospec, th.astype(np.float64), X.astype(np.float64), y)
        for c in range(C):
            assert abs(lp[c] - lp64[c]) <= 1e-5 * abs(lp64[c]), (widths, fast)
            assert rel(g[c], g64[c]) <= 2e-5, (widths, fast)
        ens.close()


def test_chain_base_decorrelates_partitioned_ensembles():
    """two ranks that own different blocks of chains must not draw the same Philox noise.  With
    chain_base = 0 on both, identical starts give identical trajectories; with chain_base = C on the second they differ,
    and chain_base = k reproduces chain k of a larger single-process ensemble bit for bit."""
    ospec, ens_a, X, y, _, _ = make('airfoil_2x16', 2)
    _, ens_b, _, _, _, _ = make('airfoil_2x16', 2)
    _, ens_c, _, _, _, _ = make('airfoil_2x16', 2, chain_base=2)
    _, ens_d, _, _, _, _ = make('airfoil_2x16', 4)
    th0 = o.synthetic_theta0(ospec, 4)
    th2 = np.concatenate([th0[:2]])
    outs = []
    for ens, th in ((ens_a, th0[:2]), (ens_b, th0[:2]), (ens_c, th0[2:]), (ens_d, th0)):
        ens.init(th, seed=3)
        ens.sample(5, 0.02, 20.0, seed=7, keep=False)
        outs.append(ens.get_state())
        ens.close()
    np.testing.assert_array_equal(outs[0][0], outs[1][0])            # same chain ids -> same noise
    np.testing.assert_array_equal(outs[3][0][:2], outs[0][0])        # block 0 of the 4-chain ensemble
    np.testing.assert_array_equal(outs[3][0][2:], outs[2][0])        # block 1 == the chain_base = 2 context
    np.testing.assert_array_equal(outs[3][1][2:], outs[2][1])


def test_wide_path_predict_and_lppd_match_oracle():
    """models on the HBM-resident wide path (here 2 x 256) are evaluated through the same entry points as the
    shared-memory ones: mile_predict, mile_lppd_accumulate and the lppd=1 fold during sampling."""
    from mile_b200 import Ensemble, FCNSpec, lppd_from_state
    from mile_b200.evaluation import _forward_rows
    F, N, Nt, C, S = 12, 700, 333, 3, 3
    widths = (256, 256, 2)
    ospec = o.ModelSpec(F, widths, 'relu', 'regr')
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, F)).astype(np.float32); y = rng.standard_normal(N).astype(np.float32)
    Xt = rng.standard_normal((Nt, F)).astype(np.float32); yt = rng.standard_normal(Nt).astype(np.float32)
    d = ospec.n_params
    thetas = (rng.standard_normal((S, C, d)) * 0.04).astype(np.float32)
    spec = FCNSpec(F, widths, 'relu', 'regr')
    ens = Ensemble(spec, C)
    ens.set_data(X, y); ens.set_test(Xt, yt)
    assert ens.get_option('wide') == 1
    out = ens.predict(thetas[0], 'test')
    ref = np.stack([o.forward(ospec, thetas[0, c].astype(np.float64), Xt.astype(np.float64)) for c in range(C)])
    assert rel(out, ref) <= 1e-5
    assert rel(_forward_rows(spec, thetas.reshape(-1, d), Xt, None, 4096), np.stack(
        [o.forward(ospec, t.astype(np.float64), Xt.astype(np.float64)) for t in thetas.reshape(-1, d)])) <= 1e-5
    for s in range(S):
        ens.lppd_accumulate(thetas[s])
    m, sst, cnt = ens.lppd_state()
    assert cnt == S
    lv = np.stack([[o.forward(ospec, thetas[s, c].astype(np.float64), Xt.astype(np.float64)) for s in range(S)] for c in range(C)])
    want = o.lppd(o.pointwise_lppd(ospec, lv, yt.astype(np.float64)))
    assert abs(lppd_from_state(m, sst, C * S) - want) <= 1e-5 * abs(want)
    # fused fold during sampling == post-hoc fold of the kept positions
    ens.lppd_reset()
    ens.init(thetas[0], seed=3)
    samples, _ = ens.sample(6, 0.005, float(np.sqrt(d)), n_thinning=3, seed=7, lppd=True)
    m, sst, cnt = ens.lppd_state()
    assert cnt == 2 and samples.shape[0] == 2
    lv = np.stack([[o.forward(ospec, samples[k, c].astype(np.float64), Xt.astype(np.float64)) for k in range(2)] for c in range(C)])
    want = o.lppd(o.pointwise_lppd(ospec, lv, yt.astype(np.float64)))
    assert abs(lppd_from_state(m, sst, C * 2) - want) <= 1e-5 * abs(want)
    ens.close()


def test_device_ess_kernel_matches_oracle():
    """csrc/mile_ess.cuh (transpose + one CTA per series: lazy direct autocovariance, Geyer initial positive / monotone
    sequence) against the oracle's restatement of blackjax.diagnostics.effective_sample_size (FFT autocovariance, fp64),
    one chain per series as warmup.py:457 calls it; with and without the reference's parameter / sample subsampling."""
    from mile_b200 import Ensemble, FCNSpec
    spec = FCNSpec(3, (4, 2), 'identity', 'regr')
    C, d = 2, spec.n_params
    ens = Ensemble(spec, C)
    rng = np.random.default_rng(12)
    for n in (601, 1200):
        phi = np.concatenate([np.linspace(-0.6, 0.995, d - 1), [0.0]])
        x = np.zeros((n, C, d))
        e = rng.standard_normal((n, C, d))
        for t in range(1, n):
            x[t] = phi * x[t - 1] + e[t]
        x = (x + 3.0 * rng.standard_normal((1, C, d))).astype(np.float32)
        ess = ens.ess_positions(x)
        assert ess.shape == (C, d)
        for c in range(C):
            want = o.effective_sample_size(x[:, c].astype(np.float64)[None])
            assert np.all(np.abs(ess[c] - want) <= 2e-4 * want), (n, c, np.max(np.abs(ess[c] / want - 1)))
        # all chains pooled per parameter (the report's ESS, metrics.py:354-425): with and without chain-specific offsets
        for xx in (x, x - x.mean(axis=0, keepdims=True) * 0.9):
            pooled = ens.ess_positions(xx, pooled=True)
            want = o.effective_sample_size(np.transpose(xx, (1, 0, 2)).astype(np.float64))
            assert pooled.shape == (d,) and np.all(np.abs(pooled - want) <= 5e-4 * want), np.max(np.abs(pooled / want - 1))
        pidx = np.array([7, 0, 25, 3], np.int32)
        sidx = np.linspace(0, n - 1, 400).astype(np.int32)
        sub = ens.ess_positions(x, param_idx=pidx, sample_idx=sidx)
        assert sub.shape == (C, 4)
        for c in range(C):
            want = o.effective_sample_size(x[sidx][:, c][:, pidx].astype(np.float64)[None])
            assert np.all(np.abs(sub[c] - want) <= 2e-4 * want)
    ens.close()


def test_phase3_on_device_equals_oracle_L_on_the_same_positions():
    """mile_mclmc_phase3_ess (capture in HBM + ESS kernels, no host copy of the positions) gives the L the oracle derives from
    the very same positions (re-sampled from the same state with the same Philox key): warmup.py:408-465."""
    name, C, n3 = 'airfoil_2x16', 3, 600
    ospec, ens, X, y, _, _ = make(name, C)
    th0 = o.synthetic_theta0(ospec, C)
    ens.init(th0, seed=5)
    ens.sample(200, 0.02, 8.0, seed=1, keep=False)
    state = ens.get_state()
    eps = np.asarray([0.02, 0.03, 0.015], np.float32)
    Ls = np.asarray([8.0, 6.0, 9.0], np.float32)
    ess = ens.phase3_ess(n3, eps, Ls, seed=9)
    after = ens.get_state()
    ens.set_state(*state)
    pos, _ = ens.sample(n3, eps, Ls, n_thinning=1, seed=9)
    np.testing.assert_array_equal(ens.get_state()[0], after[0])          # phase 3 advanced the chains by the same steps
    for c in range(C):
        want = o.adaptation_L(np.float64(eps[c]), pos[:, c].astype(np.float64))
        got = 0.4 * eps[c] * np.mean(n3 / ess[c].astype(np.float64))
        assert abs(got - want) <= 1e-3 * want, (c, got, want)
    ens.close()


def test_wide_path_full_warmup_with_device_phase3():
    """custom_mclmc_warmup on a model of the HBM-resident wide path (2 x 256, d = 69 890 > the 2000-parameter limit of
    warmup.py:442-449): phases 1+2 in the tuning loop, phase 3 captured in HBM with the ESS kernels on a random subset of
    2000 parameters; the L it returns is the oracle's L on the very same positions and parameter subset."""
    from mile_b200 import Ensemble, FCNSpec
    from mile_b200.warmup import run_warmup
    F, N, C = 12, 700, 2
    widths = (256, 256, 2)
    rng = np.random.default_rng(4)
    X = rng.standard_normal((N, F)).astype(np.float32); y = rng.standard_normal(N).astype(np.float32)
    spec = FCNSpec(F, widths, 'relu', 'regr')
    ens = Ensemble(spec, C)
    ens.set_data(X, y)
    assert ens.get_option('wide') == 1
    th0 = (rng.standard_normal((C, spec.n_params)) * 0.04).astype(np.float32)
    eps, L = run_warmup(ens, th0, 7, 400, desired_energy_var_start=0.5, desired_energy_var_end=0.1, trust_in_estimate=1.5,
                        num_effective_samples=100, step_size_init=0.005)
    assert eps.shape == L.shape == (C,) and np.all(np.isfinite(eps)) and np.all(eps > 0) and np.all(np.isfinite(L)) and np.all(L > 0)
    # the ESS table of a fresh phase 3 against the oracle on the same positions / subset
    state = ens.get_state()
    pidx = np.sort(rng.permutation(spec.n_params)[:50]).astype(np.int32)
    ess = ens.phase3_ess(40, eps, L, seed=3, param_idx=pidx)
    ens.set_state(*state)
    pos, _ = ens.sample(40, eps, L, n_thinning=1, seed=3)
    for c in range(C):
        want = o.effective_sample_size(pos[:, c][:, pidx].astype(np.float64)[None])
        assert np.all(np.abs(ess[c] - want) <= 1e-3 * want), (c, np.max(np.abs(ess[c] / want - 1)))
    ens.close()
