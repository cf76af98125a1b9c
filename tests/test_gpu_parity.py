"""GPU parity tests proper: the CUDA path (through the C ABI) against the numpy oracle on the
same seeded inputs.  Tolerances follow the north star: 1e-5 relative in fp32 for single-step
positions / momenta / energies (energies relative to |logdensity|)."""
import numpy as np
import pytest

from oracle import mile_oracle as o

pytestmark = pytest.mark.gpu

CONFIGS = ['airfoil_3x16', 'airfoil_2x16', 'bikesharing_2x16', 'protein_2x16', 'covertype_ref']


def make(name, C, n_train=None, n_test=None, **opts):
    from mile_b200 import Ensemble, FCNSpec
    ospec = o.make_spec(name)
    X, y, Xt, yt = o.synthetic_data(name, n_train=n_train, n_test=n_test)
    ens = Ensemble(FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task), C, **opts)
    ens.set_data(X, y)
    return ospec, ens, X, y, Xt, yt


def rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize('name', CONFIGS)
@pytest.mark.parametrize('G', [1, 2, 8, 14])   # 14: global-memory exchange + cooperative launch (no cluster)
@pytest.mark.parametrize('fast', [2, 1, 0])
def test_value_and_grad_matches_oracle(name, G, fast):
    """fast=2: register-chained 3xTF32 tensor evaluator, fast=1: warp-specialised FFMA pipeline (both where the shape is
    eligible: hidden width 16, Gaussian head), fast=0: generic tile kernel."""
    C = 3
    ospec, ens, X, y, _, _ = make(name, C, cluster_size=G, fast=fast)
    if fast:
        assert ens.get_option('fast') == (fast if name != 'covertype_ref' else 0)
    th = o.synthetic_theta0(ospec, C)
    lp, g = ens.value_and_grad(th)
    lp64, g64 = o.logpost_batch(ospec, th.astype(np.float64), X.astype(np.float64), y)
    lp32, g32 = o.logpost_batch(ospec, th, X, y)
    for c in range(C):
        assert abs(lp[c] - lp64[c]) <= 1e-5 * abs(lp64[c])
        assert rel(g[c], g64[c]) <= 1e-5
        # the fp32 numpy oracle can land on the other side of a ReLU kink (z ~ 1e-8) for a row;
        # allow exactly the distance the fp32 oracle itself has from its fp64 twin
        assert rel(g[c], g32[c]) <= 1e-5 + 2 * rel(g32[c], g64[c])
    ens.close()


@pytest.mark.parametrize('act', ['relu', 'sigmoid', 'tanh', 'gelu', 'leaky_relu', 'identity'])
@pytest.mark.parametrize('task,prior', [('regr', 'normal'), ('class', 'laplace')])
def test_value_and_grad_activations_tasks_priors(act, task, prior):
    from mile_b200 import Ensemble, FCNSpec
    K = 2 if task == 'regr' else 5
    kw = dict(prior_loc=0.1, prior_scale=1.7)
    ospec = o.ModelSpec(7, (9, 6, K), act, task, prior, **kw)
    rng = np.random.default_rng(3)
    X = rng.standard_normal((333, 7)).astype(np.float32)
    y = rng.standard_normal(333).astype(np.float32) if task == 'regr' else rng.integers(0, K, 333).astype(np.int32)
    th = (rng.standard_normal((2, ospec.n_params)) * 0.4).astype(np.float32)
    ens = Ensemble(FCNSpec(7, (9, 6, K), act, task, prior, **kw), 2)
    ens.set_data(X, y)
    lp, g = ens.value_and_grad(th)
    lp64, g64 = o.logpost_batch(ospec, th.astype(np.float64), X.astype(np.float64), y)
    for c in range(2):
        assert abs(lp[c] - lp64[c]) <= 1e-5 * abs(lp64[c])
        assert rel(g[c], g64[c]) <= 1e-5
    ens.close()


def test_ragged_and_tiny_row_counts():
    """Edge cases: N not a multiple of anything, N smaller than a tile, N = 1."""
    for N, fast in [(n, f) for n in (1, 3, 15, 16, 17, 31, 33, 63, 64, 65, 257, 1000) for f in (2, 1, 0)]:
        ospec, ens, X, y, _, _ = make('airfoil_2x16', 2, n_train=N, fast=fast)
        th = o.synthetic_theta0(ospec, 2)
        lp, g = ens.value_and_grad(th)
        lp64, g64 = o.logpost_batch(ospec, th.astype(np.float64), X.astype(np.float64), y)
        for c in range(2):
            assert abs(lp[c] - lp64[c]) <= 1e-5 * abs(lp64[c]), N
            assert rel(g[c], g64[c]) <= 1e-5, N
        ens.close()


def test_deep_narrow_network():
    """feasibility/feas.yaml is 9x16 + head (10 layers): exercises NLMAX=12 and the
    lexicographic leaf order ('layer10' < 'layer2' does not occur at 10 layers, does at 12)."""
    from mile_b200 import Ensemble, FCNSpec
    widths = (8,) * 11 + (2,)
    ospec = o.ModelSpec(5, widths, 'tanh', 'regr')
    rng = np.random.default_rng(0)
    X = rng.standard_normal((200, 5)).astype(np.float32)
    y = rng.standard_normal(200).astype(np.float32)
    th = (rng.standard_normal((2, ospec.n_params)) * 0.4).astype(np.float32)
    ens = Ensemble(FCNSpec(5, widths, 'tanh', 'regr'), 2)
    ens.set_data(X, y)
    lp, g = ens.value_and_grad(th)
    lp64, g64 = o.logpost_batch(ospec, th.astype(np.float64), X.astype(np.float64), y)
    for c in range(2):
        assert abs(lp[c] - lp64[c]) <= 1e-5 * abs(lp64[c])
        assert rel(g[c], g64[c]) <= 2e-5
    ens.close()


@pytest.mark.parametrize('name,G,fast', [('airfoil_3x16', 1, 2), ('airfoil_3x16', 8, 2), ('bikesharing_2x16', 8, 2), ('protein_2x16', 4, 2),
                                         ('airfoil_2x16', 2, 2), ('bikesharing_2x16', 14, 2), ('airfoil_3x16', 12, 2),
                                         ('airfoil_3x16', 1, 1), ('airfoil_3x16', 8, 1), ('airfoil_3x16', 8, 0),
                                         ('bikesharing_2x16', 8, 1), ('bikesharing_2x16', 8, 0),
                                         ('covertype_ref', 4, 0), ('protein_2x16', 8, 1), ('airfoil_2x16', 2, 1),
                                         ('bikesharing_2x16', 14, 1), ('covertype_ref', 11, 0), ('bikesharing_2x16', 0, 1)])
def test_single_step_parity(name, G, fast):
    """The north-star criterion: identical (theta,u,l,g,eps,L,z) in -> (theta',u',l',dK,dE) within 1e-5."""
    C = 3
    ospec, ens, X, y, _, _ = make(name, C, cluster_size=G, fast=fast)
    d = ospec.n_params
    th0 = o.synthetic_theta0(ospec, C)
    rng = np.random.default_rng(11)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    z = rng.standard_normal((1, C, d)).astype(np.float32)
    eps, L = 0.01, float(np.sqrt(d))
    ens.init(th0, z0)
    th_i, u_i, lp_i, g_i = ens.get_state()
    _, info = ens.sample(1, eps, L, z=z, keep=False, info=True)
    th, u, lp, g = ens.get_state()
    f64 = lambda t: o.logpost_value_and_grad(ospec, t, X.astype(np.float64), y)
    f32 = lambda t: o.logpost_value_and_grad(ospec, t, X, y)
    for c in range(C):
        s64 = o.mclmc_init(f64, th0[c].astype(np.float64), z0[c].astype(np.float64))
        assert rel(u_i[c], s64.momentum) <= 1e-6
        assert rel(g_i[c], s64.logdensity_grad) <= 1e-5
        n64, i64 = o.mclmc_step(f64, s64, eps, L, z[0, c].astype(np.float64))
        s32 = o.mclmc_init(f32, th0[c], z0[c])
        n32, i32 = o.mclmc_step(f32, s32, eps, L, z[0, c])
        scale = abs(n64.logdensity)
        # vs the fp64 twin: tight.  vs the literal fp32 restatement: allow the distance that fp32
        # restatement itself has from its twin (ReLU kinks, cancellation in 1 - exp(-delta)).
        for ref, inf, su, sg in ((n64, i64, 0.0, 0.0),
                                 (n32, i32, 2 * rel(n32.momentum, n64.momentum),
                                  2 * rel(n32.logdensity_grad, n64.logdensity_grad))):
            assert rel(th[c], ref.position) <= 1e-5
            assert rel(u[c], ref.momentum) <= 1e-5 + su
            assert abs(lp[c] - ref.logdensity) <= 1e-5 * scale
            assert rel(g[c], ref.logdensity_grad) <= 1e-4 + sg
            assert abs(info[0, c, 0] - ref.logdensity) <= 1e-5 * scale
        # energies: compare with the fp64 twin, relative to the energy scale |l|
        assert abs(info[0, c, 1] - i64.kinetic_change) <= 1e-5 * scale
        assert abs(info[0, c, 2] - i64.energy_change) <= 1e-5 * scale
    ens.close()


def test_multi_step_trajectory_and_thinning():
    """20 steps with host noise: trajectory tracks the fp64 oracle; kept positions obey idx % n_thinning == 0."""
    name, C, n, thin = 'airfoil_3x16', 2, 20, 4
    ospec, ens, X, y, _, _ = make(name, C)
    d = ospec.n_params
    th0 = o.synthetic_theta0(ospec, C)
    rng = np.random.default_rng(5)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    z = rng.standard_normal((n, C, d)).astype(np.float32)
    ens.init(th0, z0)
    samples, info = ens.sample(n, 0.02, 20.0, z=z, n_thinning=thin, info=True)
    assert samples.shape == (5, C, d)
    f64 = lambda t: o.logpost_value_and_grad(ospec, t, X.astype(np.float64), y)
    for c in range(C):
        st = o.mclmc_init(f64, th0[c].astype(np.float64), z0[c].astype(np.float64))
        st, kept, idxs, des = o.run_sampling(f64, st, 0.02, 20.0, z[:, c].astype(np.float64), n_thinning=thin)
        assert idxs == [0, 4, 8, 12, 16]
        for k in range(5):
            assert rel(samples[k, c], kept[k]) <= 2e-5
        assert np.max(np.abs(info[:, c, 2] - des)) <= 2e-5 * abs(st.logdensity)
    # split launches == one launch (step_base bookkeeping)
    ens.init(th0, z0)
    s1, _ = ens.sample(7, 0.02, 20.0, z=z[:7], n_thinning=thin, step_base=0)
    s2, _ = ens.sample(13, 0.02, 20.0, z=z[7:], n_thinning=thin, step_base=7)
    np.testing.assert_array_equal(np.concatenate([s1, s2]), samples)
    ens.close()


def test_cluster_sizes_agree_bitwise_on_state_layout():
    """Different cluster sizes change the summation order only: results agree to fp32 rounding."""
    name, C = 'bikesharing_2x16', 2
    outs = []
    for G in (1, 2, 4, 8, 13):
        ospec, ens, X, y, _, _ = make(name, C, cluster_size=G)
        th0 = o.synthetic_theta0(ospec, C)
        rng = np.random.default_rng(1)
        z0 = rng.standard_normal((C, ospec.n_params)).astype(np.float32)
        z = rng.standard_normal((3, C, ospec.n_params)).astype(np.float32)
        ens.init(th0, z0)
        ens.sample(3, 0.01, 22.0, z=z, keep=False)
        outs.append(ens.get_state())
        ens.close()
    for k in range(1, 5):
        assert rel(outs[k][0], outs[0][0]) <= 1e-6
        assert rel(outs[k][1], outs[0][1]) <= 1e-5


def test_tuning_matches_oracle():
    """HOT LOOP A (warmup.py:276-352).  dE at small step sizes is dominated by fp32 rounding of the
    log-density (|l| ~ 2e3 -> ~2e-4 absolute) in the reference as well, so the adaptive arithmetic is
    checked exactly: the oracle's predictor / streaming average is driven with the GPU's own per-step
    energy changes and positions and must reproduce the GPU's eps trajectory, moments and L."""
    name, C = 'airfoil_3x16', 2
    ospec, ens, X, y, _, _ = make(name, C)
    d = ospec.n_params
    th0 = o.synthetic_theta0(ospec, C)
    rng = np.random.default_rng(9)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    t1, t2 = 24, 16
    n = t1 + t2
    z = rng.standard_normal((n, C, d)).astype(np.float32)
    cfg = o.TuneConfig(t1, t2, 0, 0.5, 0.1, 1.5, 100, 0.01)
    tc = ens.tune_cfg(t1, t2, 0.5, 0.1, 1.5, 100)
    ens.init(th0, z0)
    ens.tune_reset(0.01)
    infos, xs = [], []
    for i in range(n):
        infos.append(ens.tune(1, i, tc, z=z[i:i + 1], info=True)[0])
        xs.append(ens.get_state()[0])
    info, xs = np.stack(infos), np.stack(xs)
    ens.tune_finish_phase2()
    eps, L, emax, mx, mx2 = ens.get_tuning(moments=True)
    final_state = ens.get_state()
    f64 = lambda t: o.logpost_value_and_grad(ospec, t, X.astype(np.float64), y)
    for c in range(C):
        ts = o.tune_init(cfg, d, np.float64)
        for i in range(n):
            assert info[i, c, 3] == 1.0
            ts = o.tune_update(cfg, ts._replace(step_size_max=np.float64(info[i, c, 2])),
                               xs[i, c].astype(np.float64), np.float64(info[i, c, 0]), True, i)
            assert abs(info[i, c, 1] - ts.step_size) <= 2e-5 * ts.step_size, (i, info[i, c, 1], ts.step_size)
        ts = o.tune_finish_phase2(cfg, ts)
        assert abs(eps[c] - ts.step_size) <= 2e-5 * ts.step_size
        assert rel(mx[c], ts.avg_x) <= 1e-5
        assert rel(mx2[c], ts.avg_x2) <= 1e-5
        assert abs(L[c] - ts.L) <= 1e-4 * ts.L
        # first step against the full fp64 oracle: energy change within the fp32 energy resolution
        st = o.mclmc_init(f64, th0[c].astype(np.float64), z0[c].astype(np.float64))
        st, ts0, inf0, ok = o.tune_step(f64, cfg, st, o.tune_init(cfg, d, np.float64), z[0, c].astype(np.float64), 0)
        assert abs(info[0, c, 0] - inf0.energy_change) <= 1e-5 * abs(st.logdensity)
        assert rel(xs[0, c], st.position) <= 1e-5
    # one 40-step launch == 40 one-step launches, bit for bit
    ens.init(th0, z0)
    ens.tune_reset(0.01)
    info2 = ens.tune(n, 0, tc, z=z, info=True)
    np.testing.assert_array_equal(info2, info)
    for a, b in zip(ens.get_state(), final_state):
        np.testing.assert_array_equal(a, b)
    ens.close()


# (handle_nans: the deterministic failure case is compared field by field with the oracle in
#  tests/test_gpu_parity_shapes.py::test_handle_nans_failure_matches_oracle)


def test_lppd_and_predict_match_oracle():
    name, C, S = 'airfoil_3x16', 3, 4
    ospec, ens, X, y, Xt, yt = make(name, C)
    ens.set_test(Xt, yt)
    rng = np.random.default_rng(2)
    thetas = (rng.standard_normal((S, C, ospec.n_params)) * 0.3).astype(np.float32)
    out = ens.predict(thetas[0], 'test')
    ref = np.stack([o.forward(ospec, thetas[0, c].astype(np.float64), Xt.astype(np.float64)) for c in range(C)])
    assert rel(out, ref) <= 1e-5
    for s in range(S):
        ens.lppd_accumulate(thetas[s])
    m, sst, cnt = ens.lppd_state()
    assert cnt == S
    from mile_b200 import lppd_from_state
    got = lppd_from_state(m, sst, C * S)
    lv = np.stack([[o.forward(ospec, thetas[s, c].astype(np.float64), Xt.astype(np.float64)) for s in range(S)]
                   for c in range(C)])
    want = o.lppd(o.pointwise_lppd(ospec, lv, yt.astype(np.float64)))
    assert abs(got - want) <= 1e-5 * abs(want)
    ens.close()


def test_fused_lppd_during_sampling_equals_posthoc():
    name, C = 'airfoil_2x16', 2
    ospec, ens, X, y, Xt, yt = make(name, C)
    ens.set_test(Xt, yt)
    th0 = o.synthetic_theta0(ospec, C)
    ens.init(th0, seed=3)
    samples, _ = ens.sample(40, 0.02, 20.0, n_thinning=10, seed=7, lppd=True)
    m, s, cnt = ens.lppd_state()
    assert cnt == 4 and samples.shape[0] == 4
    from mile_b200 import lppd_from_state
    got = lppd_from_state(m, s, C * 4)
    lv = np.stack([[o.forward(ospec, samples[k, c].astype(np.float64), Xt.astype(np.float64)) for k in range(4)]
                   for c in range(C)])
    want = o.lppd(o.pointwise_lppd(ospec, lv, yt.astype(np.float64)))
    assert abs(got - want) <= 1e-5 * abs(want)
    ens.close()


def test_philox_noise_statistics_gaussian_target():
    """Perf-mode noise (in-kernel Philox): a linear 'network' with a flat likelihood is a pure
    N(0,1) prior target -> sample mean ~ 0, variance ~ 1 (debug.ipynb cell 8 style check)."""
    from mile_b200 import Ensemble, FCNSpec
    spec = FCNSpec(3, (4, 2), 'identity', 'regr', n_batches=0.0)   # likelihood switched off
    ens = Ensemble(spec, 8)
    ens.set_data(np.zeros((4, 3), np.float32), np.zeros(4, np.float32))
    d = spec.n_params
    ens.init(np.random.default_rng(0).standard_normal((8, d)).astype(np.float32), seed=1)
    samples, _ = ens.sample(20000, 0.5, 3.0, n_thinning=5, seed=2)
    xs = samples[400:].reshape(-1, d)
    assert np.all(np.abs(xs.mean(0)) < 0.06)
    assert np.all(np.abs(xs.var(0) - 1) < 0.08)
    ens.close()


def test_caller_owned_output_buffers():
    """sample(out=) / get_state(out=) fill caller-owned (e.g. pinned) host arrays and return views of them; results are
    identical to the freshly allocated form."""
    import torch
    name, C, n, thin = 'airfoil_3x16', 3, 12, 4
    ospec, ens, X, y, _, _ = make(name, C)
    d = ospec.n_params
    th0 = o.synthetic_theta0(ospec, C)
    ens.init(th0, seed=3)
    s_ref, _ = ens.sample(n, 0.01, 20.0, n_thinning=thin, seed=9)
    st_ref = ens.get_state()
    ens.init(th0, seed=3)
    buf = torch.empty(s_ref.size + 7, dtype=torch.float32, pin_memory=True).numpy()
    s_out, _ = ens.sample(n, 0.01, 20.0, n_thinning=thin, seed=9, out=buf)
    assert np.shares_memory(s_out, buf) and s_out.shape == s_ref.shape
    np.testing.assert_array_equal(s_out, s_ref)
    outs = tuple(torch.empty(a.shape, dtype=torch.float32, pin_memory=True).numpy() for a in st_ref)
    st = ens.get_state(out=outs)
    for a, b, c_ in zip(st, st_ref, outs):
        assert a is c_
        np.testing.assert_array_equal(a, b)
    ens.close()


def test_errors_are_loud():
    from mile_b200 import Ensemble, FCNSpec
    from mile_b200.capi import MileError
    ens = Ensemble(FCNSpec(5, (16, 2)), 2)
    with pytest.raises(MileError):
        ens.value_and_grad(np.zeros((2, ens.d), np.float32))   # no data yet
    with pytest.raises(MileError):
        ens.set_option('no_such_option', 1)
    ens.close()
    gel = FCNSpec(5, (512, 512, 2), 'gelu')
    with pytest.raises(MileError):
        Ensemble(gel, 2)                                       # gelu has no wide-path implementation: loud failure


@pytest.mark.parametrize('widths,act,task', [((256, 256, 256, 256, 2), 'relu', 'regr'), ((96, 80, 3), 'tanh', 'class'),
                                             ((48, 48, 48, 2), 'relu', 'regr'),
                                             # fused output-layer kernel with a classification head: 128- and 256-wide last hidden layer
                                             ((128, 128, 7), 'sigmoid', 'class'), ((64, 256, 3), 'tanh', 'class'),
                                             # more than one 256-column block, K not a multiple of the 32-wide k-block
                                             ((272, 264, 2), 'sigmoid', 'regr')])
def test_wide_path_value_and_grad_and_step(widths, act, task):
    """Shapes that do not fit the shared-memory kernels (complexity ablation: 4x256, d = 201218) run on the
    HBM-resident layer-by-layer path (chain-batched GEMMs + integrator kernel): same parity bar."""
    from mile_b200 import Ensemble, FCNSpec
    F, N, C = 12, 700, 2
    ospec = o.ModelSpec(F, widths, act, task)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, F)).astype(np.float32)
    y = rng.standard_normal(N).astype(np.float32) if task == 'regr' else rng.integers(0, widths[-1], N).astype(np.int32)
    d = ospec.n_params
    th = (rng.standard_normal((C, d)) * (0.5 / np.sqrt(max(widths)))).astype(np.float32)
    ens = Ensemble(FCNSpec(F, widths, act, task), C)
    ens.set_data(X, y)
    lp, g = ens.value_and_grad(th)
    lp64, g64 = o.logpost_batch(ospec, th.astype(np.float64), X.astype(np.float64), y)
    for c in range(C):
        assert abs(lp[c] - lp64[c]) <= 1e-5 * abs(lp64[c])
        assert rel(g[c], g64[c]) <= 2e-5
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    z = rng.standard_normal((2, C, d)).astype(np.float32)
    ens.init(th, z0)
    samples, info = ens.sample(2, 0.01, float(np.sqrt(d)), z=z, info=True)
    thg, ug, lpg, gg = ens.get_state()
    f64 = lambda t: o.logpost_value_and_grad(ospec, t, X.astype(np.float64), y)
    for c in range(C):
        st = o.mclmc_init(f64, th[c].astype(np.float64), z0[c].astype(np.float64))
        for s in range(2):
            st, inf = o.mclmc_step(f64, st, 0.01, float(np.sqrt(d)), z[s, c].astype(np.float64))
        assert rel(thg[c], st.position) <= 1e-5
        assert rel(ug[c], st.momentum) <= 2e-5
        assert abs(lpg[c] - st.logdensity) <= 1e-5 * abs(st.logdensity)
        assert abs(info[1, c, 2] - inf.energy_change) <= 2e-5 * abs(st.logdensity)
    ens.close()
