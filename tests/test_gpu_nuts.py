"""NUTS branch on the GPU (csrc/mile_nuts.cuh through the C ABI) against the numpy oracle (oracle/nuts_oracle.py) with the
randomness supplied by the host: single transitions from identical states, the window adaptation, Philox-driven
statistics on a known target, and the reference-facing seams (KERNELS['nuts'], inference_loop with sampler name 'nuts')."""
import functools
import pickle

import numpy as np
import pytest

from oracle import mile_oracle as o
from oracle import nuts_oracle as no

pytestmark = pytest.mark.gpu

D = 5


def problem(name='airfoil_2x16', n_train=300):
    if name == 'prior_only':     # likelihood switched off: the target is the N(0, 1) prior, where U-turns end most trajectories
        ospec = o.ModelSpec(3, (4, 2), 'identity', 'regr', n_batches=0.0)
        X, y = np.zeros((4, 3), np.float32), np.zeros(4, np.float32)
        return ospec, X, y, None, None, (lambda th: o.logpost_value_and_grad(ospec, th, X.astype(np.float64), y))
    ospec = o.make_spec(name)
    X, y, Xt, yt = o.synthetic_data(name, n_train=n_train, n_test=64)
    X64 = X.astype(np.float64)
    return ospec, X, y, Xt, yt, (lambda th: o.logpost_value_and_grad(ospec, th, X64, y))


def transition_cases(ospec, lg, C, T, seed=0, burn=40, scales=(1.0, 2.5, 6.0)):
    """Typical states: every chain is burnt in by the fp64 oracle's own window adaptation; the T test transitions of chain c
    use its adapted metric and step sizes scaled by 1, 2.5 and 6 (full-depth trees, U-turns, divergences)."""
    rng = np.random.default_rng(seed)
    d = ospec.n_params
    th0 = o.synthetic_theta0(ospec, C)
    th = np.empty((C, d)); lp = np.empty(C); g = np.empty((C, d)); eps = np.empty(C); imm = np.empty((C, d))
    for c in range(C):
        th[c], lp[c], g[c], e, imm[c], _ = no.run_window_adaptation(
            lg, th0[c].astype(np.float64), rng.standard_normal((burn, d)), rng.random((burn, no.uni_len(D))), max_num_doublings=D)
        eps[c] = e * scales[c % 3]
    z = rng.standard_normal((T, C, d)).astype(np.float32)
    uni = rng.random((T, C, no.uni_len(D))).astype(np.float32)
    return th, lp, g, eps, imm, z, uni


@pytest.mark.parametrize('name,G,fast', [('airfoil_2x16', 1, 2), ('airfoil_2x16', 4, 2), ('airfoil_3x16', 8, 2),
                                         ('airfoil_2x16', 2, 0), ('airfoil_2x16', 12, 2), ('covertype_ref', 4, 0),
                                         ('prior_only', 1, 0), ('prior_only', 2, 0)])
def test_single_transitions_match_oracle(name, G, fast):
    """Each transition starts from the fp64 oracle's state (position, logdensity, gradient rounded to fp32), so only the
    arithmetic of ONE transition is compared: tree size, expansions, divergence / U-turn flags exactly; proposal, its
    logdensity and energy, acceptance rate to fp32 accuracy.  A data-dependent branch (U-turn sign, acceptance draw) that
    sits within rounding of its threshold may legitimately resolve differently in fp32: at most 1 in 10 transitions may
    differ from the fp64 oracle in their discrete outcome, and every such case must agree with the fp32 oracle instead."""
    from mile_b200 import Ensemble, FCNSpec
    C, T = 3, 8
    ospec, X, y, _, _, lg = problem(name, 300 if name != 'covertype_ref' else 400)
    th, lp, g, eps, imm, z, uni = transition_cases(ospec, lg, C, T, scales=(1.0, 0.35, 2.5) if name == 'prior_only' else (1.0, 2.5, 6.0))
    X32 = X
    lg32 = lambda t: o.logpost_value_and_grad(ospec, t, X32, y)
    ens = Ensemble(FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task, n_batches=float(ospec.n_batches)), C,
                   cluster_size=G, fast=fast)
    ens.set_data(X, y)
    ens.nuts_init(th.astype(np.float32), max_num_doublings=D)
    ens.set_nuts_params(eps, imm)
    kinds, mismatched = set(), 0
    eps_base = eps.copy()
    jit = 0.5 + np.random.default_rng(7).random((T, C)) if name == 'prior_only' else np.ones((T, C))
    for k in range(T):
        eps = eps_base * jit[k]              # (prior_only: a different step size per transition varies how trajectories end)
        ens.set_nuts_params(step_size=eps)
        ens.set_state(theta=th.astype(np.float32), lp=lp.astype(np.float32), grad=g.astype(np.float32))
        pos, info = ens.nuts_sample(1, step_base=k, z=z[k:k + 1], uni=uni[k:k + 1], info=True)
        th_g, _, lp_g, g_g = ens.get_state()
        np.testing.assert_array_equal(pos[0], th_g)
        new = []
        for c in range(C):
            t64 = th[c].astype(np.float32).astype(np.float64)
            want = no.nuts_step(lg, t64, np.float64(np.float32(lp[c])), g[c].astype(np.float32).astype(np.float64), np.float32(eps[c]),
                                imm[c].astype(np.float32).astype(np.float64), z[k, c], uni[k, c], D)
            wi = want[3]
            got = (int(info[0, c, 0]), int(info[0, c, 2]), bool(info[0, c, 3]), bool(info[0, c, 5]))
            if got != (wi.num_integration_steps, wi.num_trajectory_expansions, wi.is_divergent, wi.is_turning) or \
                    np.linalg.norm(th_g[c] - want[0]) > 1e-3 * np.linalg.norm(want[0]):
                mismatched += 1
                want = no.nuts_step(lg32, th[c].astype(np.float32), np.float32(lp[c]), g[c].astype(np.float32), np.float32(eps[c]),
                                    imm[c].astype(np.float32), z[k, c], uni[k, c], D)
                wi = want[3]
                assert got == (wi.num_integration_steps, wi.num_trajectory_expansions, wi.is_divergent, wi.is_turning), (k, c, got, wi)
            kinds.add((wi.num_integration_steps == 2 ** D - 1, wi.is_divergent, wi.is_turning,
                       wi.is_turning and not wi.is_divergent and (wi.num_integration_steps & (wi.num_integration_steps + 1)) != 0))
            scale = np.linalg.norm(want[0])
            assert np.linalg.norm(th_g[c] - want[0]) <= 2e-5 * scale, (k, c, wi)
            assert abs(lp_g[c] - want[1]) <= 2e-5 * abs(want[1]), (k, c, wi)
            assert np.linalg.norm(g_g[c] - want[2]) <= 1e-4 * np.linalg.norm(want[2]), (k, c, wi)
            assert abs(info[0, c, 1] - wi.acceptance_rate) <= 1e-3, (k, c, wi)
            assert abs(info[0, c, 4] - wi.energy) <= 2e-5 * abs(wi.energy) + 1e-3, (k, c, wi)
            assert abs(info[0, c, 6] - lp_g[c]) == 0 and abs(info[0, c, 7] - np.float32(eps[c])) == 0
            new.append(want)
        # follow the fp64 oracle to the next state
        for c in range(C):
            wn = no.nuts_step(lg, th[c].astype(np.float32).astype(np.float64), np.float64(np.float32(lp[c])),
                              g[c].astype(np.float32).astype(np.float64), np.float32(eps[c]),
                              imm[c].astype(np.float32).astype(np.float64), z[k, c], uni[k, c], D)
            th[c], lp[c], g[c] = wn[0], wn[1], wn[2]
    assert mismatched <= (C * T) // 10, mismatched
    assert len(kinds) >= 2, kinds        # the cases exercised more than one way of ending a trajectory
    if name == 'prior_only':
        assert any(k[2] and not k[1] for k in kinds), kinds      # ... among them the U-turn of the whole trajectory
        assert any(k[3] for k in kinds), kinds                   # ... and of a sub-trajectory (tree size not 2^j - 1)
    ens.close()


def test_multi_transition_launch_equals_single_launches():
    """Chunking: n transitions in one launch == n launches of one transition (bit-exact), with thinning."""
    from mile_b200 import Ensemble, FCNSpec
    C, T = 3, 6
    ospec, X, y, _, _, lg = problem()
    th, lp, g, eps, imm, z, uni = transition_cases(ospec, lg, C, T, seed=1, burn=25)
    outs = []
    for chunks, opts in (([T], {}), ([1] * T, {}), ([4, 2], {}), ([T], {'nuts_smem': 0}), ([T], {'nuts_push': 0}),
                         ([T], {'cluster_size': 12})):
        # (nuts_smem 0: end states / checkpoints in the L2-resident scratch instead of shared memory; nuts_push 0: pull form of the
        #  cluster exchange; 12 CTAs per chain: the flagged-word exchange with a different row split, so only the first five are
        #  compared bit for bit)
        ens = Ensemble(FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task), C, **opts)
        ens.set_data(X, y)
        ens.nuts_init(th.astype(np.float32), max_num_doublings=D)
        ens.set_nuts_params(eps, imm)
        done, pos, infos = 0, [], []
        for n in chunks:
            p, i = ens.nuts_sample(n, step_base=done, n_thinning=2, z=z[done:done + n], uni=uni[done:done + n], info=True)
            pos.append(p); infos.append(i); done += n
        outs.append((np.concatenate(pos), np.concatenate(infos), ens.get_state()[0]))
        ens.close()
    assert outs[0][0].shape == (T // 2, C, ospec.n_params)
    for other in outs[1:5]:
        for a, b in zip(outs[0], other):
            np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(outs[5][1][..., [0, 2, 3, 5]], outs[0][1][..., [0, 2, 3, 5]])      # same trees
    assert np.linalg.norm(outs[5][0] - outs[0][0]) <= 1e-4 * np.linalg.norm(outs[0][0])


def test_window_adaptation_matches_oracle():
    """custom_window_adaptation's loop (warmup.py:84-101,133-142): after every transition the dual-averaging step size and,
    in the slow windows, the Welford metric.  The adaptation arithmetic is checked in isolation: the oracle's adapt_step is
    fed the GPU's own (position, acceptance_rate) sequence and must reproduce the GPU's step sizes and metric."""
    from mile_b200 import Ensemble, FCNSpec
    from mile_b200.nuts import build_schedule
    C, n = 2, 40
    ospec, X, y, _, _, lg = problem()
    rng = np.random.default_rng(5)
    d = ospec.n_params
    th0 = o.synthetic_theta0(ospec, C)
    z = rng.standard_normal((n, C, d)).astype(np.float32)
    uni = rng.random((n, C, no.uni_len(D))).astype(np.float32)
    sched = build_schedule(n)
    assert any(e for _, e in sched) and any(s == 0 for s, _ in sched)
    ens = Ensemble(FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task), C)
    ens.set_data(X, y)
    ens.nuts_init(th0, max_num_doublings=D, initial_step_size=0.01)
    states = [no.adapt_init(d, 0.01, np.float64) for _ in range(C)]
    done = 0
    for chunk in (13, 1, 26):
        infos = []
        for k in range(chunk):      # one transition per launch so that every position is visible
            infos.append(ens.nuts_warmup(1, sched[done + k:done + k + 1], step_base=done + k, z=z[done + k:done + k + 1],
                                         uni=uni[done + k:done + k + 1], info=True)[0])
            pos = ens.get_state()[0]
            eps_g, imm_g = ens.nuts_params()
            for c in range(C):
                assert abs(infos[-1][c, 7] - states[c].step_size) <= 1e-4 * states[c].step_size     # step size USED
                states[c] = no.adapt_step(states[c], sched[done + k], pos[c].astype(np.float64), float(infos[-1][c, 1]))
                assert abs(eps_g[c] - states[c].step_size) <= 1e-4 * states[c].step_size, (done + k, c)
                np.testing.assert_allclose(imm_g[c], states[c].imm, rtol=2e-4, atol=1e-9)
        done += chunk
    ens.nuts_finish_warmup()
    eps_g, imm_g = ens.nuts_params()
    for c in range(C):
        e, m = no.adapt_final(states[c])
        assert abs(eps_g[c] - e) <= 1e-4 * e
        np.testing.assert_allclose(imm_g[c], m, rtol=2e-4, atol=1e-9)
    # the same warm-up in ONE launch gives the same parameters bit for bit
    ens2 = Ensemble(FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task), C)
    ens2.set_data(X, y)
    ens2.nuts_init(th0, max_num_doublings=D, initial_step_size=0.01)
    ens2.nuts_warmup(n, sched, z=z, uni=uni)
    ens2.nuts_finish_warmup()
    e2, m2 = ens2.nuts_params()
    np.testing.assert_array_equal(e2, eps_g)
    np.testing.assert_array_equal(m2, imm_g)
    np.testing.assert_array_equal(ens2.get_state()[0], ens.get_state()[0])
    ens.close(); ens2.close()


def test_philox_driven_nuts_samples_the_prior_exactly():
    """In-kernel Philox draws: with the likelihood switched off the target is the N(0, 1) prior -- after the window adaptation
    the draws have mean 0 / variance 1, the acceptance rate sits at the 0.8 target and nothing diverges."""
    from mile_b200 import Ensemble, FCNSpec
    from mile_b200.nuts import run_window_adaptation
    spec = FCNSpec(3, (4, 2), 'identity', 'regr', n_batches=0.0)
    C, d = 8, spec.n_params
    ens = Ensemble(spec, C)
    ens.set_data(np.zeros((4, 3), np.float32), np.zeros(4, np.float32))
    th0 = np.random.default_rng(0).standard_normal((C, d)).astype(np.float32)
    eps, imm = run_window_adaptation(ens, th0, 3, 600, max_num_doublings=6)
    assert np.all((eps > 0.2) & (eps < 2.5)), eps
    assert np.all((imm > 0.3) & (imm < 3.0))
    samples, info = ens.nuts_sample(3000, seed=9, info=True)
    assert samples.shape == (3000, C, d)
    assert not np.any(info[..., 3] > 0)
    assert abs(info[..., 1].mean() - 0.8) < 0.08
    xs = samples.reshape(-1, d)
    assert np.all(np.abs(xs.mean(0)) < 0.05), xs.mean(0)
    assert np.all(np.abs(xs.var(0) - 1) < 0.08), xs.var(0)
    # chains are independent streams: no two chains share a draw
    assert len({samples[0, c].tobytes() for c in range(C)}) == C
    ens.close()


def test_nuts_seams_kernel_registry_and_inference_loop(tmp_path):
    """KERNELS['nuts'](logdensity_fn, step_size, inverse_mass_matrix).init/.step and inference_loop with sampler name 'nuts'
    (sampling.py:70-81,200-215): samples/{id}/sample_{n}.npz, no warmup_params.txt, info.pkl with the NUTSInfo fields."""
    from mile_b200 import FCN, KERNELS, PriorDist, ProbabilisticModel, SamplerConfig, inference_loop
    from mile_b200.utils import load_samples_from_dir
    ospec = o.make_spec('airfoil_2x16')
    X, y, Xt, yt = o.synthetic_data('airfoil_2x16', n_train=400, n_test=100)
    module = FCN(ospec.widths, ospec.activation)
    pm = ProbabilisticModel(module, module.init(np.random.default_rng(0), ospec.n_features), PriorDist.StandardNormal.get_prior(), 'regr')
    log_post = functools.partial(pm.log_unnormalized_posterior, x=X, y=y)
    pos = module.init(np.random.default_rng(1), ospec.n_features, scale=0.5)
    sampler = KERNELS['nuts'](log_post, step_size=1e-3, inverse_mass_matrix=np.ones(ospec.n_params, np.float32), max_num_doublings=4)
    st = sampler.init(pos)
    want, gwant = o.logpost_value_and_grad(ospec, o.ravel_tree(ospec, pos).astype(np.float64), X.astype(np.float64), y)
    assert abs(st.logdensity - want) <= 1e-5 * abs(want)
    st2, info = sampler.step(4, st)
    assert 1 <= int(info.num_integration_steps) <= 15 and 0 <= float(info.acceptance_rate) <= 1
    assert info.is_divergent in (True, False) and info._fields[6:] == ('num_trajectory_expansions', 'num_integration_steps', 'acceptance_rate')
    lp2, _ = o.logpost_value_and_grad(ospec, o.ravel_tree(ospec, st2.position).astype(np.float64), X.astype(np.float64), y)
    assert abs(st2.logdensity - lp2) <= 1e-5 * abs(lp2)
    # the whole loop
    pm.attach_test_split(Xt, yt)
    C = 3
    rng = np.random.default_rng(3)
    ps = [module.init(rng, ospec.n_features, scale=0.5) for _ in range(C)]
    posb = {'fcn': {k: {kk: np.stack([p['fcn'][k][kk] for p in ps]) for kk in v} for k, v in ps[0]['fcn'].items()}}
    cfg = SamplerConfig.from_dict({'name': 'nuts', 'warmup_steps': 120, 'n_chains': C, 'n_samples': 60, 'n_thinning': 1})
    exp = tmp_path / 'exp'
    info = inference_loop(log_post, cfg, 42, posb, np.array([0, 1, 2]), exp / 'samples', exp / 'warmup')
    assert not (exp / 'warmup_params.txt').exists()
    # keep_warmup (trainer.py:324 -> saving_path_warmup; warmup.py:102-109): the position BEFORE warm-up step n of every chain
    warm = load_samples_from_dir(exp / 'warmup')
    assert warm['fcn']['layer0']['kernel'].shape == (C, 120, ospec.n_features, 16)
    np.testing.assert_array_equal(warm['fcn']['layer1']['kernel'][:, 0], posb['fcn']['layer1']['kernel'])
    # (the first transitions run at the initial step size 1.0 and are rejected: compare the end of the warm-up with its start)
    assert np.any(warm['fcn']['layer1']['kernel'][:, -1] != warm['fcn']['layer1']['kernel'][:, 0])
    with open(exp / 'samples' / 'info.pkl', 'rb') as f:
        pk = pickle.load(f)
    assert set(pk) == {'num_integration_steps', 'acceptance_rate', 'num_trajectory_expansions', 'is_divergent', 'energy', 'is_turning'}
    assert all(v.shape == (C, 60) for v in pk.values())
    assert pk['num_integration_steps'].min() >= 1 and pk['num_integration_steps'].max() <= 1023 + 512
    samples = load_samples_from_dir(exp / 'samples')
    assert samples['fcn']['layer0']['kernel'].shape == (C, 60, ospec.n_features, 16)
    flat = pm.spec.ravel(samples)
    lv = np.stack([[o.forward(ospec, flat[c, s].astype(np.float64), Xt.astype(np.float64)) for s in range(60)] for c in range(C)])
    wl = o.lppd(o.pointwise_lppd(ospec, lv, yt.astype(np.float64)))
    assert abs(info['lppd'] - wl) <= 1e-4 * abs(wl)
    # the sampler moved and fits: the posterior-mean prediction beats the initial one
    assert np.isfinite(wl) and len({flat[0, s].tobytes() for s in range(60)}) > 20


@pytest.mark.parametrize('name,widths', [('airfoil_3x16', None), ('airfoil_2x16', (16, 16, 16, 16, 16, 2))])
def test_partition_nuts_matches_reduced_oracle(name, widths, tmp_path):
    """NUTS with the frozen-parameter mask (experiments/replicate_uci/partition_NUTS.yaml: partition_sampling + sampler nuts;
    src/training/partition_sampling.py:70-81,226-270): the transitions equal the oracle's on the REDUCED vector (sampled
    parameters only; reduced log-density = prior of the sampled layers + likelihood of the merged network, trainer.py:651-659),
    the frozen parameters keep their values bit for bit."""
    from mile_b200 import Ensemble, FCNSpec
    from tests.test_gpu_partition import reduced_logpost
    ospec = o.make_spec(name)
    if widths is not None:      # a deeper net runs on the generic evaluator (the shipped partition config has 9 layers)
        ospec = o.ModelSpec(ospec.n_features, widths, 'relu', 'regr')
    X, y, _, _ = o.synthetic_data(name, n_train=300)
    C, T = 2, 5
    spec = FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task)
    frozen = spec.hidden_layer_mask()
    active = np.flatnonzero(~frozen)
    d = ospec.n_params
    rng = np.random.default_rng(31)
    th0 = o.synthetic_theta0(ospec, C)
    imm = np.exp(0.5 * rng.standard_normal((C, d))).astype(np.float32)
    eps = np.array([2e-3, 6e-3], np.float32)
    z = rng.standard_normal((T, C, d)).astype(np.float32)
    uni = rng.random((T, C, no.uni_len(D))).astype(np.float32)
    ens = Ensemble(spec, C)
    ens.set_data(X, y)
    ens.set_frozen_mask(frozen)
    ens.nuts_init(th0, max_num_doublings=D)
    ens.set_nuts_params(eps, imm)
    th_i, _, lp_i, g_i = ens.get_state()
    pos, info = ens.nuts_sample(T, z=z, uni=uni, info=True)
    th_g, _, lp_g, g_g = ens.get_state()
    for c in range(C):
        f64 = reduced_logpost(ospec, X, y, th0[c].astype(np.float64), active)
        lp, g = f64(th0[c, active].astype(np.float64))
        assert abs(lp_i[c] - lp) <= 1e-5 * abs(lp) and np.all(g_i[c, frozen] == 0)
        assert np.linalg.norm(g_i[c, active] - g) <= 1e-5 * np.linalg.norm(g)
        th = th0[c, active].astype(np.float64)
        for k in range(T):
            th, lp, g, wi = no.nuts_step(f64, th, lp, g, eps[c], imm[c, active].astype(np.float64), z[k, c, active], uni[k, c], D)
            assert (int(info[k, c, 0]), int(info[k, c, 2]), bool(info[k, c, 3]), bool(info[k, c, 5])) == \
                (wi.num_integration_steps, wi.num_trajectory_expansions, wi.is_divergent, wi.is_turning), (k, c, wi)
            assert np.linalg.norm(pos[k, c, active] - th) <= 5e-5 * np.linalg.norm(th), (k, c)
            np.testing.assert_array_equal(pos[k, c, frozen], th0[c, frozen])
        assert abs(lp_g[c] - lp) <= 5e-5 * abs(lp)
        assert np.all(g_g[c, frozen] == 0)
    ens.close()
    # the reference-facing seam: partition_inference_loop with sampler name 'nuts'
    if widths is None:
        from mile_b200 import FCN, PriorDist, ProbabilisticModel, SamplerConfig, partition_inference_loop
        from mile_b200.utils import load_samples_from_dir
        module = FCN(ospec.widths, ospec.activation)
        pm = ProbabilisticModel(module, module.init(np.random.default_rng(0), ospec.n_features), PriorDist.StandardNormal.get_prior(), 'regr')
        ps = [module.init(rng, ospec.n_features, scale=0.5) for _ in range(C)]
        posb = {'fcn': {k: {kk: np.stack([p['fcn'][k][kk] for p in ps]) for kk in v} for k, v in ps[0]['fcn'].items()}}
        cfg = SamplerConfig.from_dict({'name': 'nuts', 'warmup_steps': 40, 'n_chains': C, 'n_samples': 20, 'n_thinning': 1,
                                       'partition_sampling': True})
        log_post = functools.partial(pm.log_unnormalized_posterior_partition, x=X, y=y)
        partition_inference_loop(log_post, cfg, 5, posb, np.arange(C), tmp_path / 'exp' / 'samples')
        s = load_samples_from_dir(tmp_path / 'exp' / 'samples')
        assert s['fcn']['layer1']['kernel'].shape == (C, 20, 16, 16)
        for c in range(C):      # hidden layers frozen at their initial values, outer layers moved
            np.testing.assert_array_equal(s['fcn']['layer1']['kernel'][c, -1], posb['fcn']['layer1']['kernel'][c])
            assert np.any(s['fcn']['layer0']['kernel'][c, -1] != posb['fcn']['layer0']['kernel'][c])


def test_nuts_and_mclmc_agree_on_the_posterior_predictive():
    """Two samplers, one posterior: from the same warmed-up states, NUTS (window adaptation + sampling, in-kernel Philox) and
    MCLMC (tuned step size / L) must give the same test-set LPPD and RMSE within the chain-to-chain spread (stated
    tolerances: |dLPPD| <= 3 SE + 0.03, |dRMSE| <= 0.05) -- the 'matching LPPD' of the north star, between the two kernels."""
    from mile_b200 import Ensemble, FCNSpec
    from mile_b200.nuts import build_schedule
    name, C = 'airfoil_2x16', 8
    ospec = o.make_spec(name)
    X, y, Xt, yt = o.synthetic_data(name, n_train=400, n_test=200)
    fs = FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task)

    def metrics(samples):      # [S, C, d]
        S = samples.shape[0]
        lv = np.stack([[o.forward(ospec, samples[s, c].astype(np.float64), Xt.astype(np.float64)) for s in range(S)]
                       for c in range(C)])
        pw = o.pointwise_lppd(ospec, lv, yt.astype(np.float64))
        per_chain = np.array([o.lppd(pw[c:c + 1]) for c in range(C)])
        return o.lppd(pw), per_chain, np.sqrt(np.mean((lv[..., 0].mean(axis=(0, 1)) - yt) ** 2))

    ens = Ensemble(fs, C)
    ens.set_data(X, y)
    ens.init(o.synthetic_theta0(ospec, C), seed=1)
    ens.tune_reset(0.01)
    ens.tune(1800, 0, ens.tune_cfg(1600, 200, 0.5, 0.1, 1.5, 100), seed=2)
    ens.tune_finish_phase2()
    eps, L, _ = ens.get_tuning()
    ens.sample(1500, eps, L, seed=3, keep=False)               # burn-in shared by both samplers
    start = ens.get_state()
    m_samples, _ = ens.sample(3000, eps, L, n_thinning=10, seed=4)
    ens.nuts_init(start[0], max_num_doublings=7, initial_step_size=float(np.mean(eps)))
    ens.nuts_warmup(400, build_schedule(400), seed=5)
    ens.nuts_finish_warmup()
    n_samples, info = ens.nuts_sample(300, seed=6, info=True)
    ens.close()
    # (divergent transitions -- |dE| > 1000 in the steep small-sigma regions of this posterior -- are rejected proposals; the
    #  fp64 oracle shows the same rate on this problem, so only sanity is asserted here)
    assert np.mean(info[..., 3]) < 0.5 and 0.4 < np.mean(info[..., 1]) < 0.97, (np.mean(info[..., 3]), np.mean(info[..., 1]))
    ml, mpc, mrmse = metrics(m_samples)
    nl, npc, nrmse = metrics(n_samples)
    se = np.sqrt(mpc.var(ddof=1) / C + npc.var(ddof=1) / C)
    assert abs(ml - nl) <= 3 * se + 0.03, (ml, nl, se)
    assert abs(mrmse - nrmse) <= 0.05, (mrmse, nrmse)
