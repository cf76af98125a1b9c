"""Pin of the oracle (and of the CUDA path) to the REAL reference arithmetic -- active only when
tests/golden/blackjax_vectors.npz exists.  That file is written by tools/dump_blackjax_golden.py on a host that has
jax 0.4.28 + blackjax 1.2.2 (the reference's pinned stack, pyproject.toml:9-25); this image has neither, so here the
whole module skips and parity stays "unpinned" (DESIGN.md section 6).  What runs once the file is present:

  * numpy oracle (fp32 and fp64 twin) vs the dumped `mclmc.init` state and three `build_kernel` steps, under BOTH
    readings of the kernel (refresh_mode 0: single post-step refresh; 1: `with_isokinetic_maruyama`); exactly one must
    match, and the test says which;
  * C oracle vs the same vectors;
  * the tuner (`custom_mclmc_warmup.run`, src/training/warmup.py:486-568) when the dump holds the warmup block;
  * the ESS restatement vs `blackjax.diagnostics.effective_sample_size`;
  * `-m gpu`: the CUDA path vs the dumped states with the matching refresh_mode.
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import mile_oracle as o

PATH = Path(__file__).resolve().parent / 'golden' / 'blackjax_vectors.npz'
pytestmark = pytest.mark.skipif(not PATH.exists(), reason='tests/golden/blackjax_vectors.npz not generated '
                                '(needs jax + blackjax: tools/dump_blackjax_golden.py)')


def rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(b), 1e-30)


def _load():
    G = np.load(PATH)
    spec = o.ModelSpec(int(G['n_features']), tuple(int(w) for w in G['widths']), 'relu', 'regr')
    return G, spec


def _replay(G, spec, dtype, mode):
    """Replays init + the three steps with the dumped normal draws; returns the worst relative deviation."""
    X, y = G['X'].astype(dtype), G['y'].astype(dtype)
    f = lambda t: o.logpost_value_and_grad(spec, t, X, y)
    st = o.mclmc_init(f, G['theta0'].astype(dtype), G['init_z'].astype(dtype))
    worst = max(rel(st.momentum, G['init_momentum']), rel(st.logdensity_grad, G['init_grad']),
                abs(st.logdensity - G['init_logdensity']) / abs(G['init_logdensity']))
    eps, L = float(G['step_size']), float(G['L'])
    for s in range(G['step_position'].shape[0]):
        z = G['z_post'][s].astype(dtype) if mode == 0 else G['z_mar'][s].astype(dtype)
        st, info = o.mclmc_step(f, st, eps, L, z, refresh='post' if mode == 0 else 'maruyama')
        scale = abs(float(G['step_logdensity'][s]))
        worst = max(worst, rel(st.position, G['step_position'][s]), rel(st.momentum, G['step_momentum'][s]),
                    abs(st.logdensity - G['step_logdensity'][s]) / scale,
                    abs(info.kinetic_change - G['step_info'][s, 1]) / scale,
                    abs(info.energy_change - G['step_info'][s, 2]) / scale)
    return worst


def matching_refresh_mode(G, spec):
    dev = {m: _replay(G, spec, np.float64, m) for m in (0, 1)}
    ok = [m for m in (0, 1) if dev[m] <= 1e-5]
    assert len(ok) == 1, f'exactly one refresh placement must reproduce blackjax {G["blackjax_version"]}: deviations {dev}'
    return ok[0], dev


def test_numpy_oracle_matches_blackjax_and_reports_refresh_mode():
    G, spec = _load()
    mode, dev = matching_refresh_mode(G, spec)
    print(f'\n[golden] blackjax {G["blackjax_version"]} / jax {G["jax_version"]} ({G["logdensity_source"]}): '
          f'refresh_mode {mode} matches (max rel dev {dev[mode]:.2e}; the other reading deviates by {dev[1 - mode]:.2e})')
    assert _replay(G, spec, np.float32, mode) <= 5e-5      # the fp32 restatement carries its own rounding


def test_c_oracle_matches_blackjax():
    from oracle import c_oracle
    if not c_oracle.available():
        pytest.skip('C oracle not built')
    G, spec = _load()
    mode, _ = matching_refresh_mode(G, spec)
    if mode != 0:
        pytest.skip('the C oracle implements the post-step refresh only')
    lp, g = c_oracle.logpost_batch(spec, G['theta0'][None], G['X'], G['y'])
    assert abs(lp[0] - G['init_logdensity']) <= 1e-5 * abs(G['init_logdensity'])
    assert rel(g[0], G['init_grad']) <= 1e-5


def test_tuner_matches_reference_warmup():
    G, spec = _load()
    if 'warm_step_size' not in G.files:
        pytest.skip('the dump was made without --reference: no warmup block')
    mode, _ = matching_refresh_mode(G, spec)
    W = int(G['warm_steps'])
    ev0, ev1, trust, neff, eps0 = (float(v) for v in G['warm_cfg'])
    cfg = o.TuneConfig(int(W * 0.8), int(W * 0.1), int(W * 0.1), ev0, ev1, trust, neff, eps0)
    X, y = G['X'].astype(np.float64), G['y'].astype(np.float64)
    f = lambda t: o.logpost_value_and_grad(spec, t, X, y)
    z = (G['warm_z_post'] if mode == 0 else G['warm_z_mar']).astype(np.float64)
    st, eps, L, _ = o.run_warmup(f, cfg, G['theta0'].astype(np.float64), G['warm_init_z'].astype(np.float64), z,
                                 refresh='post' if mode == 0 else 'maruyama')
    # 200 chaotic steps in fp32 vs an fp64 replay: the tuned scalars agree to a few per cent, not to 1e-5
    assert abs(eps - G['warm_step_size']) <= 0.05 * G['warm_step_size']
    assert abs(L - G['warm_L']) <= 0.10 * G['warm_L']


def test_ess_matches_blackjax():
    G, _ = _load()
    ess = o.effective_sample_size(G['ess_x'].astype(np.float64))
    assert rel(ess, G['ess']) <= 1e-4


@pytest.mark.gpu
def test_cuda_path_matches_blackjax():
    from mile_b200 import Ensemble, FCNSpec
    G, spec = _load()
    mode, _ = matching_refresh_mode(G, spec)
    ens = Ensemble(FCNSpec(spec.n_features, spec.widths, 'relu', 'regr'), 1, refresh_mode=mode)
    ens.set_data(G['X'], G['y'])
    ens.init(G['theta0'][None], G['init_z'][None])
    th, u, lp, g = ens.get_state()
    assert rel(u[0], G['init_momentum']) <= 1e-5 and rel(g[0], G['init_grad']) <= 1e-5
    assert abs(lp[0] - G['init_logdensity']) <= 1e-5 * abs(G['init_logdensity'])
    n = G['step_position'].shape[0]
    z = G['z_post'][:, None, :] if mode == 0 else G['z_mar'][:, :, None, :]
    for s in range(n):
        _, info = ens.sample(1, float(G['step_size']), float(G['L']), z=np.ascontiguousarray(z[s:s + 1]), step_base=s,
                             keep=False, info=True)
        th, u, lp, g = ens.get_state()
        scale = abs(float(G['step_logdensity'][s]))
        assert rel(th[0], G['step_position'][s]) <= 1e-5
        assert rel(u[0], G['step_momentum'][s]) <= 2e-5
        assert abs(lp[0] - G['step_logdensity'][s]) <= 1e-5 * scale
        assert abs(info[0, 0, 2] - G['step_info'][s, 2]) <= 1e-5 * scale
    ens.close()


# ---- NUTS branch (present when the dump holds the nuts_* / wa_* arrays) -------------------------------------------------
def _has_nuts():
    return PATH.exists() and 'nuts_position' in np.load(PATH).files


@pytest.mark.skipif(not _has_nuts(), reason='the golden file has no NUTS block')
def test_nuts_oracle_matches_blackjax_transitions():
    """oracle/nuts_oracle.py (fp64 and fp32) replays the dumped blackjax.nuts transitions with the replayed draws: first the
    replay itself is judged (tree sizes / expansions / flags must agree: they depend on the direction and acceptance draws),
    then the arithmetic (position, log-density, gradient, acceptance rate, energy)."""
    from oracle import nuts_oracle as no
    G, spec = _load()
    D, eps = int(G['nuts_max_doublings']), float(G['nuts_step_size'])
    for dt, tol in ((np.float64, 2e-5), (np.float32, 5e-5)):
        X, y = G['X'].astype(dt), G['y'].astype(dt)
        f = lambda t: o.logpost_value_and_grad(spec, t, X, y)
        th = G['theta0'].astype(dt)
        lp, g = f(th)
        assert abs(lp - G['nuts_init_logdensity']) <= 1e-5 * abs(G['nuts_init_logdensity'])
        assert rel(g, G['nuts_init_grad']) <= 1e-5
        for k in range(G['nuts_position'].shape[0]):
            th, lp, g, info = no.nuts_step(f, th, dt(lp), g, eps, G['nuts_imm'].astype(dt), G['nuts_z'][k], G['nuts_uni'][k], D)
            want = G['nuts_info'][k]
            assert (info.num_integration_steps, info.num_trajectory_expansions, info.is_divergent, info.is_turning) == \
                (int(want[0]), int(want[2]), bool(want[3]), bool(want[5])), \
                f'transition {k}: the replayed draws (or the tree logic) do not reproduce blackjax {G["blackjax_version"]}: {info} vs {want}'
            assert rel(th, G['nuts_position'][k]) <= tol, (k, dt)
            assert abs(lp - G['nuts_logdensity'][k]) <= tol * abs(G['nuts_logdensity'][k])
            assert rel(g, G['nuts_grad'][k]) <= 10 * tol
            assert abs(info.acceptance_rate - want[1]) <= 1e-4 and abs(info.energy - want[4]) <= tol * abs(want[4]) + 1e-3
            # continue from blackjax's own state so that one flipped branch cannot cascade
            th, lp, g = G['nuts_position'][k].astype(dt), dt(G['nuts_logdensity'][k]), G['nuts_grad'][k].astype(dt)


@pytest.mark.skipif(not _has_nuts(), reason='the golden file has no NUTS block')
def test_window_adaptation_matches_blackjax():
    from oracle import nuts_oracle as no
    from mile_b200.nuts import build_schedule
    G, _ = _load()
    want = [(int(a), bool(b)) for a, b in G['wa_schedule_1000']]
    assert no.build_schedule(1000) == want and build_schedule(1000) == want
    sched = no.build_schedule(60)
    st = no.adapt_init(6, 0.1, np.float64)
    for i in range(60):
        st = no.adapt_step(st, sched[i], G['wa_positions'][i].astype(np.float64), float(G['wa_acceptance'][i]))
        assert abs(st.step_size - G['wa_trace'][i, 0]) <= 1e-4 * G['wa_trace'][i, 0], i
        np.testing.assert_allclose(st.imm, G['wa_trace'][i, 1:], rtol=1e-4, atol=1e-7)
    e, m = no.adapt_final(st)
    assert abs(e - G['wa_final_step_size']) <= 1e-4 * G['wa_final_step_size']
    np.testing.assert_allclose(m, G['wa_final_imm'], rtol=1e-4, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.skipif(not _has_nuts(), reason='the golden file has no NUTS block')
def test_cuda_nuts_matches_blackjax_transitions():
    from mile_b200 import Ensemble, FCNSpec
    G, spec = _load()
    D = int(G['nuts_max_doublings'])
    ens = Ensemble(FCNSpec(spec.n_features, spec.widths, spec.activation, spec.task), 1)
    ens.set_data(G['X'], G['y'])
    ens.nuts_init(G['theta0'][None], max_num_doublings=D)
    ens.set_nuts_params(float(G['nuts_step_size']), G['nuts_imm'][None])
    for k in range(G['nuts_position'].shape[0]):
        pos, info = ens.nuts_sample(1, step_base=k, z=G['nuts_z'][k][None, None], uni=G['nuts_uni'][k][None, None], info=True)
        want = G['nuts_info'][k]
        assert (int(info[0, 0, 0]), int(info[0, 0, 2]), bool(info[0, 0, 3]), bool(info[0, 0, 5])) == \
            (int(want[0]), int(want[2]), bool(want[3]), bool(want[5])), (k, info[0, 0], want)
        assert rel(pos[0, 0], G['nuts_position'][k]) <= 5e-5
        th, _, lp, g = ens.get_state()
        assert abs(lp[0] - G['nuts_logdensity'][k]) <= 5e-5 * abs(G['nuts_logdensity'][k])
        assert abs(info[0, 0, 1] - want[1]) <= 1e-3
        ens.set_state(theta=G['nuts_position'][k][None], lp=G['nuts_logdensity'][k:k + 1], grad=G['nuts_grad'][k][None])
    ens.close()
