#!/usr/bin/env python
"""Dump golden vectors from the REAL reference stack (jax 0.4.28 + blackjax 1.2.2 [+ flax 0.8.5 and the MILE sources])
so that the oracle under oracle/ -- and through it the CUDA path -- can be pinned to the reference's own arithmetic.

    python tools/dump_blackjax_golden.py [--reference /path/to/MILE] [--out tests/golden/blackjax_vectors.npz]

This image has no jax/blackjax, so the file this script writes does not exist in the tree yet and
parity stays "unpinned" until someone runs it once on a host with the pinned stack (poetry.lock of the reference) or a
driver-provided baseline/_ref install.  tests/test_blackjax_golden.py activates as soon as the file exists and checks
the numpy oracle, the C oracle and (on a GPU) the CUDA path against it, reporting which refresh placement
(`refresh_mode` 0 = single post-step refresh, 1 = `with_isokinetic_maruyama` half-step refreshes) the installed
blackjax implements.

What is dumped (all float32, the reference's dtype; flat vectors in jax.flatten_util.ravel_pytree order):
  * the synthetic airfoil-shaped problem (X [1052,5], y, theta0) from mile_b200/synthetic.py (numpy only);
  * `blackjax.mcmc.mclmc.init` -> (position, momentum, logdensity, logdensity_grad) and the normal draw behind the
    momentum (reference call site: src/training/warmup.py:539-541);
  * three `build_kernel(logdensity_fn, integrator=isokinetic_mclachlan, sqrt_diag_cov)` steps with their MCLMCInfo and
    the normal draws for BOTH readings of the kernel: z_post[s] = normal(key_s) and z_mar[s] = normal(split(key_s))
    (reference call site: src/training/warmup.py:286-291, src/training/sampling.py:133-150);
  * with --reference: the log-density is the reference's own ProbabilisticModel.log_unnormalized_posterior on its own
    FCN (src/training/probabilistic.py:115-138, src/models/tabular/fcn.py:11-28), and one
    `custom_mclmc_warmup(...).run(key, position, 200)` (src/training/warmup.py:486-568) with every normal draw it
    consumed; without it a pure-jax restatement of the log-density is used and the warmup block is skipped;
  * `blackjax.diagnostics.effective_sample_size` on a fixed [1, 500, 7] array;
  * the NUTS branch: `blackjax.mcmc.nuts.init` + four `nuts.build_kernel()` transitions (max_num_doublings 5, a fixed
    diagonal inverse mass matrix) with their NUTSInfo, and the random draws of every transition REPLAYED from its key in the
    layout of oracle/nuts_oracle.py (momentum normals; direction / merge / per-leapfrog uniforms): `jax.random.bernoulli(key,
    p)` is `uniform(key) < p`, and blackjax 1.2.2 derives the keys as split(rng_key, 2) -> (momentum, integrator);
    fold_in(integrator, expansion) -> split 3 -> (direction, trajectory, proposal); fold_in(trajectory, leaf).  If the
    installed blackjax derives them differently, tests/test_blackjax_golden.py reports the mismatch of the replay itself
    (tree sizes) before any arithmetic is judged.  Plus one `window_adaptation.base` trace (init / update / final over a
    fixed sequence of positions and acceptance rates) and `build_schedule(1000)`.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default=None, help='checkout of zhiyuan-yang/MILE (adds its root to sys.path)')
    ap.add_argument('--out', default=str(ROOT / 'tests' / 'golden' / 'blackjax_vectors.npz'))
    ap.add_argument('--warmup-steps', type=int, default=200)
    args = ap.parse_args()

    import jax
    import jax.numpy as jnp
    import blackjax
    from blackjax.mcmc import mclmc
    from blackjax.mcmc.integrators import isokinetic_mclachlan
    from blackjax.diagnostics import effective_sample_size
    from jax.flatten_util import ravel_pytree

    from mile_b200 import synthetic as syn
    key_name = 'airfoil_3x16'
    spec = syn.workload_spec(key_name)
    X, y, Xt, yt = syn.synthetic_data(key_name, seed=1234)
    d = spec.n_params
    theta0 = syn.synthetic_theta0(d, 1, seed0=1000, scale=0.3)[0]
    widths = list(spec.widths)
    F = spec.n_features

    # parameter pytree in the reference's naming (src/flax_building_blocks/basic.py:54): fcn/layer{i}/{bias,kernel};
    # ravel_pytree flattens dicts in sorted-key order: per layer bias then kernel
    def unravel_np(th):
        out, off, fan_in = {}, 0, F
        for l, w in enumerate(widths):
            b = th[off:off + w]; off += w
            k = th[off:off + fan_in * w].reshape(fan_in, w); off += fan_in * w
            out[f'layer{l}'] = {'bias': jnp.asarray(b), 'kernel': jnp.asarray(k)}
            fan_in = w
        return {'fcn': out}

    position = unravel_np(theta0)
    flat0, unravel = ravel_pytree(position)
    assert np.array_equal(np.asarray(flat0), theta0), 'ravel_pytree order differs from the documented leaf order'
    Xj, yj = jnp.asarray(X), jnp.asarray(y)

    source = 'pure-jax restatement of probabilistic.py:92-138'
    logdensity_fn = None
    warm = None
    if args.reference:
        sys.path.insert(0, str(Path(args.reference).resolve()))
        try:
            from functools import partial
            from src.config.data import Task
            from src.config.models.base import Activation
            from src.config.models.fcn import FCNConfig
            from src.models.tabular.fcn import FCN
            from src.training.priors import PriorDist
            from src.training.probabilistic import ProbabilisticModel
            module = FCN(config=FCNConfig(hidden_structure=widths, activation=Activation.RELU))
            pm = ProbabilisticModel(module=module, params=position, prior=PriorDist.StandardNormal.get_prior(),
                                    task=Task.REGRESSION, n_batches=1)
            logdensity_fn = partial(pm.log_unnormalized_posterior, x=Xj, y=yj)
            source = 'reference ProbabilisticModel.log_unnormalized_posterior on reference FCN'
        except Exception as e:   # noqa: BLE001
            print(f'[dump] could not build the reference log-density ({e!r}); falling back to the pure-jax restatement')
    if logdensity_fn is None:
        import jax.scipy.stats as stats

        def logdensity_fn(p):
            h = Xj
            n = len(widths)
            for l in range(n):
                lay = p['fcn'][f'layer{l}']
                h = h @ lay['kernel'] + lay['bias']
                if l < n - 1:
                    h = jax.nn.relu(h)
            ll = jnp.nansum(stats.norm.logpdf(x=yj, loc=h[..., 0], scale=jnp.exp(h[..., 1]).clip(min=1e-6, max=1e6)))
            return jnp.sum(stats.norm.logpdf(ravel_pytree(p)[0], loc=0.0, scale=1.0)) + ll

    flat = lambda t: np.asarray(ravel_pytree(t)[0], np.float32)
    out = {'X': X, 'y': y, 'theta0': theta0, 'widths': np.asarray(widths), 'n_features': np.asarray(F),
           'logdensity_source': np.asarray(source), 'jax_version': np.asarray(jax.__version__),
           'blackjax_version': np.asarray(getattr(blackjax, '__version__', 'unknown'))}

    # ---- init + three kernel steps --------------------------------------------------------------------------
    root = jax.random.PRNGKey(0)
    k_init, k0, k1, k2 = jax.random.split(root, 4)
    state = mclmc.init(position=position, logdensity_fn=logdensity_fn, rng_key=k_init)
    out['init_z'] = np.asarray(jax.random.normal(k_init, (d,), jnp.float32))
    out['init_momentum'] = flat(state.momentum)
    out['init_logdensity'] = np.asarray(state.logdensity, np.float32)
    out['init_grad'] = flat(state.logdensity_grad)
    eps, L = 0.01, float(np.sqrt(d))
    kernel = mclmc.build_kernel(logdensity_fn=logdensity_fn, integrator=isokinetic_mclachlan,
                                sqrt_diag_cov=jnp.ones((d,)))
    pos, mom, lp, grad, info, z_post, z_mar = [], [], [], [], [], [], []
    for k in (k0, k1, k2):
        state, inf = kernel(rng_key=k, state=state, L=L, step_size=eps)
        pos.append(flat(state.position)); mom.append(flat(state.momentum)); lp.append(np.float32(state.logdensity))
        grad.append(flat(state.logdensity_grad))
        info.append([np.float32(inf.logdensity), np.float32(inf.kinetic_change), np.float32(inf.energy_change)])
        z_post.append(np.asarray(jax.random.normal(k, (d,), jnp.float32)))
        ka, kb = jax.random.split(k)
        z_mar.append(np.stack([np.asarray(jax.random.normal(ka, (d,), jnp.float32)),
                               np.asarray(jax.random.normal(kb, (d,), jnp.float32))]))
    out.update(step_size=np.float32(eps), L=np.float32(L), step_position=np.stack(pos), step_momentum=np.stack(mom),
               step_logdensity=np.asarray(lp), step_grad=np.stack(grad), step_info=np.asarray(info, np.float32),
               z_post=np.stack(z_post), z_mar=np.stack(z_mar))

    # ---- one full warmup through the reference's own tuner ------------------------------------------------------
    if args.reference:
        try:
            from src.training.warmup import custom_mclmc_warmup
            W = args.warmup_steps
            t1, t2, t3 = int(W * 0.8), int(W * 0.1), int(W * 0.1)
            cfg = dict(diagonal_preconditioning=False, desired_energy_var_start=0.5, desired_energy_var_end=0.1,
                       trust_in_estimate=1.5, num_effective_samples=100, step_size_init=0.01)
            wkey = jax.random.PRNGKey(7)
            res = custom_mclmc_warmup(logdensity_fn, **cfg).run(wkey, position, W)
            st, par = res.state, res.parameters
            part1, part2 = jax.random.split(wkey, 2)                              # warmup.py:210
            keys1 = jax.random.split(part1, t1 + t2 + 1)[:-1]                      # warmup.py:367-373
            keys3 = jax.random.split(part2, t3)                                    # warmup.py:424
            zs_post = [np.asarray(jax.random.normal(k, (d,), jnp.float32)) for k in list(keys1) + list(keys3)]
            zs_mar = []
            for k in list(keys1) + list(keys3):
                ka, kb = jax.random.split(k)
                zs_mar.append(np.stack([np.asarray(jax.random.normal(ka, (d,), jnp.float32)),
                                        np.asarray(jax.random.normal(kb, (d,), jnp.float32))]))
            out.update(warm_steps=np.asarray(W), warm_cfg=np.asarray([cfg['desired_energy_var_start'],
                       cfg['desired_energy_var_end'], cfg['trust_in_estimate'], cfg['num_effective_samples'],
                       cfg['step_size_init']], np.float32),
                       warm_init_z=np.asarray(jax.random.normal(wkey, (d,), jnp.float32)),   # same key as the tuner: warmup.py:540,552
                       warm_z_post=np.stack(zs_post), warm_z_mar=np.stack(zs_mar),
                       warm_position=flat(st.position), warm_momentum=flat(st.momentum),
                       warm_logdensity=np.float32(st.logdensity), warm_step_size=np.float32(par.step_size),
                       warm_L=np.float32(par.L))
            warm = True
        except Exception as e:   # noqa: BLE001
            print(f'[dump] reference warmup not dumped: {e!r}')

    # ---- ESS on a fixed array -------------------------------------------------------------------------------------
    rng = np.random.default_rng(5)
    x = np.empty((1, 500, 7), np.float32)
    x[0, 0] = rng.standard_normal(7)
    rho = np.linspace(0.1, 0.95, 7).astype(np.float32)
    for i in range(1, 500):                                  # AR(1) columns with different autocorrelation
        x[0, i] = rho * x[0, i - 1] + np.sqrt(1 - rho ** 2) * rng.standard_normal(7).astype(np.float32)
    out['ess_x'] = x
    out['ess'] = np.asarray(effective_sample_size(jnp.asarray(x)), np.float32)

    # ---- NUTS branch (sampling.py:70-81,107-210; warmup.py:27-152) ----------------------------------------------
    try:
        from blackjax.mcmc import nuts as bj_nuts
        from blackjax.adaptation.window_adaptation import base as wa_base, build_schedule
        D = 5
        imm = np.exp(0.3 * np.random.default_rng(9).standard_normal(d)).astype(np.float32)
        n_eps = np.float32(2e-3)
        nkernel = bj_nuts.build_kernel()
        nstate = bj_nuts.init(position, logdensity_fn)
        out['nuts_init_logdensity'] = np.float32(nstate.logdensity)
        out['nuts_init_grad'] = flat(nstate.logdensity_grad)
        n_len = 2 * D + 2 ** D
        uniform = lambda k: np.float32(jax.random.uniform(k, (), jnp.float32))
        zs, unis, npos, nlp, ngrad, ninfo = [], [], [], [], [], []
        for k in jax.random.split(jax.random.PRNGKey(11), 4):
            nstate, inf = nkernel(k, nstate, logdensity_fn, n_eps, jnp.asarray(imm), max_num_doublings=D)
            km, ki = jax.random.split(k, 2)                                   # nuts.py kernel
            zs.append(np.asarray(jax.random.normal(km, (d,), jnp.float32)))   # util.generate_gaussian_noise over the raveled position
            u = np.zeros(n_len, np.float32)
            for j in range(D):                                                # trajectory.dynamic_multiplicative_expansion
                dk, tk, pk = jax.random.split(jax.random.fold_in(ki, j), 3)
                u[j], u[D + j] = uniform(dk), uniform(pk)
                for i in range(2 ** j):                                       # trajectory.dynamic_progressive_integration
                    u[2 * D + 2 ** j - 1 + i] = uniform(jax.random.fold_in(tk, i))
            unis.append(u)
            npos.append(flat(nstate.position)); nlp.append(np.float32(nstate.logdensity)); ngrad.append(flat(nstate.logdensity_grad))
            ninfo.append([np.float32(inf.num_integration_steps), np.float32(inf.acceptance_rate),
                          np.float32(inf.num_trajectory_expansions), np.float32(inf.is_divergent), np.float32(inf.energy),
                          np.float32(inf.is_turning)])
        out.update(nuts_max_doublings=np.asarray(D), nuts_step_size=n_eps, nuts_imm=imm, nuts_z=np.stack(zs),
                   nuts_uni=np.stack(unis), nuts_position=np.stack(npos), nuts_logdensity=np.asarray(nlp),
                   nuts_grad=np.stack(ngrad), nuts_info=np.asarray(ninfo, np.float32))
        # window adaptation arithmetic on a fixed input sequence (no sampler involved)
        a_init, a_update, a_final = wa_base(True, target_acceptance_rate=0.8)
        sched = build_schedule(60)
        rng = np.random.default_rng(13)
        xs = rng.standard_normal((60, 6)).astype(np.float32) * np.linspace(0.5, 2.0, 6).astype(np.float32)
        accs = rng.uniform(0.3, 1.0, 60).astype(np.float32)
        ast = a_init(jnp.zeros(6, jnp.float32), 0.1)
        trace = []
        for i in range(60):
            ast = a_update(ast, (sched[i][0], sched[i][1]), jnp.asarray(xs[i]), accs[i])
            trace.append(np.concatenate([[np.float32(ast.step_size)], np.asarray(ast.inverse_mass_matrix, np.float32)]))
        fe, fm = a_final(ast)
        out.update(wa_positions=xs, wa_acceptance=accs, wa_trace=np.stack(trace), wa_final_step_size=np.float32(fe),
                   wa_final_imm=np.asarray(fm, np.float32), wa_schedule_1000=np.asarray(build_schedule(1000)).astype(np.int32))
    except Exception as e:   # noqa: BLE001
        print(f'[dump] NUTS block not dumped: {e!r}')

    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(args.out, **out)
    print(f'[dump] wrote {args.out}: log-density = {source}; warmup block = {bool(warm)}; '
          f'jax {jax.__version__}, blackjax {getattr(blackjax, "__version__", "?")}')


if __name__ == '__main__':
    main()
