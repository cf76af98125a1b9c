"""Developer timing of a WHOLE reference-sized run through the reference-facing seam: inference_loop with the sampler block
of experiments/illustrative_example_readme/mclmc.yaml (12 chains, 50 000 warmup steps, 10 000 samples, thinning 10) on
the synthetic airfoil problem, including the phase-3 ESS and the sample files.
Usage: python tools/full_run.py [npz|store] [mclmc|nuts]   (nuts: the sampler block of illustrative_example_readme/nuts.yaml:
100 warm-up transitions, 1000 samples, thinning 1)"""
import functools, logging, os, sys, tempfile, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
fmt = sys.argv[1] if len(sys.argv) > 1 else 'npz'
logging.basicConfig(level=logging.DEBUG, format='%(relativeCreated)8.0f ms %(name)s %(message)s')
os.environ['MILE_SAMPLE_FORMAT'] = fmt
sampler = sys.argv[2] if len(sys.argv) > 2 else 'mclmc'
from mile_b200 import FCN, PriorDist, ProbabilisticModel, inference_loop   # noqa: E402
from mile_b200.config import SamplerConfig                                  # noqa: E402
from mile_b200 import synthetic as syn                                      # noqa: E402

spec = syn.workload_spec('airfoil_3x16')
X, y, Xt, yt = syn.synthetic_data('airfoil_3x16', seed=1234)
module = FCN(spec.widths, spec.activation)
rng = np.random.default_rng(0)
pm = ProbabilisticModel(module, module.init(rng, spec.n_features), PriorDist.StandardNormal.get_prior(), 'regr')
pm.attach_test_split(Xt, yt)
C = 12
pos = [module.init(rng, spec.n_features, scale=0.3) for _ in range(C)]
tree = {'fcn': {k: {kk: np.stack([p['fcn'][k][kk] for p in pos]) for kk in v} for k, v in pos[0]['fcn'].items()}}
if sampler == 'nuts':
    cfg = SamplerConfig(name='nuts', warmup_steps=100, n_chains=12, n_samples=1000, n_thinning=1)
else:
    cfg = SamplerConfig(name='mclmc', warmup_steps=50000, n_chains=12, n_samples=10000, n_thinning=10,
                        desired_energy_var_start=0.5, desired_energy_var_end=0.1, trust_in_estimate=1.5,
                        num_effective_samples=100, step_size_init=0.01)
log_post = functools.partial(pm.log_unnormalized_posterior, x=X, y=y)
with tempfile.TemporaryDirectory() as tmp:
    t0 = time.perf_counter()
    info = inference_loop(log_post, cfg, 0, tree, np.arange(C), Path(tmp) / 'exp' / 'samples')
    dt = time.perf_counter() - t0
    nfiles = sum(1 for _ in (Path(tmp) / 'exp').rglob('*.npz'))
if sampler == 'nuts':
    ev = int(np.sum(info['num_integration_steps']))
    print(f'format {fmt}: NUTS inference_loop of 12 chains x (100 warm-up + 1000 sampling) transitions: {dt:.2f} s wall, {nfiles} npz '
          f'files; {ev:,} gradient evaluations in the sampling phase (mean tree {ev / (C * 1000):.0f}), acceptance '
          f"{float(np.mean(info['acceptance_rate'])):.2f}, divergent {float(np.mean(info['is_divergent'])):.2f}, LPPD {info.get('lppd', float('nan')):.4f}")
else:
    print(f'format {fmt}: inference_loop of 12 chains x (50000 warmup + 10000 sampling) steps: {dt:.2f} s wall, {nfiles} npz files; '
          f'{C * 60000 / dt:,.0f} chain-steps/s over the whole run')
