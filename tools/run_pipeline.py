"""The whole MILE flow on the CUDA path, end to end, on synthetic data of a named shape -- what `python train.py -c
experiments/illustrative_example_readme/mclmc.yaml` does in the reference (train.py:23-69 -> BDETrainer.train_bde,
src/training/trainer.py:292-328): warm-start training of the deep-ensemble members, sampling with the YAML's sampler block
(read from the committed fixture tests/golden/sampler_configs.json), and the report's numbers (evaluate_bde, chain-wise
R-hat / ESS) from the files the run wrote.

    python tools/run_pipeline.py [illustrative_example_readme/mclmc.yaml | illustrative_example_readme/nuts.yaml | ...] [--fast]
"""
import argparse, functools, json, logging, sys, tempfile, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mile_b200 import FCN, FCNSpec, ProbabilisticModel, SamplerConfig, evaluate_bde, inference_loop      # noqa: E402
from mile_b200 import metrics as M                                                                    # noqa: E402
from mile_b200 import synthetic as syn                                                                # noqa: E402
from mile_b200.training import train_warmstart                                                        # noqa: E402
from mile_b200.utils import load_params_batch, load_samples_from_dir                                  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('config', nargs='?', default='illustrative_example_readme/mclmc.yaml')
ap.add_argument('--fast', action='store_true', help='a tenth of the warm-up / sampling steps')
args = ap.parse_args()
logging.basicConfig(level=logging.WARNING)
cfgs = json.loads((ROOT / 'tests' / 'golden' / 'sampler_configs.json').read_text())
block = cfgs[args.config]
sd = dict(block['sampler'])
if args.fast:
    sd['warmup_steps'] = max(100, sd['warmup_steps'] // 10)
    sd['n_samples'] = max(100, sd['n_samples'] // 10)
cfg = SamplerConfig.from_dict(sd)
widths = tuple(block['model']['hidden_structure'])
workload = 'airfoil_3x16' if len(widths) == 4 else 'airfoil_2x16'
spec_s = syn.workload_spec(workload)
X, y, Xt, yt = syn.synthetic_data(workload, seed=1234)
n_valid = len(X) // 8                                              # 70 / 10 / 20 split of the reference's data block
Xv, yv, Xtr, ytr = X[:n_valid], y[:n_valid], X[n_valid:], y[n_valid:]
C = int(cfg.n_chains)
module = FCN(widths, block['model']['activation'])
fs = FCNSpec(spec_s.n_features, widths, block['model']['activation'], 'regr')
out = {'config': args.config, 'sampler': cfg.name.value, 'chains': C, 'rows': {'train': len(Xtr), 'valid': len(Xv), 'test': len(Xt)}}
with tempfile.TemporaryDirectory() as tmp:
    exp = Path(tmp) / 'exp'
    t0 = time.perf_counter()
    _, tm = train_warmstart(fs, (Xtr, ytr), (Xv, yv), (Xt, yt), exp, range(C), optimizer={'name': 'adamw', 'learning_rate': 5e-3},
                            max_epochs=20 if args.fast else 100, batch_size=32, patience=10, seed=0)
    t1 = time.perf_counter()
    out['warmstart'] = {'seconds': t1 - t0, 'epochs': int(tm['epochs']), 'test_nll_mean': float(np.nanmean(tm['test'][:, 0])),
                        'test_rmse_mean': float(np.nanmean(tm['test'][:, 1]))}
    init = load_params_batch(sorted((exp / 'warmstart').glob('params_*.npz')))          # what start_sampling reads (trainer.py:560-575)
    pm = ProbabilisticModel(module, init, cfg.prior, 'regr')
    pm.attach_test_split(Xt, yt)
    log_post = functools.partial(pm.log_unnormalized_posterior, x=Xtr, y=ytr)
    info = inference_loop(log_post, cfg, 0, init, np.arange(C), exp / 'samples')
    t2 = time.perf_counter()
    out['sampling'] = {'seconds': t2 - t1, 'fused_lppd': float(info['lppd'])}
    if 'is_divergent' in info:
        out['sampling'].update(divergent_per_chain=[round(float(v), 3) for v in np.mean(info['is_divergent'], axis=1)],
                               mean_tree=float(np.mean(info['num_integration_steps'])),
                               acceptance_per_chain=[round(float(v), 3) for v in np.mean(info['acceptance_rate'], axis=1)])
    samples = load_samples_from_dir(exp / 'samples')                                      # what the report notebook reads
    lv, em = evaluate_bde(samples, module, Xt, yt, 'regr', verbose=False)
    # per-chain view: RMSE of the posterior-mean prediction (no predictive noise) and the largest predicted log-sigma
    mean_rmse = [float(np.sqrt(np.mean((yt - lv[c, :, :, 0].mean(axis=0)) ** 2))) for c in range(C)]
    max_logsig = [float(lv[c, :, :, 1].max()) for c in range(C)]
    flat = pm.spec.ravel(samples)                                                         # [C, S, d]
    rhat = np.asarray(M.split_chain_r_hat(flat, 4))
    ess = np.asarray(M.effective_sample_size(flat))
    t3 = time.perf_counter()
    out['report'] = {'seconds': t3 - t2, 'lppd': float(em['lppd']), 'rmse': float(em['rmse']),
                     'rmse_of_mean_prediction_per_chain': [round(v, 4) for v in mean_rmse],
                     'max_predicted_log_sigma_per_chain': [round(v, 2) for v in max_logsig], 'samples_per_chain': int(flat.shape[1]),
                     'split_rhat_median': float(np.median(rhat)), 'split_rhat_max': float(np.max(rhat)),
                     'ess_median': float(np.median(ess)), 'ess_min': float(np.min(ess))}
    out['total_seconds'] = t3 - t0
print(json.dumps(out))
