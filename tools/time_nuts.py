"""Developer timing of the NUTS branch: gradient evaluations per second inside mile_nuts_kernel (window adaptation, then
sampling transitions with in-kernel Philox noise) beside the MCLMC step loop on the same workload.
Usage: python tools/time_nuts.py [workload] [chains] [warmup_transitions] [sampling_transitions]"""
import json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mile_b200 import Ensemble, FCNSpec                    # noqa: E402
from mile_b200 import synthetic as syn                     # noqa: E402
from mile_b200.nuts import build_schedule                  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'airfoil_3x16'
C = int(sys.argv[2]) if len(sys.argv) > 2 else 12
n_warm = int(sys.argv[3]) if len(sys.argv) > 3 else 400
n_samp = int(sys.argv[4]) if len(sys.argv) > 4 else 200
spec = syn.workload_spec(name)
X, y, Xt, yt = syn.synthetic_data(name, seed=1234)
fs = FCNSpec(spec.n_features, spec.widths, spec.activation, spec.task)
ens = Ensemble(fs, C)
ens.set_data(X, y)
rng = np.random.default_rng(0)
th0 = (0.3 * rng.standard_normal((C, fs.n_params))).astype(np.float32)
ens.nuts_init(th0, max_num_doublings=10, initial_step_size=1.0)
sched = build_schedule(n_warm)
t0 = time.perf_counter()
done, evals_w = 0, 0
while done < n_warm:
    n = min(100, n_warm - done)
    info = ens.nuts_warmup(n, sched[done:done + n], step_base=done, seed=1, info=True)
    evals_w += int(info[..., 0].sum())
    done += n
ens.nuts_finish_warmup()
t_w = time.perf_counter() - t0
eps, imm = ens.nuts_params()
t0 = time.perf_counter()
samples, info = ens.nuts_sample(n_samp, seed=2, info=True)
t_s = time.perf_counter() - t0
evals_s = int(info[..., 0].sum())
# MCLMC step loop on the same ensemble shape for scale (2 gradient evaluations per step)
ens2 = Ensemble(fs, C)
ens2.set_data(X, y)
ens2.init(th0, seed=3)
ens2.sample(2000, 0.05, 5.0, seed=4, keep=False)
t0 = time.perf_counter()
ens2.sample(20000, 0.05, 5.0, seed=5, keep=False)
t_m = time.perf_counter() - t0
print(json.dumps({'workload': name, 'chains': C, 'nuts_warmup': {'transitions': n_warm, 'grad_evals': evals_w, 'seconds': t_w,
                  'grad_evals_per_s': evals_w / t_w}, 'nuts_sampling': {'transitions': n_samp, 'grad_evals': evals_s,
                  'seconds': t_s, 'grad_evals_per_s': evals_s / t_s, 'mean_tree_size': float(info[..., 0].mean()),
                  'mean_acceptance': float(info[..., 1].mean()), 'divergent': int(info[..., 3].sum())},
                  'step_size': eps.tolist(), 'mclmc_grad_evals_per_s': 2 * 20000 * C / t_m}))
