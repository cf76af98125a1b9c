"""Developer timing: MCLMC steps/s of the tuning loop (phases 1+2, 83 % of a reference run's gradient evaluations) against
the sampling loop, same ensemble.  Usage: python tools/time_tune_vs_sample.py [workload] [chains]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
from mile_b200 import Ensemble
from mile_b200 import synthetic as syn
key = sys.argv[1] if len(sys.argv) > 1 else 'airfoil_3x16'
C = int(sys.argv[2]) if len(sys.argv) > 2 else 12
spec = syn.workload_spec(key)
X, y, Xt, yt = syn.synthetic_data(key, seed=1234)
d = spec.n_params
ens = Ensemble(spec, C, device=0)
ens.set_data(X, y)
ens.init(syn.synthetic_theta0(d, C, seed0=1000, scale=0.3), seed=17)
ens.tune_reset(0.01)
n = 4000
tc = ens.tune_cfg(3 * n, n, 0.5, 0.1, 1.5, 100)
ens.tune(500, 0, tc, seed=1)
for phase, base in (('phase 1', 500), ('phase 2', 3 * n)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ens.tune(n, base, tc, seed=2)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'{key} C={C} tune {phase}: {C * n / dt:,.0f} chain-steps/s')
ens.tune_finish_phase2()
eps, L, _ = ens.get_tuning()
torch.cuda.synchronize(); t0 = time.perf_counter()
ens.sample(n, eps, L, n_thinning=10, seed=3)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f'{key} C={C} sample (host buffers, thinning 10): {C * n / dt:,.0f} chain-steps/s; eps {eps.mean():.4f} L {L.mean():.2f}')
ens.close()
