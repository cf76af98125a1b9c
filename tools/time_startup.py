"""Developer timing of the process start-up of the CUDA path (what a whole run pays before its first step)."""
import sys, time
t0 = time.perf_counter()
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np                                         # noqa: E402
t1 = time.perf_counter()
from mile_b200 import Ensemble, FCNSpec, capi              # noqa: E402
from mile_b200 import synthetic as syn                     # noqa: E402
t2 = time.perf_counter()
capi.load()
t3 = time.perf_counter()
spec = syn.workload_spec('airfoil_3x16')
X, y, Xt, yt = syn.synthetic_data('airfoil_3x16', seed=1234)
fs = FCNSpec(spec.n_features, spec.widths, spec.activation, spec.task)
t4 = time.perf_counter()
ens = Ensemble(fs, 12)
t5 = time.perf_counter()
ens.set_data(X, y)
t6 = time.perf_counter()
th0 = (0.3 * np.random.default_rng(0).standard_normal((12, fs.n_params))).astype(np.float32)
ens.init(th0, seed=1)
t7 = time.perf_counter()
ens.tune_reset(0.01)
cfg = ens.tune_cfg(80, 10, 0.5, 0.1, 1.5, 100)
ens.tune(90, 0, cfg, seed=2)
t8 = time.perf_counter()
ens.tune(90, 90, cfg, seed=2)
t9 = time.perf_counter()
print(f'numpy import {t1 - t0:.3f} s | package import {t2 - t1:.3f} | dlopen {t3 - t2:.3f} | synthetic data {t4 - t3:.3f} | '
      f'mile_create (CUDA context) {t5 - t4:.3f} | set_data {t6 - t5:.3f} | first launch (init) {t7 - t6:.3f} | '
      f'first tuning launch {t8 - t7:.3f} | second tuning launch {t9 - t8:.3f}')
