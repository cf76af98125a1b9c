"""A phase-3 call for an ncu capture of the ESS kernels: airfoil 3x16, 12 chains, 5000 captured steps -> 8088 series."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mile_b200 import Ensemble, FCNSpec                    # noqa: E402
from mile_b200 import synthetic as syn                     # noqa: E402

spec = syn.workload_spec('airfoil_3x16')
X, y, _, _ = syn.synthetic_data('airfoil_3x16', seed=1234)
fs = FCNSpec(spec.n_features, spec.widths, spec.activation, spec.task)
ens = Ensemble(fs, 12)
ens.set_data(X, y)
ens.init((0.3 * np.random.default_rng(0).standard_normal((12, fs.n_params))).astype(np.float32), seed=1)
ens.sample(3000, 0.05, 5.0, seed=2, keep=False)
t = time.perf_counter()
ess = ens.phase3_ess(5000, 0.05, 5.0, seed=3)
print(f'phase 3 (5000 steps + ESS of {ess.size} series): {time.perf_counter() - t:.3f} s; median ESS {np.median(ess):.1f}')
ens.close()
