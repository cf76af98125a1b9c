"""A short NUTS launch for an ncu capture of mile_nuts_kernel: airfoil 3x16, 12 chains, max_num_doublings 5, a few
transitions with the in-kernel Philox noise (after a short window adaptation so that the trees are typical)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mile_b200 import Ensemble, FCNSpec                    # noqa: E402
from mile_b200 import synthetic as syn                     # noqa: E402
from mile_b200.nuts import build_schedule                  # noqa: E402

spec = syn.workload_spec('airfoil_3x16')
X, y, _, _ = syn.synthetic_data('airfoil_3x16', seed=1234)
fs = FCNSpec(spec.n_features, spec.widths, spec.activation, spec.task)
ens = Ensemble(fs, 12)
ens.set_data(X, y)
th0 = (0.3 * np.random.default_rng(0).standard_normal((12, fs.n_params))).astype(np.float32)
ens.nuts_init(th0, max_num_doublings=5, initial_step_size=0.01)
ens.nuts_warmup(40, build_schedule(40), seed=1)
ens.nuts_finish_warmup()
_, info = ens.nuts_sample(8, seed=2, info=True)
print('tree sizes', info[..., 0].mean(), 'launches', ens.launches)
ens.close()
