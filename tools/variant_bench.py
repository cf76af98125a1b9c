"""Developer tool: chain-steps/s of a library variant (path to a .so built with experimental -D flags)."""
import sys, time; sys.path.insert(0, '.')
from pathlib import Path
import numpy as np
from mile_b200 import capi
if len(sys.argv) > 1:
    capi.lib_path = lambda: Path(sys.argv[1]).resolve()
from mile_b200 import Ensemble, FCNSpec
from oracle import mile_oracle as o
import torch
for name, C, steps in (('bikesharing_2x16', 10, 300), ('airfoil_3x16', 1024, 100), ('airfoil_3x16', 12, 1000), ('protein_2x16', 10, 100)):
    ospec = o.make_spec(name)
    X, y, _, _ = o.synthetic_data(name)
    ens = Ensemble(FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task), C)
    ens.set_data(X, y)
    ens.init(o.synthetic_theta0(ospec, C), seed=1)
    ens.sample(20, 0.02, float(np.sqrt(ospec.n_params)), keep=False, seed=3)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ens.sample(steps, 0.02, float(np.sqrt(ospec.n_params)), keep=False, seed=4)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'{name} C={C}: {C * steps / dt:12.0f} chain-steps/s', flush=True)
    ens.close()
