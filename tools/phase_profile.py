"""Phase-level cycle breakdown of the persistent kernel (developer tool, not part of the product).
Builds a -DMILE_PROFILE variant of the library into tools/_prof/ and runs a short sampling launch."""
import ctypes, subprocess, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
out = ROOT / 'tools' / '_prof'
out.mkdir(exist_ok=True)
lib = out / 'libmile_b200.so'
src = ROOT / 'mile_b200' / 'csrc'
import os
if not (os.environ.get('MILE_PROF_NOBUILD') == '1' and lib.exists()):   # (build here, run on the GPU box with MILE_PROF_NOBUILD=1)
    subprocess.run(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo', '-DMILE_PROFILE',
                    '-shared', '-Xcompiler', '-fPIC', '-o', str(lib), str(src / 'mile_api.cu'), str(src / 'mile_microbench.cu')], check=True)
if len(sys.argv) > 1 and sys.argv[1] == 'build':
    sys.exit(0)
from mile_b200 import capi
capi.lib_path = lambda: lib
from mile_b200 import Ensemble, FCNSpec
from oracle import mile_oracle as o
NAMES = {0: 'x-tile/loop | mma: integrator warp + weight image', 1: 'forward | mma: tile loop (thread 0 = warp 0)', 2: 'loglik | mma: reduce-scatter + 2nd cluster.sync', 3: 'backward', 4: 'dW accumulate', 5: 'cross-chunk reduce',
         8: 'esh_update (B)', 9: 'position_update (A)', 10: 'grad_eval total (outer)', 11: 'cluster reduce (rest: block_sum)', 12: 'refresh', 13: 'publish partials | 1st cluster.sync', 14: 'wait + sum ranks'}
CASES = [('airfoil_3x16', 12, 200, {}), ('airfoil_3x16', 12, 200, {'cluster_size': 8}), ('airfoil_3x16', 12, 200, {'cluster_size': 4}),
         ('bikesharing_2x16', 10, 50, {}), ('protein_2x16', 10, 50, {}), ('airfoil_3x16', 1024, 20, {})]
if len(sys.argv) > 1 and sys.argv[1] == 'airfoil':
    CASES = CASES[:2]
if len(sys.argv) > 1 and sys.argv[1] == 'fast1':
    CASES = [(n, c, s, dict(o_, fast=1)) for n, c, s, o_ in CASES]
for name, C, steps, opts in CASES:
    ospec = o.make_spec(name)
    X, y, _, _ = o.synthetic_data(name)
    ens = Ensemble(FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task), C, **opts)
    ens.set_data(X, y)
    ens.init(o.synthetic_theta0(ospec, C), seed=1)
    prof = (ctypes.c_ulonglong * 32)()
    ens.lib.mile_debug_read_profile(prof, 1)
    ens.sample(steps, 0.02, float(np.sqrt(ospec.n_params)), keep=False, seed=3)
    ens.lib.mile_debug_read_profile(prof, 1)
    tot = sum(prof[i] for i in (0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 12, 13, 14))
    print(f'--- {name} C={C} opts={opts} fast={ens.get_option("fast")} sync={ens.get_option("sync_mode")} G={ens.get_option("cluster_size")} TR={ens.get_option("tile_rows")} '
          f'resident={ens.get_option("resident")}: {tot / steps:.0f} cycles/step')
    for i, n in NAMES.items():
        print(f'  {n:28s} {prof[i] / steps:10.0f} cycles/step  {100 * prof[i] / tot:5.1f}%')
    print('  barrier wait per warp (cycles/step):', ' '.join(f'{prof[16 + w] / steps:.0f}' for w in range(16)))
    ens.close()
