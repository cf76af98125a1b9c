"""Developer diagnostic: per-layer-block relative error of the wide-path gradient (4x256 full shape) against the fp64
oracle, per chain, for a chosen GEMM core and with the fused output-layer kernel on / off.
Usage: python tools/wide_grad_blocks.py [tensor] [head_fused]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import mile_oracle as o  # noqa: E402
from mile_b200 import Ensemble, FCNSpec  # noqa: E402

tensor = int(sys.argv[1]) if len(sys.argv) > 1 else 0
head = int(sys.argv[2]) if len(sys.argv) > 2 else 1
C = 8
ospec = o.make_spec('wide_4x256')
X, y, _, _ = o.synthetic_data('wide_4x256')
th = o.synthetic_theta0(ospec, C, scale=0.05)
lp64, g64 = o.logpost_batch(ospec, th.astype(np.float64), X.astype(np.float64), y)
ens = Ensemble(FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task), C, tensor=tensor)
ens.set_option('head_fused', head)
ens.set_data(X, y)
lp, g = ens.value_and_grad(th)
rel = lambda a, b: np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30)
dims = [ospec.n_features] + list(ospec.widths)
# flat order: per layer (sorted keys) bias then kernel
names = sorted(f'layer{l}' for l in range(len(ospec.widths)))
off = 0
blocks = []
for nm in names:
    l = int(nm[5:])
    blocks.append((f'{nm}.bias', off, off + dims[l + 1])); off += dims[l + 1]
    blocks.append((f'{nm}.kernel', off, off + dims[l] * dims[l + 1])); off += dims[l] * dims[l + 1]
assert off == ospec.n_params
print('tensor', tensor, 'head_fused', head)
for c in range(C):
    print(f'chain {c}: total {rel(g[c], g64[c]):.2e} lp {abs(lp[c]-lp64[c])/abs(lp64[c]):.1e} | ' +
          ' '.join(f'{nm}={rel(g[c, a:b], g64[c, a:b]):.1e}' for nm, a, b in blocks))
