import sys; sys.path.insert(0, '.')
import numpy as np
from mile_b200 import Ensemble, FCNSpec
from oracle import mile_oracle as o
name = sys.argv[1] if len(sys.argv) > 1 else 'airfoil_3x16'
ospec = o.make_spec(name)
X, y, _, _ = o.synthetic_data(name)
ens = Ensemble(FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task), 2, cluster_size=1)
ens.set_data(X, y)
th = o.synthetic_theta0(ospec, 2)
lp, g = ens.value_and_grad(th); print('eval ok', lp)
lp, g = ens.value_and_grad(th); print('eval2 ok', lp)
ens.init(th, seed=1); print('init ok')
s, i = ens.sample(1, 0.01, 20.0, seed=3, info=True); print('sample1 ok', i)
s, i = ens.sample(3, 0.01, 20.0, seed=3, info=True); print('sample3 ok', i[:, 0])
