"""Writes tests/golden/oracle_vectors.npz: a small frozen trajectory of the fp64 oracle (NOT of the reference -- the
reference cannot run in this image).  It is a regression pin: the oracle and the CUDA path must both
keep reproducing these numbers, so a silent change to either is caught even when the two drift together.

    python tests/golden/make_oracle_vectors.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import mile_oracle as o   # noqa: E402

out = {}
for name, n_train in (('airfoil_3x16', 96), ('bikesharing_2x16', 160), ('covertype_ref', 120)):
    spec = o.make_spec(name)
    X, y, Xt, yt = o.synthetic_data(name, n_train=n_train, n_test=24)
    C, S, d = 2, 4, spec.n_params
    th0 = o.synthetic_theta0(spec, C).astype(np.float32)
    rng = np.random.default_rng(2024)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    z = rng.standard_normal((S, C, d)).astype(np.float32)
    f = lambda t: o.logpost_value_and_grad(spec, t, X.astype(np.float64), y)
    pos, mom, lp, dE, grad0 = [], [], [], [], []
    for c in range(C):
        st = o.mclmc_init(f, th0[c].astype(np.float64), z0[c].astype(np.float64))
        grad0.append(st.logdensity_grad)
        e = []
        for s in range(S):
            st, info = o.mclmc_step(f, st, 0.02, 15.0, z[s, c].astype(np.float64))
            e.append(info.energy_change)
        pos.append(st.position); mom.append(st.momentum); lp.append(st.logdensity); dE.append(e)
    lv = np.stack([o.forward(spec, p, Xt.astype(np.float64)) for p in pos])
    out.update({f'{name}.X': X, f'{name}.y': y, f'{name}.Xt': Xt, f'{name}.yt': yt, f'{name}.theta0': th0, f'{name}.z0': z0,
                f'{name}.z': z, f'{name}.position': np.array(pos), f'{name}.momentum': np.array(mom),
                f'{name}.logdensity': np.array(lp), f'{name}.energy_change': np.array(dE), f'{name}.grad0': np.array(grad0),
                f'{name}.lppd': np.float64(o.lppd(o.pointwise_lppd(spec, lv[:, None], yt)))})
np.savez_compressed(Path(__file__).resolve().parent / 'oracle_vectors.npz', **out)
print('written', sum(v.nbytes for v in out.values()) // 1024, 'KiB')
