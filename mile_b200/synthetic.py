"""Synthetic workloads of the named BASELINE.json shapes: seeded inputs for bench.py, smoke()
and the tests.  numpy only; no arithmetic of the hot path lives here.

    X ~ N(0, 1) f32 [N, F] (the reference z-scores features, tabular.py:135-147)
    regression      y = tanh(X w0) + 0.3 eps, z-scored
    classification  y = argmax(X Wc + Gumbel)
    warm start      theta0_c ~ N(0, scale^2) with chain seed seed0 + c (stand-in for the deep-ensemble members)
"""
from __future__ import annotations

import math

import numpy as np

CONFIGS = {
    # name: (N_train, N_test, F, widths, activation, task)
    'airfoil_3x16': (1052, 301, 5, (16, 16, 16, 2), 'relu', 'regr'),
    'airfoil_2x16': (1052, 301, 5, (16, 16, 2), 'relu', 'regr'),
    'bikesharing_2x16': (12165, 3476, 12, (16, 16, 2), 'relu', 'regr'),
    'protein_2x16': (32010, 9146, 9, (16, 16, 2), 'relu', 'regr'),
    'covertype_ref': (3200, 4000, 54, (32, 7), 'sigmoid', 'class'),
    'covertype_full': (232404, 290506, 54, (32, 7), 'sigmoid', 'class'),
    'wide_4x256': (12165, 3476, 12, (256, 256, 256, 256, 2), 'relu', 'regr'),
}


def workload_spec(name: str):
    """FCNSpec of a named workload."""
    from .engine import FCNSpec
    _, _, F, widths, act, task = CONFIGS[name]
    return FCNSpec(F, widths, act, task)


def synthetic_data(name: str, seed: int = 1234, n_train: int | None = None, n_test: int | None = None):
    """Returns X, y, Xt, yt (fp32 / int32)."""
    N, Nt, F, widths, _, task = CONFIGS[name]
    N = n_train or N
    Nt = n_test or Nt
    rng = np.random.default_rng(seed)
    Xall = rng.standard_normal((N + Nt, F)).astype(np.float32)
    if task == 'regr':
        w0 = rng.standard_normal(F).astype(np.float32) / np.float32(math.sqrt(F))
        yall = np.tanh(Xall @ w0) + np.float32(0.3) * rng.standard_normal(N + Nt).astype(np.float32)
        yall = ((yall - yall.mean()) / yall.std()).astype(np.float32)
    else:
        K = widths[-1]
        Wc = rng.standard_normal((F, K)).astype(np.float32)
        yall = np.argmax(Xall @ Wc + rng.gumbel(size=(N + Nt, K)).astype(np.float32), axis=1).astype(np.int32)
    return Xall[:N], yall[:N], Xall[N:], yall[N:]


def synthetic_theta0(n_params: int, n_chains: int, scale: float = 0.3, seed0: int = 1000):
    return np.stack([np.random.default_rng(seed0 + c).standard_normal(n_params).astype(np.float32)
                     * np.float32(scale) for c in range(n_chains)])
