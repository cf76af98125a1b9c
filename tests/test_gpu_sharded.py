"""Data-sharded variant (rows split across ranks + gradient all-reduce): on one GPU (world = 1) the step loop made
of EVAL + integrator launches must reproduce the fused persistent kernel; with >= 2 GPUs a torchrun job checks the
2-rank result against the single-GPU full-data result."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import mile_oracle as o

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize('name', ['covertype_ref', 'airfoil_3x16'])
def test_world1_sharded_equals_fused(name):
    from mile_b200 import Ensemble, FCNSpec, ShardedEnsemble
    ospec = o.make_spec(name)
    X, y, _, _ = o.synthetic_data(name, n_train=900)
    C, d, n = 3, ospec.n_params, 4
    spec = FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task)
    th0 = o.synthetic_theta0(ospec, C)
    rng = np.random.default_rng(0)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    z = rng.standard_normal((n, C, d)).astype(np.float32)
    a = Ensemble(spec, C); a.set_data(X, y); a.init(th0, z0)
    sa, ia = a.sample(n, 0.01, 20.0, z=z, n_thinning=2, info=True)
    b = ShardedEnsemble(spec, C, rank=0, world=1); b.set_data(X, y); b.init(th0, z0)
    sb, ib = b.sample(n, 0.01, 20.0, z=z, n_thinning=2, info=True)
    # (the two paths reduce over different thread groupings: agreement to fp32 rounding, not bit-exact)
    for u, v in zip(a.get_state(), b.get_state()):
        assert rel(v, u) <= 1e-5
    assert rel(sb, sa) <= 1e-5 and np.max(np.abs(ib - ia)) <= 2e-5 * np.max(np.abs(ia[..., 0]))
    # tuning loop as well
    tc = a.tune_cfg(4, 4, 0.5, 0.1, 1.5, 100)
    zt = rng.standard_normal((8, C, d)).astype(np.float32)
    a.init(th0, z0); a.tune_reset(0.01); ta = a.tune(8, 0, tc, z=zt, info=True); a.tune_finish_phase2()
    b.init(th0, z0); b.tune_reset(0.01); tb = b.tune(8, 0, tc, z=zt, info=True); b.tune_finish_phase2()
    # dE at tiny step sizes is dominated by fp32 rounding of the log-density (different summation orders in the two
    # paths), so the adaptive arithmetic is checked for self-consistency: the oracle's predictor driven by the
    # sharded path's own energy changes must reproduce its step-size trajectory; the first step (identical inputs)
    # must agree with the fused kernel within the fp32 energy resolution.
    lp_scale = np.abs(a.get_state()[2]).max()
    assert np.max(np.abs(tb[0, :, 0] - ta[0, :, 0])) <= 2e-5 * lp_scale
    cfg = o.TuneConfig(4, 4, 0, 0.5, 0.1, 1.5, 100, 0.01)
    for c in range(C):
        ts = o.tune_init(cfg, d, np.float64)
        for i in range(8):
            assert tb[i, c, 3] == 1.0
            ts = o.tune_update(cfg, ts._replace(step_size_max=np.float64(tb[i, c, 2])), np.zeros(d), np.float64(tb[i, c, 0]), True, i)
            assert abs(tb[i, c, 1] - ts.step_size) <= 2e-5 * ts.step_size
    eb, Lb, _ = b.get_tuning()
    assert np.all(np.isfinite(eb)) and np.all(np.isfinite(Lb)) and np.all(Lb > 0)
    a.close(); b.close()


def test_two_rank_sharded_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29533', str(ROOT / 'tests' / 'sharded_worker.py')],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'SHARDED-OK' in r.stdout
