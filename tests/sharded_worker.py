"""torchrun worker for tests/test_gpu_sharded.py: 2 ranks, rows of a covertype-shaped problem split in two,
NCCL all-reduce of the packed [C,d+1] gradient buffer; rank 0 compares against the full-data single-GPU run."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import mile_oracle as o  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    from mile_b200 import Ensemble, FCNSpec, ShardedEnsemble
    name = 'covertype_ref'
    ospec = o.make_spec(name)
    X, y, _, _ = o.synthetic_data(name, n_train=3001)       # odd row count: ragged shards
    C, d, n = 4, ospec.n_params, 5
    spec = FCNSpec(ospec.n_features, ospec.widths, ospec.activation, ospec.task)
    th0 = o.synthetic_theta0(ospec, C)
    rng = np.random.default_rng(0)
    z0 = rng.standard_normal((C, d)).astype(np.float32)
    z = rng.standard_normal((n, C, d)).astype(np.float32)
    sh = ShardedEnsemble(spec, C, device=local)
    rows = ShardedEnsemble.shard_rows(X.shape[0], rank, world)
    sh.set_data(X[rows], y[rows])
    sh.init(th0, z0)
    s_sh, i_sh = sh.sample(n, 0.01, 30.0, z=z, info=True)
    st_sh = sh.get_state()
    # replicas must be bit-identical across ranks
    t = torch.from_numpy(st_sh[0]).cuda()
    g = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    assert all(torch.equal(g[0], gi) for gi in g), 'chain state diverged between ranks'
    if rank == 0:
        full = Ensemble(spec, C, device=local)
        full.set_data(X, y)
        full.init(th0, z0)
        s_f, i_f = full.sample(n, 0.01, 30.0, z=z, info=True)
        rel = lambda a, b: np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b)
        for a, b in zip(st_sh, full.get_state()):
            assert rel(a, b) <= 1e-5, rel(a, b)
        assert rel(s_sh, s_f) <= 1e-5
        assert np.max(np.abs(i_sh - i_f)) <= 2e-5 * np.max(np.abs(i_f[..., 0]))
        print('SHARDED-OK', rel(st_sh[0], full.get_state()[0]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
