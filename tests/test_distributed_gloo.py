"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: chain partition + final LPPD merge."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mile_oracle as o


def test_partition_chains_covers_everything():
    from mile_b200.distributed import partition_chains
    for n, w in [(12, 8), (10, 4), (1024, 8), (3, 4), (12, 1)]:
        parts = [list(partition_chains(n, w, r)) for r in range(w)]
        assert sum(parts, []) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, lp, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mile_b200.distributed import allreduce_mean_max, merge_lppd_states, partition_chains
    mine = list(partition_chains(lp.shape[0], world, rank))
    m = np.full((len(mine), lp.shape[2]), -np.inf, np.float32)
    s = np.zeros_like(m)
    for t in range(lp.shape[1]):
        m, s = o.online_logsumexp_update(m, s, lp[mine, t].astype(np.float32))
    val, total = merge_lppd_states(m, s, lp.shape[1])
    mean, mx = allreduce_mean_max(np.array([float(rank)]))
    out[rank] = (val, total, float(mean[0]), float(mx[0]))
    dist.destroy_process_group()


def test_lppd_merge_world_size_2():
    rng = np.random.default_rng(0)
    lp = rng.standard_normal((5, 7, 11)) * 2          # 5 chains (uneven split 3 + 2), 7 samples, 11 test points
    want = o.lppd(lp)
    ctx = mp.get_context('spawn')
    mgr = ctx.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lp, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for r in range(2):
        val, total, mean, mx = out[r]
        assert total == 35 and abs(val - want) < 1e-5 and mean == 0.5 and mx == 1.0
