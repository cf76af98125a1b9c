"""Multi-GPU plumbing.  The ensemble shards naturally: chains are independent (the reference runs them under
`jax.pmap` with no collective, src/training/sampling.py:181-184), so each rank owns a contiguous block of chains
and the only exchange step is the final merge of the per-chain online logsumexp states of the posterior
predictive (and of scalar diagnostics).  torch.distributed is the transport: NCCL over NVLink on GPUs, gloo in
the CPU tests."""
from __future__ import annotations

import numpy as np


def partition_chains(n_chains: int, world_size: int, rank: int) -> range:
    """Contiguous block of chain ids for `rank` (the reference's `train_plan` waves, trainer.py:80-82)."""
    base, rem = divmod(n_chains, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def merge_lppd_states(m: np.ndarray, s: np.ndarray, count: int, device=None):
    """All-gather the per-chain (running max, scaled sum-exp) states [C_local, Nt] of every rank and return
    (LPPD, total samples).  LPPD = mean_n logsumexp_{c,s}(lp) - log(C*S) (src/inference/metrics.py:296-312).
    Works without an initialised process group (single process)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return _lppd(m, s, m.shape[0] * count), m.shape[0] * count
    dev = device if device is not None else ('cuda' if dist.get_backend() == 'nccl' else 'cpu')
    world = dist.get_world_size()
    # chains per rank may differ by one: exchange sizes first, then pad to the maximum
    n_local = torch.tensor([m.shape[0], count], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    cmax = int(max(int(t[0]) for t in sizes))
    pad_m = np.full((cmax, m.shape[1]), -np.inf, np.float32)
    pad_s = np.zeros((cmax, m.shape[1]), np.float32)
    pad_m[:m.shape[0]], pad_s[:m.shape[0]] = m, s
    local = torch.from_numpy(np.stack([pad_m, pad_s])).to(dev)
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ms, ss, total = [], [], 0
    for t, sz in zip(gathered, sizes):
        c, cnt = int(sz[0]), int(sz[1])
        arr = t.cpu().numpy()
        ms.append(arr[0, :c]); ss.append(arr[1, :c])
        total += c * cnt
    M, S = np.concatenate(ms), np.concatenate(ss)
    return _lppd(M, S, total), total


def _lppd(m, s, total):
    from .engine import lppd_from_state
    return lppd_from_state(m, s, total)


def allreduce_mean_max(values: np.ndarray, device=None):
    """Diagnostics reduction: (mean over ranks, max over ranks) of a small float vector."""
    import torch
    import torch.distributed as dist
    v = np.asarray(values, np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return v, v
    dev = device if device is not None else ('cuda' if dist.get_backend() == 'nccl' else 'cpu')
    a = torch.from_numpy(v).to(dev)
    b = a.clone()
    dist.all_reduce(a, op=dist.ReduceOp.SUM)
    dist.all_reduce(b, op=dist.ReduceOp.MAX)
    return (a / dist.get_world_size()).cpu().numpy(), b.cpu().numpy()
