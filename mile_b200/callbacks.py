"""Sample writer (mirror of src/training/callbacks.py:17-44).  The reference fires one io_callback per kept
position from inside the jitted scan; here kept positions accumulate in an HBM ring and are written in
batches by a small thread pool, with the reference's file layout: samples/{chain}/sample_{n}.npz, members in
leaf (sorted-key) order 'fcn.layer0.bias', 'fcn.layer0.kernel', ..."""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

from .utils import get_flattened_keys, sorted_tree


def save_position(position: dict, base: Path, idx, n: int):
    """callbacks.py:17-44: one compressed npz per (chain idx, sample n)."""
    position = sorted_tree(position)
    names = get_flattened_keys(position)
    leaves = _leaves(position)
    path = Path(base) / f'{int(np.asarray(idx).item())}/sample_{int(n)}.npz'
    path.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(path, **{name: np.array(leaf) for name, leaf in zip(names, leaves)})
    return position


def _leaves(tree):
    out = []
    for k in tree:
        v = tree[k]
        out.extend(_leaves(v) if isinstance(v, dict) else [v])
    return out


class SampleWriter:
    """Asynchronous batch writer of kept positions."""

    def __init__(self, spec, base: Path, step_ids, max_workers: int = 8):
        self.spec, self.base, self.step_ids = spec, Path(base), [int(s) for s in step_ids]
        self.pool = ThreadPoolExecutor(max_workers=max_workers)
        self.futures = []
        for s in self.step_ids:
            (self.base / str(s)).mkdir(parents=True, exist_ok=True)

    def submit(self, samples: np.ndarray, sample_indices):
        """samples [S, C, d]; sample_indices: the step index n of each kept position."""
        for c, cid in enumerate(self.step_ids):
            self.futures.append(self.pool.submit(self._write_chain, samples[:, c].copy(), cid, list(sample_indices)))

    def _write_chain(self, block, cid, idxs):
        for k, n in enumerate(idxs):
            save_position(self.spec.unravel(block[k]), self.base, np.asarray(cid), n)

    def close(self):
        for f in self.futures:
            f.result()
        self.pool.shutdown()
