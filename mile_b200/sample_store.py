"""Compact sample store.

The reference writes one compressed npz per kept position and chain (src/training/callbacks.py:17-44: 12 000 files for
the illustrative run) and re-reads them one by one (src/training/utils.py:131-175).  Once the sampler runs at
microseconds per step that file traffic is the dominant residual cost, so the kept positions can instead go to ONE
array on disk in the layout the device already holds:

    <dir>/samples.npy     float32 [C, S, d]  (chain, kept sample, flat parameter in ravel_pytree order), memory-mapped
    <dir>/index.json      chain ids, step index n of every kept sample, leaf names / shapes / offsets, model shape

`export_npz` is the compatibility exporter: it reproduces the reference layout samples/{chain}/sample_{n}.npz (same
member names and order) so that `inference.ipynb` and `load_samples_from_dir` of the reference work unchanged;
`from_npz_dir` imports a reference-written directory.  `mile_b200.utils.load_samples_from_dir` reads either form.
"""
from __future__ import annotations

import json
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

DATA, INDEX = 'samples.npy', 'index.json'


def _leaf_table(spec):
    """(name, shape, offset) of every leaf in npz member order: sorted keys, per layer bias then kernel."""
    b_off, k_off = spec.offsets()
    d = spec.dims
    rows = []
    for l in spec.layer_order:
        rows.append((f'fcn.layer{l}.bias', (d[l + 1],), b_off[l]))
        rows.append((f'fcn.layer{l}.kernel', (d[l], d[l + 1]), k_off[l]))
    return rows


class SampleStore:
    def __init__(self, path: Path, meta: dict, array: np.memmap):
        self.path, self.meta, self.array = Path(path), meta, array

    # ---- writing ----------------------------------------------------------------------------------------
    @classmethod
    def create(cls, path, spec, step_ids, n_samples: int) -> 'SampleStore':
        path = Path(path)
        path.mkdir(parents=True, exist_ok=True)
        C, d = len(step_ids), spec.n_params
        meta = {'format': 'mile_b200.sample_store/1', 'chains': [int(s) for s in step_ids], 'n_params': d,
                'n_samples': int(n_samples), 'filled': 0, 'sample_index': [],
                'leaves': [{'name': n, 'shape': list(s), 'offset': int(o)} for n, s, o in _leaf_table(spec)],
                'model': {'n_features': spec.n_features, 'widths': list(spec.widths), 'activation': spec.activation,
                          'task': spec.task}}
        arr = np.lib.format.open_memmap(path / DATA, mode='w+', dtype=np.float32, shape=(C, int(n_samples), d))
        return cls(path, meta, arr)

    def append(self, samples: np.ndarray, sample_indices):
        """samples [S_k, C, d] as the sampler returns them; sample_indices: step index n of each kept position."""
        k = samples.shape[0]
        f = self.meta['filled']
        if f + k > self.meta['n_samples']:
            raise ValueError('sample store overflow')
        self.array[:, f:f + k, :] = np.transpose(samples, (1, 0, 2))
        self.meta['filled'] = f + k
        self.meta['sample_index'].extend(int(n) for n in sample_indices)

    def close(self):
        self.array.flush()
        with open(self.path / INDEX, 'w') as fh:
            json.dump(self.meta, fh)

    # ---- reading ----------------------------------------------------------------------------------------
    @classmethod
    def exists(cls, path) -> bool:
        return (Path(path) / INDEX).exists() and (Path(path) / DATA).exists()

    @classmethod
    def open(cls, path) -> 'SampleStore':
        path = Path(path)
        meta = json.loads((path / INDEX).read_text())
        arr = np.load(path / DATA, mmap_mode='r')
        return cls(path, meta, arr)

    @property
    def samples(self) -> np.ndarray:
        """[C, S_filled, d]"""
        return self.array[:, :self.meta['filled'], :]

    def to_tree(self) -> dict:
        """Same result as the reference's load_samples_from_dir: leaves [n_chains, n_samples, ...]."""
        from .utils import _unflatten
        x = np.asarray(self.samples)
        names, arrays = [], []
        for leaf in self.meta['leaves']:
            size = int(np.prod(leaf['shape']))
            names.append(leaf['name'])
            arrays.append(x[:, :, leaf['offset']:leaf['offset'] + size].reshape(x.shape[:2] + tuple(leaf['shape'])))
        return _unflatten(names, arrays)

    # ---- compatibility with the reference layout -----------------------------------------------------------
    def export_npz(self, samples_dir, max_workers: int = 8):
        """Writes samples/{chain}/sample_{n}.npz exactly like src/training/callbacks.py:36-43."""
        import os
        from . import capi
        from .callbacks import _npy_header, write_npz_block
        samples_dir = Path(samples_dir)
        x, idx, leaves = self.samples, self.meta['sample_index'], self.meta['leaves']
        sizes = [int(np.prod(lf['shape'])) for lf in leaves]
        # the members must tile a row in order (they do: the leaf table is in ravel order)
        assert [lf['offset'] for lf in leaves] == [int(v) for v in np.cumsum([0] + sizes[:-1])] and sum(sizes) == x.shape[2]
        names = [(lf['name'] + '.npy').encode() for lf in leaves]
        headers = [_npy_header(lf['shape'], np.float32) for lf in leaves]
        paths = []
        for cid in self.meta['chains']:
            (samples_dir / str(cid)).mkdir(parents=True, exist_ok=True)
            paths += [str(samples_dir / str(cid) / f'sample_{n}.npz').encode() for n in idx]
        block = np.ascontiguousarray(x, dtype=np.float32)            # [C, S, d]: row (c, k) -> file of chain c, position k
        write_npz_block(capi.load(), paths, names, headers, sizes, block, max(1, min(max_workers, os.cpu_count() or 1)),
                        samples_dir)

    @classmethod
    def from_npz_dir(cls, samples_dir, spec, out_path) -> 'SampleStore':
        """Imports a reference-written samples/ directory."""
        samples_dir = Path(samples_dir)
        chains = sorted([d for d in samples_dir.iterdir() if d.is_dir() and d.name.isdigit()], key=lambda p: int(p.name))
        files0 = sorted(chains[0].glob('sample_*.npz'), key=lambda p: int(p.stem.split('_')[-1]))
        store = cls.create(out_path, spec, [int(c.name) for c in chains], len(files0))
        table = _leaf_table(spec)
        for ci, cd in enumerate(chains):
            files = sorted(cd.glob('sample_*.npz'), key=lambda p: int(p.stem.split('_')[-1]))
            for k, fp in enumerate(files):
                with np.load(fp) as z:
                    for name, shape, off in table:
                        store.array[ci, k, off:off + int(np.prod(shape))] = np.asarray(z[name], np.float32).ravel()
        store.meta['filled'] = len(files0)
        store.meta['sample_index'] = [int(p.stem.split('_')[-1]) for p in files0]
        store.close()
        return store


def main(argv=None):
    """python -m mile_b200.sample_store export <store dir> <samples dir>"""
    import argparse
    ap = argparse.ArgumentParser(description='sample store <-> reference npz layout')
    ap.add_argument('cmd', choices=['export', 'info'])
    ap.add_argument('store')
    ap.add_argument('samples_dir', nargs='?')
    a = ap.parse_args(argv)
    st = SampleStore.open(a.store)
    if a.cmd == 'info':
        print(json.dumps({k: st.meta[k] for k in ('chains', 'n_params', 'n_samples', 'filled')}))
    else:
        st.export_npz(a.samples_dir or Path(a.store).parent / 'samples')


if __name__ == '__main__':
    main()
