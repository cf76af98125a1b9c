"""Parameter (de)serialisation with the reference's on-disk layout (mirror of src/training/utils.py:69-180
and src/utils.py:57-76).  The reference's `tree` file is a pickled jax PyTreeDef, which cannot be produced or
read without jax; the loaders here rebuild the nesting from the npz member names instead and `save_tree`
writes a plain nested-key description."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np


def sorted_tree(tree: dict) -> dict:
    """JAX flattens dicts in sorted key order; keep every nested dict sorted the same way."""
    return {k: (sorted_tree(tree[k]) if isinstance(tree[k], dict) else tree[k]) for k in sorted(tree)}


def get_flattened_keys(d: dict, sep: str = '.') -> list:
    """src/utils.py:57-76."""
    keys = []
    for k, v in d.items():
        if isinstance(v, dict):
            keys.extend([f'{k}{sep}{kk}' for kk in get_flattened_keys(v)])
        else:
            keys.append(k)
    return keys


def _unflatten(names, arrays) -> dict:
    tree: dict = {}
    for name, arr in zip(names, arrays):
        node = tree
        parts = name.split('.')
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = arr
    return tree


def save_tree(dir, tree: dict):
    """utils.py:90-93.  The reference pickles a jax PyTreeDef into `<exp>/tree`.  That object can only be built where
    jax is importable: there the same pickle is written (so `inference.ipynb` loads it unchanged); everywhere a plain
    description of the nesting goes to `tree.json`, which `load_tree` below turns into a stand-in treedef."""
    import pickle
    tree = sorted_tree(tree)
    with open(Path(dir) / 'tree.json', 'w') as f:
        json.dump(get_flattened_keys(tree), f)
    try:
        import jax  # noqa: F401  (absent on this image; present on a host that runs the reference's notebook)
        with open(Path(dir) / 'tree', 'wb') as f:
            pickle.dump(jax.tree.structure(tree), f)
    except Exception:
        pass


class TreeShim:
    """Stand-in for the pickled PyTreeDef: knows the leaf names in flatten order and rebuilds the nested dict, which is
    all the reference does with it (`jax.tree.unflatten(tree, leaves)`, utils.py:105-108,127,161)."""

    def __init__(self, names):
        self.names = list(names)
        self.num_leaves = len(self.names)

    def unflatten(self, leaves):
        leaves = list(leaves)
        if len(leaves) != self.num_leaves:
            raise ValueError(f'expected {self.num_leaves} leaves, got {len(leaves)}')
        return _unflatten(self.names, leaves)

    def __repr__(self):
        return f'TreeShim({self.names})'


def load_tree(dir):
    """utils.py:96-99: the pickled PyTreeDef when it exists and jax can unpickle it, else the tree.json stand-in."""
    import pickle
    dir = Path(dir)
    if (dir / 'tree').exists():
        try:
            with open(dir / 'tree', 'rb') as f:
                return pickle.load(f)
        except Exception:
            pass
    with open(dir / 'tree.json') as f:
        return TreeShim(json.load(f))


def save_params(dir, params: dict, idx: int | None = None):
    """utils.py:69-87: warmstart/params_{i}.npz with members in leaf order."""
    dir = Path(dir)
    dir.mkdir(parents=True, exist_ok=True)
    params = sorted_tree(params)
    if not (dir.parent / 'tree.json').exists():
        save_tree(dir.parent, params)
    names = get_flattened_keys(params)
    from .callbacks import _leaves
    name = f'params_{idx}.npz' if idx is not None else 'params.npz'
    np.savez_compressed(dir / name, **dict(zip(names, [np.asarray(l) for l in _leaves(params)])))


def load_params(params_path, tree_path=None) -> dict:
    """utils.py:102-108 (tree_path accepted for signature parity; nesting comes from the member names)."""
    with np.load(params_path) as z:
        return _unflatten(z.files, [np.array(z[k]) for k in z.files])


def load_params_batch(params_path: list, tree_path=None) -> dict:
    """utils.py:111-128: stack chains on axis 0 in numeric-suffix order; a single path is not stacked."""
    paths = sorted((Path(p) for p in params_path), key=lambda x: int(x.stem.split('_')[-1]))
    if len(paths) == 1:
        return load_params(paths[0], tree_path)
    loaded = []
    names = None
    for p in paths:
        with np.load(p) as z:
            names = z.files
            loaded.append([np.array(z[k]) for k in names])
    return _unflatten(names, [np.stack([l[i] for l in loaded]) for i in range(len(names))])


def load_samples_from_dir(dir, tree_path=None) -> dict:
    """utils.py:131-161: leaves [n_chains, n_samples, ...] from samples/{chain}/sample_{n}.npz."""
    dir = Path(dir)
    from .sample_store import SampleStore
    store = dir.parent / 'samples_store'
    has_npz = dir.exists() and any(d.is_dir() and any(d.glob('*.npz')) for d in dir.iterdir())
    if not has_npz and SampleStore.exists(store):          # compact store written by inference_loop (sample_store.py)
        return SampleStore.open(store).to_tree()
    chain_dirs = sorted([d for d in dir.iterdir() if d.is_dir()], key=lambda x: int(x.stem.split('_')[-1]))
    stacks, names = [], None
    for cd in chain_dirs:
        files = sorted([p for p in cd.iterdir() if p.suffix == '.npz'], key=lambda x: int(x.stem.split('_')[-1]))
        if not files:
            raise ValueError('No samples found in the directory')
        per = []
        for fp in files:
            with np.load(fp) as z:
                names = z.files
                per.append([np.array(z[k]) for k in names])
        stacks.append([np.stack([p[i] for p in per]) for i in range(len(names))])
    return _unflatten(names, [np.stack([c[i] for c in stacks]) for i in range(len(names))])
