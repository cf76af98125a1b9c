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
    """Plain description of the nesting (the reference pickles a jax PyTreeDef here, utils.py:90-93)."""
    with open(Path(dir) / 'tree.json', 'w') as f:
        json.dump(get_flattened_keys(sorted_tree(tree)), f)


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
