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


def _npy_header(shape, dtype) -> bytes:
    import io
    buf = io.BytesIO()
    np.lib.format.write_array_header_1_0(buf, {'descr': np.lib.format.dtype_to_descr(np.dtype(dtype)), 'fortran_order': False,
                                               'shape': tuple(int(s) for s in shape)})
    return buf.getvalue()


class _FastNpz:
    """np.savez_compressed for MANY small files of one fixed structure: the .npy headers are built once, every member is
    handed to zipfile.writestr as bytes (zlib runs outside the GIL), none of numpy's per-array Python bookkeeping.
    The files are ordinary compressed npz archives (np.load reads them; members 'fcn.layer0.bias', ...).  A reference run
    writes 12 000 of them: 0.67 ms each through np.savez_compressed was the dominant cost of a whole run
    (tools/full_run.py)."""

    def __init__(self, names, shapes, dtype=np.float32):
        self.names = [n + '.npy' for n in names]
        self.headers = [_npy_header(s, dtype) for s in shapes]
        self.sizes = [int(np.prod(s)) for s in shapes]
        self.dtype = np.dtype(dtype)

    def write(self, path, leaves):
        import zipfile
        with zipfile.ZipFile(path, 'w', zipfile.ZIP_DEFLATED) as zf:
            for name, hdr, leaf in zip(self.names, self.headers, leaves):
                zf.writestr(name, hdr + np.ascontiguousarray(leaf, dtype=self.dtype).tobytes())


def write_npz_block(lib, paths, names, headers, sizes, block, n_threads, where=''):
    """One call of the native writer (csrc/mile_npz.cu): file i = the members `names` (bytes, with '.npy') cut from row i of
    `block` [n_files, sum(sizes)] float32 in order, each behind its prebuilt .npy header."""
    import ctypes as C
    from . import capi
    nm = len(names)
    assert block.dtype == np.float32 and block.flags['C_CONTIGUOUS'] and block.size == len(paths) * sum(sizes)
    c_paths = (C.c_char_p * len(paths))(*paths)
    c_names = (C.c_char_p * nm)(*names)
    c_hdrs = (C.c_char_p * nm)(*headers)
    c_hlen = (C.c_int32 * nm)(*[len(h) for h in headers])
    c_size = (C.c_int64 * nm)(*sizes)
    rc = lib.mile_write_npz_batch(C.cast(c_paths, C.c_void_p), len(paths), C.cast(c_names, C.c_void_p),
                                  C.cast(c_hdrs, C.c_void_p), C.cast(c_hlen, C.c_void_p), C.cast(c_size, C.c_void_p),
                                  nm, capi.host_ptr(block), int(n_threads))
    if rc != 0:
        raise OSError(f'mile_write_npz_batch failed ({rc}) under {where}')


class SampleWriter:
    """Asynchronous batch writer of kept positions: every submitted block [S, C, d] becomes S x C npz files through ONE call
    of the native writer (`mile_write_npz_batch`: deflate + zip on a few host threads outside the GIL), issued from a
    background thread so that the sampler's next launch overlaps the file writing.  The flat position vector IS the
    concatenation of the leaves in the files' member order (ravel order = sorted keys, bias before kernel), so no
    per-sample unravelling happens on the host."""

    def __init__(self, spec, base: Path, step_ids, max_workers: int = 8):
        import os
        from . import capi
        self.spec, self.base, self.step_ids = spec, Path(base), [int(s) for s in step_ids]
        self.lib = capi.load()
        self.n_threads = max(1, min(int(max_workers), os.cpu_count() or 1))
        self.pool = ThreadPoolExecutor(max_workers=1)          # keeps the blocks in submission order
        self.futures = []
        proto = sorted_tree(spec.unravel(np.arange(spec.n_params, dtype=np.float32)))
        leaves = _leaves(proto)
        # the leaves of the sorted tree must tile the flat vector in order (this is what lets the writer take rows as they are)
        flat = np.concatenate([np.ravel(l) for l in leaves])
        assert np.array_equal(flat, np.arange(spec.n_params, dtype=np.float32)), 'leaf order differs from the ravel order'
        self._names = [(n + '.npy').encode() for n in get_flattened_keys(proto)]
        self._headers = [_npy_header(np.shape(l), np.float32) for l in leaves]
        self._sizes = [int(np.size(l)) for l in leaves]
        for s in self.step_ids:
            (self.base / str(s)).mkdir(parents=True, exist_ok=True)

    def submit(self, samples: np.ndarray, sample_indices):
        """samples [S, C, d]; sample_indices: the step index n of each kept position."""
        block = np.ascontiguousarray(np.transpose(samples, (1, 0, 2)), dtype=np.float32)        # [C, S, d]
        b = str(self.base)            # (plain string formatting: pathlib joins cost ~10 us each, 12 000 of them per run)
        idx = [int(n) for n in sample_indices]
        paths = [f'{b}/{cid}/sample_{n}.npz'.encode() for cid in self.step_ids for n in idx]
        self.futures.append(self.pool.submit(self._write_block, block, paths))

    def _write_block(self, block, paths):
        write_npz_block(self.lib, paths, self._names, self._headers, self._sizes, block, self.n_threads, self.base)

    def close(self):
        for f in self.futures:
            f.result()
        self.pool.shutdown()
