"""Compact sample store (mile_b200/sample_store.py) against the reference's on-disk layout: what the store returns, what
its exporter writes and what the npz loader reads must be the same trees (src/training/callbacks.py:17-44,
src/training/utils.py:131-175), and the tree.json stand-in must rebuild the nesting like the pickled PyTreeDef does."""
import numpy as np

from mile_b200 import FCNSpec
from mile_b200.callbacks import SampleWriter
from mile_b200.sample_store import SampleStore
from mile_b200.utils import TreeShim, load_samples_from_dir, load_tree, save_params


def _same_tree(a, b):
    assert list(a) == list(b)
    for k in a:
        if isinstance(a[k], dict):
            _same_tree(a[k], b[k])
        else:
            np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]))


def test_store_roundtrip_export_and_import(tmp_path):
    spec = FCNSpec(5, (16, 16, 2))
    rng = np.random.default_rng(0)
    C, S, thin = 3, 7, 10
    samples = rng.standard_normal((S, C, spec.n_params)).astype(np.float32)
    idx = [k * thin for k in range(S)]
    store = SampleStore.create(tmp_path / 'exp' / 'samples_store', spec, [4, 5, 11], S)
    store.append(samples[:3], idx[:3])
    store.append(samples[3:], idx[3:])
    store.close()
    st = SampleStore.open(tmp_path / 'exp' / 'samples_store')
    assert st.samples.shape == (C, S, spec.n_params)
    np.testing.assert_array_equal(st.samples, np.transpose(samples, (1, 0, 2)))
    tree = st.to_tree()
    assert tree['fcn']['layer0']['kernel'].shape == (C, S, 5, 16)
    # the loader reads the store when no npz files exist ...
    (tmp_path / 'exp' / 'samples').mkdir()
    _same_tree(load_samples_from_dir(tmp_path / 'exp' / 'samples'), tree)
    # ... the exporter reproduces the reference layout, identical to what the per-sample writer produces ...
    st.export_npz(tmp_path / 'exp' / 'samples')
    assert (tmp_path / 'exp' / 'samples' / '11' / 'sample_60.npz').exists()
    with np.load(tmp_path / 'exp' / 'samples' / '4' / 'sample_0.npz') as z:
        assert z.files == ['fcn.layer0.bias', 'fcn.layer0.kernel', 'fcn.layer1.bias', 'fcn.layer1.kernel',
                           'fcn.layer2.bias', 'fcn.layer2.kernel']
    _same_tree(load_samples_from_dir(tmp_path / 'exp' / 'samples'), tree)
    w = SampleWriter(spec, tmp_path / 'ref' / 'samples', [4, 5, 11])
    w.submit(samples, idx)
    w.close()
    _same_tree(load_samples_from_dir(tmp_path / 'ref' / 'samples'), tree)
    # ... and a reference-written directory imports into the same array
    st2 = SampleStore.from_npz_dir(tmp_path / 'ref' / 'samples', spec, tmp_path / 'ref' / 'samples_store')
    np.testing.assert_array_equal(SampleStore.open(st2.path).samples, st.samples)
    assert st2.meta['chains'] == [4, 5, 11] and st2.meta['sample_index'] == idx


def test_tree_shim_rebuilds_nesting(tmp_path):
    spec = FCNSpec(5, (16, 2))
    theta = np.arange(spec.n_params, dtype=np.float32)
    params = spec.unravel(theta)
    save_params(tmp_path / 'warmstart', params, 0)
    tree = load_tree(tmp_path)
    assert isinstance(tree, TreeShim) and tree.num_leaves == 4
    with np.load(tmp_path / 'warmstart' / 'params_0.npz') as z:
        rebuilt = tree.unflatten([z[k] for k in z.files])
    _same_tree(rebuilt, {'fcn': {'layer0': params['fcn']['layer0'], 'layer1': params['fcn']['layer1']}})


def test_native_npz_writer_matches_savez_compressed(tmp_path):
    """mile_write_npz_batch (csrc/mile_npz.cu) against np.savez_compressed, the call of the reference's save_position
    (src/training/callbacks.py:40-44): same member names in the same order, same dtypes/shapes/values, valid zip CRCs."""
    import zipfile
    from mile_b200.callbacks import save_position
    spec = FCNSpec(3, (8, 1))
    rng = np.random.default_rng(1)
    samples = rng.standard_normal((4, 2, spec.n_params)).astype(np.float32)
    samples[0, 0, :5] = [np.nan, np.inf, -np.inf, 0.0, -0.0]
    w = SampleWriter(spec, tmp_path / 'a', [0, 1], max_workers=3)
    w.submit(samples, [0, 10, 20, 30])
    w.close()
    for c in range(2):
        for k, n in enumerate([0, 10, 20, 30]):
            save_position(spec.unravel(samples[k, c]), tmp_path / 'b', np.asarray(c), n)
            pa, pb = tmp_path / 'a' / str(c) / f'sample_{n}.npz', tmp_path / 'b' / str(c) / f'sample_{n}.npz'
            assert zipfile.ZipFile(pa).testzip() is None
            with np.load(pa) as za, np.load(pb) as zb:
                assert za.files == zb.files
                for m in za.files:
                    assert za[m].dtype == zb[m].dtype and za[m].shape == zb[m].shape
                    np.testing.assert_array_equal(za[m].view(np.uint32), zb[m].view(np.uint32))
