"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/mile_b200.h declares (no compute calls without a GPU); host layout logic."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope='module')
def lib_path():
    from mile_b200 import build
    return build.build()


def test_header_symbols_match_binding_and_are_exported(lib_path):
    from mile_b200 import capi
    header = (ROOT / 'include' / 'mile_b200.h').read_text()
    declared = set(re.findall(r'\b(mile_[a-z0-9_]+)\s*\(', header))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    so = ctypes.CDLL(str(lib_path))
    for s in declared:
        assert hasattr(so, s), s
    assert so.mile_version() == 100


def test_struct_layouts_match_header():
    from mile_b200 import capi
    assert ctypes.sizeof(capi.ModelDesc) == 4 * (2 + 3 * capi.MILE_MAX_LAYERS + 3 + 3)
    assert ctypes.sizeof(capi.TuneCfg) == 24


def test_create_fails_loudly_without_gpu(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from mile_b200 import Ensemble, FCNSpec
    from mile_b200.capi import MileError
    with pytest.raises(MileError):
        Ensemble(FCNSpec(5, (16, 2)), 2)


def test_fcnspec_layout_roundtrip_and_matches_oracle():
    from mile_b200 import FCNSpec
    from oracle import mile_oracle as o
    for widths in [(16, 16, 16, 2), (32, 7), (3,) * 11 + (2,)]:
        spec = FCNSpec(5, widths)
        ospec = o.ModelSpec(5, widths)
        assert spec.offsets() == ospec.offsets() and spec.n_params == ospec.n_params
        th = np.arange(3 * spec.n_params, dtype=np.float32).reshape(3, -1)
        tree = spec.unravel(th)
        assert tree['fcn']['layer0']['kernel'].shape == (3, 5, widths[0])
        np.testing.assert_array_equal(spec.ravel(tree), th)
        np.testing.assert_array_equal(o.ravel_tree(ospec, spec.unravel(th[0])), th[0])


def test_product_package_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under mile_b200/ may import it."""
    for p in (ROOT / 'mile_b200').rglob('*.py'):
        src = p.read_text()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), p
