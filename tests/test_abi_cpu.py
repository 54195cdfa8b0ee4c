"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/iris_b200.h declares
(no compute calls without a GPU), and its hash-grid level table is the oracle's."""
import os
import re

import numpy as np

from iris_b200 import _capi as C
from oracle import field as OF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "iris_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(iris_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(C.PROTOTYPES), declared ^ set(C.PROTOTYPES)
    L = C.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.iris_version()


def test_level_table_matches_oracle():
    from iris_b200.core import field_levels
    sc, res, size, off, n = field_levels()
    assert n == OF.N_ENTRIES
    for l, (scale, r, s, o) in enumerate(OF.LEVELS):
        assert np.float32(scale) == sc[l] and r == res[l] and s == size[l] and o == off[l], l


def test_no_compute_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from iris_b200.core import Scene
    with pytest.raises(RuntimeError):
        Scene(np.zeros((3, 3), np.float32), np.array([[0, 1, 2]], np.int32))


def test_product_does_not_import_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "iris_b200")):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)
