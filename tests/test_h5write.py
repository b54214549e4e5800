"""The HDF5 writer of the batched observers (arboris_b200/h5write.py), round-tripped through the
reader that parses the REFERENCE's own .h5 fixtures (oracle/h5lite.py; test infrastructure)."""
import os
import struct

import numpy as np

from arboris_b200 import h5write
from oracle import h5lite

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same(a, b):
    assert sorted(a) == sorted(b)
    for k in a:
        if isinstance(a[k], dict):
            _same(a[k], b[k])
        else:
            assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k


def test_roundtrip_nested_groups(tmp_path):
    rng = np.random.default_rng(0)
    tree = {"timeline": np.arange(0., 1., .01),
            "transforms": {"Arm": rng.normal(size=(99, 4, 4)), "Forearm": rng.normal(size=(99, 3, 4, 4))},
            "gpositions": {"Shoulder": rng.normal(size=(99, 1))},
            "deep": {"er": {"still": {"x": np.array([1., 2., 3.])}}},
            "scalar_like": np.array([42.]), "empty": np.zeros((0, 6))}
    p = str(tmp_path/"t.h5")
    h5write.write(p, tree)
    _same(tree, h5lite.read(p))
    raw = open(p, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0          # superblock version 0
    assert struct.unpack("<Q", raw[40:48])[0] == len(raw)            # end-of-file address


def test_group_larger_than_one_symbol_table_node(tmp_path):
    tree = {"g": {"body%03d" % i: np.full((2, 2), float(i)) for i in range(300)}}
    p = str(tmp_path/"big.h5")
    h5write.write(p, tree)
    _same(tree, h5lite.read(p))


def test_rewrites_the_reference_fixture_layout(tmp_path):
    """tests/simplearm_flat.h5 of the reference (converted in tests/golden/reference_h5.npz):
    writing the same names / shapes and reading them back gives the same arrays."""
    ref = np.load(os.path.join(GOLDEN, "reference_h5.npz"))
    tree = {"timeline": ref["simplearm_flat/timeline"],
            "transforms": {k: ref["simplearm_flat/transforms/" + k] for k in ("Arm", "Forearm", "Hand")}}
    p = str(tmp_path/"simplearm_flat.h5")
    h5write.write(p, tree)
    _same(tree, h5lite.read(p))
