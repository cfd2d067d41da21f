"""CPU tests of the C ABI surface: the library loads, exports every symbol that
include/centrolign_b200.h declares, validates arguments, and refuses to compute without a GPU
(no CPU fallback).  No kernels are launched here."""
import ctypes
import os
import re

import pytest

from centrolign_b200 import popoa
from centrolign_b200.batch import AlignmentParameters, batch_from_graph_pairs, graph_from_edges

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib_or_skip():
    if not os.path.exists(popoa._LIB_PATH):
        pytest.skip("libcentrolign_b200.so not built (nvcc missing?)")
    return popoa.load_library()


def test_header_symbols_exported():
    lib = _lib_or_skip()
    header = open(os.path.join(ROOT, "include", "centrolign_b200.h")).read()
    declared = set(re.findall(r"\b(clb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(popoa.EXPORTS), declared ^ set(popoa.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layout_matches_header():
    assert ctypes.sizeof(popoa._Params) == 4 + 4 + 4 + 12 + 12
    assert ctypes.sizeof(popoa._GraphBatch) == 9 * 8
    assert ctypes.sizeof(popoa.BatchStats) == 9 * 8


def test_argument_validation_and_no_cpu_fallback():
    lib = _lib_or_skip()
    g = graph_from_edges("ACG", [(0, 1), (1, 2)], [0], [2])
    batch = batch_from_graph_pairs([(g, g)])
    if lib.clb_device_count() == 0:
        with pytest.raises(popoa.ClbError) as ei:
            popoa.po_poa_batch(batch, AlignmentParameters())
        assert ei.value.code == 3 and "no CPU fallback" in str(ei.value)
    with pytest.raises(popoa.ClbError) as ei:
        popoa.po_poa_batch(batch, AlignmentParameters(1, 1, (1, 2, 3, 4), (4, 3, 2, 1)))
    assert ei.value.code == 1


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(popoa, "_lib", None)
    monkeypatch.setattr(popoa, "_LIB_PATH", "/nonexistent/libcentrolign_b200.so")
    g = graph_from_edges("A", [], [0], [0])
    with pytest.raises(popoa.ClbError):
        popoa.po_poa_batch(batch_from_graph_pairs([(g, g)]), AlignmentParameters())
