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


def test_chain_job_entry_points_without_a_device():
    """The two-halves form of the batched chaining call: argument errors and the no-device case come back as codes (no crash,
    no job object left behind); running zero jobs and destroying NULL are allowed; the context warm-up is silent."""
    from centrolign_b200 import chain

    lib = _lib_or_skip()
    chain._bind(lib)
    job = ctypes.c_void_p(None)
    assert lib.clb_chain_job_create(0, None, None, None, None, None, None, ctypes.byref(job)) == 1 and not job.value  # CLB_EINVAL
    assert lib.clb_chain_job_create(0, None, None, None, None, None, None, None) == 1
    assert lib.clb_chain_jobs_run(0, 0, None) == 0
    assert lib.clb_chain_jobs_run(0, 1, None) == 1
    lib.clb_chain_job_destroy(None)
    lib.clb_warm_up.restype = None
    lib.clb_warm_up.argtypes = [ctypes.c_int]
    lib.clb_warm_up(0)
    lib.clb_warm_up(0)  # only the first call does anything
    if lib.clb_device_count() == 0:
        import sys

        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import test_chain_gpu as t

        cp, keep = chain._c_problem(t._handmade(3))
        n = ctypes.c_int64(0)
        out = (ctypes.c_int64 * 4)()
        rc = lib.clb_chain_job_create(0, ctypes.byref(cp), None, None, out, ctypes.byref(n), None, ctypes.byref(job))
        assert rc == 3 and not job.value  # CLB_ECUDA: no device, no CPU fallback


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(popoa, "_lib", None)
    monkeypatch.setattr(popoa, "_LIB_PATH", "/nonexistent/libcentrolign_b200.so")
    g = graph_from_edges("A", [], [0], [0])
    with pytest.raises(popoa.ClbError):
        popoa.po_poa_batch(batch_from_graph_pairs([(g, g)]), AlignmentParameters())


def _pred_lists(n, edges):
    preds = [[] for _ in range(n)]
    for a, b in edges:
        preds[b].append(a)
    return preds


def test_topological_numbering_is_valid_and_keeps_bubbles_near():
    """Host logic of the gap-fill path that needs no GPU: the numbering the flattening code gives the nodes
    (clb_topological_ranks) is a topological order of any DAG, rejects cycles, and keeps both alleles of adjacent
    SNP bubbles within distance 2 of their predecessors (a plain LIFO order leaves distance 3,
    which sends the strip through the generic step of the fill kernel)."""
    import numpy as np
    from centrolign_b200.batch import random_bubble_chain, random_dag
    _lib_or_skip()
    rng = np.random.default_rng(3)
    for t in range(60):
        n = int(rng.integers(1, 60))
        _, edges = random_dag(rng, n, int(rng.integers(0, 3 * n)))
        ranks = popoa.topological_ranks(_pred_lists(n, edges))
        assert sorted(ranks.tolist()) == list(range(1, n + 1))
        assert all(ranks[a] < ranks[b] for a, b in edges)
    for t in range(20):  # SNP bubbles at random, many of them adjacent: still a topological order
        labels, edges = random_bubble_chain(rng, int(rng.integers(50, 400)), snp_rate=0.3, del_rate=0.0)
        ranks = popoa.topological_ranks(_pred_lists(len(labels), edges)).astype(np.int64)
        assert all(ranks[a] < ranks[b] for a, b in edges)
    # isolated PAIRS of adjacent SNP bubbles (three in a row cannot all stay within distance 2 in any order)
    length = 200
    edges = [(i - 1, i) for i in range(1, length)]
    n = length
    for p in range(5, length - 5, 10):
        for q in (p, p + 1):
            edges += [(q - 1, n), (n, q + 1)]
            n += 1
    for seed in range(5):
        e = list(edges)
        np.random.default_rng(seed).shuffle(e)  # the edge order decides the successor order the stack sees
        ranks = popoa.topological_ranks(_pred_lists(n, e)).astype(np.int64)
        assert all(ranks[a] < ranks[b] for a, b in e)
        assert max(ranks[b] - ranks[a] for a, b in e) <= 2, seed
    with pytest.raises(popoa.ClbError) as ei:
        popoa.topological_ranks([[1], [0]])
    assert ei.value.code == 2
    with pytest.raises(popoa.ClbError) as ei:
        popoa.topological_ranks([[5]])
    assert ei.value.code == 1
