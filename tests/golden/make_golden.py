"""Generate tests/golden/popoa_golden.npz by running the UNMODIFIED reference
(oracle/_ref/libclref.so, built by oracle/Makefile from /root/reference sources) on seeded
windows.  Run in the build container only:  python tests/golden/make_golden.py

Stored per case: the window (flat batch arrays), the parameter set, and the reference's
score + alignment.  tests/test_oracle.py (CPU) and tests/test_popoa_gpu.py (GPU) replay them.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from centrolign_b200.batch import (AlignmentParameters, batch_from_graph_pairs, concat_batches,  # noqa: E402
                                   graph_from_edges, random_bubble_chain, random_dag, sources_and_sinks, synth_windows)

PARAM_SETS = [
    AlignmentParameters(1, 1, (1,), (1,)),  # the reference's unit-test parameters (test_alignment.cpp:711-715)
    AlignmentParameters(2, 2, (0,), (1,)),  # zero gap-open: maximal tie pressure
    AlignmentParameters(20, 80, (60,), (30,)),  # production, truncated to 1 piece (src/stitcher.cpp:54-56)
    AlignmentParameters(20, 80, (60, 800), (30, 5)),  # production, 2 pieces
    AlignmentParameters(20, 80, (60, 800, 2500), (30, 5, 1)),  # production (src/stitcher.cpp:13-22)
    AlignmentParameters(3, 2, (1, 4, 9), (3, 2, 1)),  # small 3-piece: pieces cross within short gaps
]


def reaches(n, edges, sources, sinks):
    succ = [[] for _ in range(n)]
    for a, b in edges:
        succ[a].append(b)
    seen = set(sources)
    stack = list(sources)
    while stack:
        v = stack.pop()
        for u in succ[v]:
            if u not in seen:
                seen.add(u)
                stack.append(u)
    return any(s in seen for s in sinks)


def random_window(rng, kind):
    sides = []
    for _ in range(2):
        if kind == "dag":  # test_alignment.cpp:1656-1694: 5-10 nodes, 8-18 edges, random sources/sinks
            n = int(rng.integers(5, 11))
            labels, edges = random_dag(rng, n, int(rng.integers(8, 19)))
            while True:
                src = sorted(set(int(x) for x in rng.integers(0, n, int(rng.integers(1, 3)))))
                snk = sorted(set(int(x) for x in rng.integers(0, n, int(rng.integers(1, 3)))))
                if reaches(n, edges, src, snk):
                    break
        elif kind == "dag_big":
            n = int(rng.integers(20, 90))
            labels, edges = random_dag(rng, n, int(rng.integers(n, 3 * n)), alphabet="AC")
            src, snk = sources_and_sinks(n, edges)
        else:  # bubble chains: low-entropy backbone, SNPs, deletion edges
            length = int(rng.integers(3, 140)) if kind == "chain" else int(rng.integers(1, 6))
            labels, edges = random_bubble_chain(rng, length)
            src, snk = sources_and_sinks(len(labels), edges)
            if rng.random() < 0.3:
                rng.shuffle(src)
                rng.shuffle(snk)
        sides.append(graph_from_edges(labels, edges, src, snk))
    return tuple(sides)


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from checkers import CpuChecker
    ref = CpuChecker("reference")
    rng = np.random.default_rng(20261017)
    pairs = []
    for kind, count in (("dag", 160), ("dag_big", 30), ("chain", 90), ("tiny", 40)):
        pairs += [random_window(rng, kind) for _ in range(count)]
    batches = [batch_from_graph_pairs(pairs)]
    # a few HOR-like synthetic windows (csrc/synth.c), small enough for the fixture
    batches.append(synth_windows(12, first_index=0, seed=7, len_min=40, len_max=700, alt_len=31, alt_period=150))
    batch = concat_batches(batches)
    nw = batch.n_windows
    param_idx = np.arange(nw) % len(PARAM_SETS)
    scores = np.zeros(nw, np.int64)
    alns, aln_off = [], np.zeros(nw + 1, np.int64)
    for w in range(nw):
        s, a = ref.po_poa(batch, w, PARAM_SETS[param_idx[w]])
        scores[w] = s
        alns.append(a)
        aln_off[w + 1] = aln_off[w] + len(a)
    out = {"param_sets": np.stack([p.packed() for p in PARAM_SETS]),
           "param_num_pw": np.asarray([p.num_pw for p in PARAM_SETS], np.int32), "param_idx": param_idx.astype(np.int32),
           "score": scores, "aln": np.concatenate(alns).astype(np.int32), "aln_off": aln_off}
    for name, side in (("g1", batch.g1), ("g2", batch.g2)):
        for f in ("node_off", "label", "edge_off", "pred_off", "pred", "src_off", "src", "snk_off", "snk"):
            out[f"{name}_{f}"] = getattr(side, f)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "popoa_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {nw} windows, {int(batch.cells().sum())} cells, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
