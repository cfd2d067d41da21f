import os, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import importlib.util
spec = importlib.util.spec_from_file_location("t", "tests/test_popoa_gpu.py")
t = importlib.util.module_from_spec(spec); spec.loader.exec_module(t)
from centrolign_b200.batch import *
from centrolign_b200.popoa import po_poa_batch
oracle = CpuChecker("port")
def check(batch, p, tag):
    scores, alns = po_poa_batch(batch, p)
    bad = 0
    for w in range(batch.n_windows):
        s, a = oracle.po_poa(batch, w, p)
        if s != scores[w] or not np.array_equal(a, alns[w]):
            bad += 1
            first = -1
            if len(a) == len(alns[w]):
                d = np.nonzero((np.asarray(a) != np.asarray(alns[w])).any(axis=1))[0]
                first = (int(d[0]), a[d[0]].tolist(), alns[w][d[0]].tolist(), int(d[-1])) if len(d) else -1
            print(f"  {tag}: window {w} n1={batch.g1.n(w)} n2={batch.g2.n(w)} score {scores[w]} vs {s} len {len(alns[w])} vs {len(a)} first {first}")
    print(f"{tag}: {bad} bad of {batch.n_windows}")
rng = np.random.default_rng(4242 + 3)
pairs = []
for k in range(24):
    sides = []
    for _ in range(2):
        labels, edges = t._irregular_chain(rng, int(rng.integers(400, 1300)))
        src, snk = sources_and_sinks(len(labels), edges)
        sides.append(graph_from_edges(labels, edges, src, snk))
    pairs.append(tuple(sides))
b1 = batch_from_graph_pairs(pairs)
b2 = synth_windows(6, first_index=50, seed=5, len_min=1500, len_max=3000)
b3 = synth_windows(4, first_index=70, seed=6, len_min=1500, len_max=3000, alt_period=0, snp_rate=0.0)
check(b1, t.PROD, "irregular"); check(b2, t.PROD, "synth"); check(b3, t.PROD, "linear")
