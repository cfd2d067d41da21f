import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")  # run from the repo root: compute-sanitizer --tool memcheck python tools/sanitize_tiled.py
import numpy as np
from centrolign_b200.batch import *
from centrolign_b200.popoa import po_poa_batch
import test_popoa_gpu as t
from checkers import CpuChecker
oracle = CpuChecker("port")
rng = np.random.default_rng(5)
pairs = []
for k in range(6):
    sides = []
    for side in range(2):
        labels, edges = t._irregular_chain(rng, 500 + 60 * k)
        src, snk = sources_and_sinks(len(labels), edges)
        sides.append(graph_from_edges(labels, edges, src, snk))
    pairs.append(tuple(sides))
b = concat_batches([batch_from_graph_pairs(pairs), synth_windows(2, first_index=50, seed=5, len_min=900, len_max=1300)])
scores, alns = po_poa_batch(b, t.PROD)
ok = all(oracle.po_poa(b, w, t.PROD)[0] == scores[w] and np.array_equal(oracle.po_poa(b, w, t.PROD)[1], alns[w]) for w in range(b.n_windows))
print("windows", b.n_windows, "parity", ok)
