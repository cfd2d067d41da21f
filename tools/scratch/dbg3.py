import sys
sys.path.insert(0, ".")
import numpy as np
from centrolign_b200.batch import *
from centrolign_b200.popoa import po_poa_batch
oracle = CpuChecker("port")
PROD = AlignmentParameters()
rng = np.random.default_rng(5)
pairs = []
for k in range(2):
    sides = []
    for side in range(2):
        labels, edges = random_bubble_chain(rng, 600 + 40 * k, snp_rate=0.0, del_rate=0.0)
        src, snk = sources_and_sinks(len(labels), edges)
        sides.append(graph_from_edges(labels, edges, src, snk))
    pairs.append(tuple(sides))
b = batch_from_graph_pairs(pairs)
scores, alns = po_poa_batch(b, PROD)
for w in range(b.n_windows):
    print("window", w, scores[w], oracle.po_poa(b, w, PROD)[0])
