import os, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
from centrolign_b200.batch import *
from centrolign_b200.popoa import po_poa_batch
oracle = CpuChecker("port")
PROD = AlignmentParameters()
def chain(rng, length, snp, dele, short, longs, multi, srcs):
    labels, edges = random_bubble_chain(rng, length, snp_rate=snp, del_rate=dele)
    labels = list(labels); edges = set(edges)
    if short:
        for _ in range(max(2, length // 60)):
            p = int(rng.integers(0, length - 9)); edges.add((p, p + 2 + int(rng.integers(0, 6))))
    if longs:
        for _ in range(max(2, length // 120)):
            q = int(rng.integers(10, length))
            for _ in range(1 + (int(rng.integers(0, 2)) if multi else 0)):
                p = int(rng.integers(max(0, q - 400), q - 8)); edges.add((p, q))
    for _ in range(srcs):
        a = len(labels); labels.append("ACGT"[int(rng.integers(0, 4))]); edges.add((a, int(rng.integers(1, length))))
    edges = [(int(a), int(b)) for a, b in edges]; rng.shuffle(edges)
    return "".join(labels), edges
def run(tag, which, **kw):
    rng = np.random.default_rng(99)
    pairs = []
    for k in range(16):
        sides = []
        for side in range(2):
            if which in (side, 2): labels, edges = chain(rng, int(rng.integers(500, 1200)), **kw)
            else: labels, edges = chain(rng, int(rng.integers(500, 1200)), 0.0, 0.0, False, False, False, 0)
            src, snk = sources_and_sinks(len(labels), edges)
            sides.append(graph_from_edges(labels, edges, src, snk))
        pairs.append(tuple(sides))
    b = batch_from_graph_pairs(pairs)
    scores, alns = po_poa_batch(b, PROD)
    bad = sum(1 for w in range(b.n_windows) if oracle.po_poa(b, w, PROD)[0] != scores[w])
    print(f"{tag} side={which}: {bad} bad scores of {b.n_windows}")
for which in (0, 1):
    run("plain", which, snp=0.0, dele=0.0, short=False, longs=False, multi=False, srcs=0)
    run("snp", which, snp=0.08, dele=0.0, short=False, longs=False, multi=False, srcs=0)
    run("del", which, snp=0.0, dele=0.04, short=False, longs=False, multi=False, srcs=0)
    run("short", which, snp=0.0, dele=0.0, short=True, longs=False, multi=False, srcs=0)
    run("long", which, snp=0.0, dele=0.0, short=False, longs=True, multi=False, srcs=0)
    run("multi", which, snp=0.0, dele=0.0, short=False, longs=True, multi=True, srcs=0)
    run("srcs", which, snp=0.0, dele=0.0, short=False, longs=False, multi=False, srcs=2)
