"""Window batches: the flat, concatenated layout that crosses the C ABI.

One *window* is one inter-anchor gap-fill problem: a pair of partial-order (PO) graphs
with source / sink node sets, exactly the argument list of the reference's
``po_poa(graph1, graph2, sources1, sources2, sinks1, sinks2, params)``
(reference: include/centrolign/alignment.hpp:78-85).  A *batch* concatenates many windows
side by side (struct-of-arrays), which is what ``clb_popoa_batch`` in
``include/centrolign_b200.h`` consumes.

Node ids are window-relative and in the caller's (arbitrary) order; predecessor lists are
in the graph's ``previous()`` order (include/centrolign/graph.hpp:111) because the
reference's traceback tie-breaking depends on it.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")

# production parameters of the reference's Stitcher (src/stitcher.cpp:13-22)
PRODUCTION_PARAMS = (20, 80, (60, 800, 2500), (30, 5, 1))


@dataclass
class AlignmentParameters:
    """Mirror of ``AlignmentParameters<NumPW>`` (include/centrolign/alignment.hpp:56-65)."""

    match: int = 20
    mismatch: int = 80
    gap_open: Tuple[int, ...] = (60, 800, 2500)
    gap_extend: Tuple[int, ...] = (30, 5, 1)

    @property
    def num_pw(self) -> int:
        return len(self.gap_open)

    def packed(self) -> np.ndarray:
        """``[match, mismatch, open0..2, extend0..2]`` as uint32 (unused pieces zero)."""
        if not (1 <= self.num_pw <= 3) or len(self.gap_extend) != self.num_pw:
            raise ValueError("1..3 gap pieces with matching open/extend lists required")
        out = np.zeros(8, dtype=np.uint32)
        out[0], out[1] = self.match, self.mismatch
        out[2 : 2 + self.num_pw] = self.gap_open
        out[5 : 5 + self.num_pw] = self.gap_extend
        return out

    def truncated(self, num_pw: int) -> "AlignmentParameters":
        """``truncate_parameters<N,T>`` (include/centrolign/alignment.hpp:208-219)."""
        return AlignmentParameters(self.match, self.mismatch, tuple(self.gap_open[:num_pw]),
                                   tuple(self.gap_extend[:num_pw]))


@dataclass
class GraphSide:
    """One side (graph1 or graph2) of every window of a batch, concatenated."""

    node_off: np.ndarray  # int64 [nw+1]
    label: np.ndarray  # uint8 [N]
    edge_off: np.ndarray  # int64 [nw+1]
    pred_off: np.ndarray  # uint32 [N+nw]; window w's n_w+1 entries start at node_off[w]+w
    pred: np.ndarray  # uint32 [E]
    src_off: np.ndarray  # int64 [nw+1]
    src: np.ndarray  # uint32
    snk_off: np.ndarray  # int64 [nw+1]
    snk: np.ndarray  # uint32

    def n(self, w: int) -> int:
        return int(self.node_off[w + 1] - self.node_off[w])

    def window(self, w: int):
        """(label, pred_off, pred, sources, sinks) views of window ``w``."""
        n0, n1 = int(self.node_off[w]), int(self.node_off[w + 1])
        e0, e1 = int(self.edge_off[w]), int(self.edge_off[w + 1])
        return (self.label[n0:n1], self.pred_off[n0 + w : n1 + w + 1], self.pred[e0:e1],
                self.src[int(self.src_off[w]) : int(self.src_off[w + 1])],
                self.snk[int(self.snk_off[w]) : int(self.snk_off[w + 1])])


@dataclass
class WindowBatch:
    g1: GraphSide
    g2: GraphSide

    @property
    def n_windows(self) -> int:
        return len(self.g1.node_off) - 1

    def sizes(self) -> Tuple[np.ndarray, np.ndarray]:
        return np.diff(self.g1.node_off), np.diff(self.g2.node_off)

    def cells(self) -> np.ndarray:
        """DP matrix size per window, ``(n1+1)*(n2+1)`` (include/centrolign/stitcher.hpp:241)."""
        n1, n2 = self.sizes()
        return (n1 + 1) * (n2 + 1)

    def aln_capacity(self) -> np.ndarray:
        n1, n2 = self.sizes()
        return n1 + n2


Graph = Tuple[Sequence[int], Sequence[Sequence[int]], Sequence[int], Sequence[int]]
"""(labels, predecessor lists in previous() order, sources, sinks)"""


def _side_from_graphs(graphs: List[Graph]) -> GraphSide:
    nw = len(graphs)
    node_off = np.zeros(nw + 1, np.int64)
    edge_off = np.zeros(nw + 1, np.int64)
    src_off = np.zeros(nw + 1, np.int64)
    snk_off = np.zeros(nw + 1, np.int64)
    labels, pred_offs, preds, srcs, snks = [], [], [], [], []
    for w, (lab, pl, src, snk) in enumerate(graphs):
        n = len(lab)
        assert len(pl) == n
        labels.append(np.asarray([ord(c) if isinstance(c, str) else int(c) for c in lab], np.uint8))
        po = np.zeros(n + 1, np.uint32)
        for v, p in enumerate(pl):
            po[v + 1] = po[v] + len(p)
        pred_offs.append(po)
        preds.append(np.asarray([u for p in pl for u in p], np.uint32))
        srcs.append(np.asarray(list(src), np.uint32))
        snks.append(np.asarray(list(snk), np.uint32))
        node_off[w + 1] = node_off[w] + n
        edge_off[w + 1] = edge_off[w] + int(po[n])
        src_off[w + 1] = src_off[w] + len(src)
        snk_off[w + 1] = snk_off[w] + len(snk)

    def cat(xs, dt):
        return np.ascontiguousarray(np.concatenate(xs)) if xs else np.zeros(0, dt)

    return GraphSide(node_off, cat(labels, np.uint8).astype(np.uint8), edge_off, cat(pred_offs, np.uint32).astype(np.uint32),
                     cat(preds, np.uint32).astype(np.uint32), src_off, cat(srcs, np.uint32).astype(np.uint32), snk_off,
                     cat(snks, np.uint32).astype(np.uint32))


def batch_from_graph_pairs(pairs: List[Tuple[Graph, Graph]]) -> WindowBatch:
    """Build a batch from small python-level graphs (tests, fixtures)."""
    return WindowBatch(_side_from_graphs([p[0] for p in pairs]), _side_from_graphs([p[1] for p in pairs]))


def graph_from_edges(labels: str, edges: Sequence[Tuple[int, int]], sources, sinks) -> Graph:
    """Replay ``add_node`` / ``add_edge`` calls the way ``BaseGraph`` does
    (reference: src/graph.cpp ``add_edge`` appends to ``prev`` of the head node)."""
    pl: List[List[int]] = [[] for _ in labels]
    for a, b in edges:
        pl[b].append(a)
    return (labels, pl, list(sources), list(sinks))


def concat_batches(batches: List[WindowBatch]) -> WindowBatch:
    def cat_side(sides: List[GraphSide]) -> GraphSide:
        def offs(name):
            out = [np.zeros(1, np.int64)]
            base = 0
            for s in sides:
                a = getattr(s, name)
                out.append(a[1:] + base)
                base += int(a[-1])
            return np.concatenate(out)

        def cat(name):
            return np.ascontiguousarray(np.concatenate([getattr(s, name) for s in sides]))

        return GraphSide(offs("node_off"), cat("label"), offs("edge_off"), cat("pred_off"), cat("pred"),
                         offs("src_off"), cat("src"), offs("snk_off"), cat("snk"))

    return WindowBatch(cat_side([b.g1 for b in batches]), cat_side([b.g2 for b in batches]))


def select_windows(batch: WindowBatch, idx: Sequence[int]) -> WindowBatch:
    def sel(side: GraphSide) -> GraphSide:
        graphs = []
        for w in idx:
            lab, po, pr, src, snk = side.window(int(w))
            graphs.append((lab, po, pr, src, snk))
        nw = len(graphs)
        node_off = np.zeros(nw + 1, np.int64)
        edge_off = np.zeros(nw + 1, np.int64)
        src_off = np.zeros(nw + 1, np.int64)
        snk_off = np.zeros(nw + 1, np.int64)
        for k, (lab, po, pr, src, snk) in enumerate(graphs):
            node_off[k + 1] = node_off[k] + len(lab)
            edge_off[k + 1] = edge_off[k] + len(pr)
            src_off[k + 1] = src_off[k] + len(src)
            snk_off[k + 1] = snk_off[k] + len(snk)

        def cat(i, dt):
            return np.ascontiguousarray(np.concatenate([g[i] for g in graphs])) if graphs else np.zeros(0, dt)

        return GraphSide(node_off, cat(0, np.uint8), edge_off, cat(1, np.uint32), cat(2, np.uint32), src_off,
                         cat(3, np.uint32), snk_off, cat(4, np.uint32))

    return WindowBatch(sel(batch.g1), sel(batch.g2))


def successor_form(batch: WindowBatch, rng: np.random.Generator = None) -> WindowBatch:
    """The same windows with each side's CSR holding SUCCESSOR lists in ``next()`` order (what the
    wavefront variant ``pwfa_po_poa`` enumerates, include/centrolign/alignment.hpp:1788-1826) in the
    ``pred_off`` / ``pred`` fields.  ``BaseGraph::add_edge(a, b)`` appends ``b`` to ``next(a)`` and ``a`` to
    ``previous(b)``; replaying the predecessor lists head by head (as oracle/ref_shim.cpp does) therefore
    gives ``next(a)`` = heads in ascending id order.  With ``rng`` the order inside every successor list is
    shuffled instead, to exercise the order dependence."""

    def conv(side: GraphSide) -> GraphSide:
        offs, succs = [], []
        for w in range(len(side.node_off) - 1):
            lab, po, pr, _, _ = side.window(w)
            n = len(lab)
            heads = np.repeat(np.arange(n, dtype=np.uint32), np.diff(po.astype(np.int64)))
            tails = pr.astype(np.int64)
            if rng is not None and len(tails):
                perm = rng.permutation(len(tails))
                heads, tails = heads[perm], tails[perm]
            order = np.argsort(tails, kind="stable")
            no = np.zeros(n + 1, np.uint32)
            np.cumsum(np.bincount(tails, minlength=n), out=no[1:])
            offs.append(no)
            succs.append(heads[order].astype(np.uint32))
        cat = lambda xs: np.ascontiguousarray(np.concatenate(xs)) if xs else np.zeros(0, np.uint32)  # noqa: E731
        return GraphSide(side.node_off, side.label, side.edge_off, cat(offs).astype(np.uint32), cat(succs).astype(np.uint32),
                         side.src_off, side.src, side.snk_off, side.snk)

    return WindowBatch(conv(batch.g1), conv(batch.g2))


# ----------------------------------------------------------------------------------------
# synthetic HOR-like windows (csrc/synth.c)
# ----------------------------------------------------------------------------------------
class _SynthBatch(ctypes.Structure):
    _fields_ = [("n_windows", ctypes.c_int64)] + [
        (name, ctypes.c_void_p * 2)
        for name in ("node_off", "label", "edge_off", "pred_off", "pred", "src_off", "src", "snk_off", "snk")
    ]


def _load_synth():
    path = os.path.join(_CSRC, "libclsynth.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(path)
    lib.clsynth_generate.restype = ctypes.POINTER(_SynthBatch)
    lib.clsynth_generate.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_double, ctypes.c_int64, ctypes.c_int64]
    lib.clsynth_generate_list.restype = ctypes.POINTER(_SynthBatch)
    lib.clsynth_generate_list.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_double, ctypes.c_double,
                                          ctypes.c_double, ctypes.c_int64, ctypes.c_int64]
    lib.clsynth_backbone_lengths.restype = None
    lib.clsynth_backbone_lengths.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    lib.clsynth_free.argtypes = [ctypes.POINTER(_SynthBatch)]
    return lib


def synth_windows(n_windows: int, first_index: int = 0, seed: int = 20261017, len_min: float = 880.0,
                  len_max: float = 17600.0, snp_rate: float = 0.05, alt_len: int = 171,
                  alt_period: int = 2000, indices=None) -> WindowBatch:
    """BASELINE.json configs[1] windows: backbone length log-uniform in
    [len_min, len_max] (=> ~1 k - 20 k nodes per side with bubbles), seed = window index.
    ``indices`` (int64 array) selects arbitrary windows of the stream instead of a contiguous range: a shard of one batch."""
    lib = _load_synth()
    if indices is not None:
        idx = np.ascontiguousarray(indices, np.int64)
        n_windows = int(len(idx))
        ptr = lib.clsynth_generate_list(n_windows, idx.ctypes.data, seed, len_min, len_max, snp_rate, alt_len, alt_period)
    else:
        ptr = lib.clsynth_generate(n_windows, first_index, seed, len_min, len_max, snp_rate, alt_len, alt_period)
    B = ptr.contents
    nw = n_windows

    def arr(p, count, dt):
        if count == 0:
            return np.zeros(0, dt)
        buf = (ctypes.c_char * (count * np.dtype(dt).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dt, count=count).copy()

    sides = []
    for s in range(2):
        node_off = arr(B.node_off[s], nw + 1, np.int64)
        edge_off = arr(B.edge_off[s], nw + 1, np.int64)
        src_off = arr(B.src_off[s], nw + 1, np.int64)
        snk_off = arr(B.snk_off[s], nw + 1, np.int64)
        N, E = int(node_off[-1]), int(edge_off[-1])
        sides.append(GraphSide(node_off, arr(B.label[s], N, np.uint8), edge_off, arr(B.pred_off[s], N + nw, np.uint32),
                               arr(B.pred[s], E, np.uint32), src_off, arr(B.src[s], int(src_off[-1]), np.uint32), snk_off,
                               arr(B.snk[s], int(snk_off[-1]), np.uint32)))
    lib.clsynth_free(ptr)
    return WindowBatch(sides[0], sides[1])


def synth_backbone_lengths(n_windows: int, first_index: int = 0, seed: int = 20261017, len_min: float = 880.0,
                           len_max: float = 17600.0) -> np.ndarray:
    """Backbone lengths of a range of stream windows without generating them (cell-balanced sharding of one batch)."""
    out = np.zeros(n_windows, np.int64)
    _load_synth().clsynth_backbone_lengths(n_windows, first_index, seed, len_min, len_max, out.ctypes.data)
    return out


# ----------------------------------------------------------------------------------------
# small random DAG windows for parity tests, modelled on the reference's generators
# (include/centrolign/test_util.hpp:86-131 random_graph, :133-228 random_challenge_graph)
# ----------------------------------------------------------------------------------------
def random_dag(rng: np.random.Generator, n_nodes: int, n_edges: int, alphabet: str = "ACGT") -> Tuple[str, List[Tuple[int, int]]]:
    labels = "".join(alphabet[i] for i in rng.integers(0, len(alphabet), n_nodes))
    perm = rng.permutation(n_nodes)  # hide the topological order from the node ids
    all_edges = [(a, b) for a in range(n_nodes) for b in range(a + 1, n_nodes)]
    rng.shuffle(all_edges)
    edges = [(int(perm[a]), int(perm[b])) for a, b in all_edges[:n_edges]]
    return labels, edges


def random_bubble_chain(rng: np.random.Generator, length: int, snp_rate: float = 0.15, del_rate: float = 0.05,
                        alphabet: str = "AC") -> Tuple[str, List[Tuple[int, int]]]:
    """Low-entropy backbone with SNP bubbles and deletion (skip) edges."""
    labels = [alphabet[i] for i in rng.integers(0, len(alphabet), length)]
    edges = [(i - 1, i) for i in range(1, length)]
    for p in range(1, length - 1):
        if rng.random() < snp_rate:
            a = len(labels)
            labels.append("ACGT"[int(rng.integers(0, 4))])
            edges += [(p - 1, a), (a, p + 1)]
    for p in range(length - 2):
        if rng.random() < del_rate:
            q = min(length - 1, p + 2 + int(rng.integers(0, 4)))
            if (p, q) not in edges:
                edges.append((p, q))
    rng.shuffle(edges)
    return "".join(labels), [(int(a), int(b)) for a, b in edges]


def sources_and_sinks(n: int, edges) -> Tuple[List[int], List[int]]:
    indeg = [0] * n
    outdeg = [0] * n
    for a, b in edges:
        outdeg[a] += 1
        indeg[b] += 1
    return [v for v in range(n) if indeg[v] == 0], [v for v in range(n) if outdeg[v] == 0]
