"""Multi-GPU sharding of window batches: one process per GPU, no data-path collective.

The windows of a batch are independent (reference: include/centrolign/stitcher.hpp:157-203 reads
only its own SubGraphInfo pair per window), so a batch is split across ranks by cell-balanced
bins and every rank runs ``clb_popoa_batch`` on its own device.  ``torch.distributed`` is used only
for plumbing: the barrier / max-over-ranks timing in bench.py and, here, returning per-window
results to rank 0 in the original order.  NCCL on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .batch import AlignmentParameters, WindowBatch, select_windows


def balanced_partition(cells: Sequence[int], n_parts: int) -> List[np.ndarray]:
    """Longest-processing-time-first assignment of windows to ``n_parts`` ranks by DP cell count
    ((n1+1)*(n2+1), include/centrolign/stitcher.hpp:241).  Returns, per rank, the window ids in
    ascending order.  Deterministic: every rank computes the same partition locally."""
    cells = np.asarray(cells, dtype=np.int64)
    order = np.argsort(-cells, kind="stable")
    load = np.zeros(n_parts, dtype=np.int64)
    parts: List[List[int]] = [[] for _ in range(n_parts)]
    for w in order:
        k = int(np.argmin(load))  # first minimum: ties go to the lowest rank
        parts[k].append(int(w))
        load[k] += int(cells[w])
    return [np.asarray(sorted(p), dtype=np.int64) for p in parts]


def stream_shard(rank: int, windows_per_rank: int) -> Tuple[int, int]:
    """Weak-scaling shard of the synthetic window stream used by bench.py: rank r owns window
    indices [r*n, (r+1)*n) of the seeded generator (csrc/synth.c)."""
    return rank * windows_per_rank, windows_per_rank


Runner = Callable[[WindowBatch, AlignmentParameters], Tuple[np.ndarray, List[np.ndarray]]]


def run_sharded(batch: WindowBatch, params: AlignmentParameters, runner: Optional[Runner] = None, device: Optional[int] = None):
    """Align ``batch`` with every rank of the default process group taking its cell-balanced
    share; rank 0 returns (scores, alignments) in the original window order, other ranks
    return ``None``.  Without an initialised process group this is a plain single-GPU call.

    ``runner`` defaults to the CUDA path (``po_poa_batch``); tests inject a CPU checker."""
    import torch.distributed as dist

    if runner is None:
        from .popoa import po_poa_batch

        def runner(b, p, _dev=device):  # noqa: E306
            import torch

            dev = _dev if _dev is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
            return po_poa_batch(b, p, device=dev)

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return runner(batch, params)
    rank, world = dist.get_rank(), dist.get_world_size()
    parts = balanced_partition(batch.cells(), world)
    mine = parts[rank]
    if len(mine):
        scores, alns = runner(select_windows(batch, mine), params)
    else:
        scores, alns = np.zeros(0, np.int64), []
    payload = (mine, np.asarray(scores), [np.asarray(a) for a in alns])
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)  # results only; the DP itself exchanges nothing
    if rank != 0:
        return None
    out_scores = np.zeros(batch.n_windows, np.int64)
    out_alns: List[Optional[np.ndarray]] = [None] * batch.n_windows
    for ids, sc, al in gathered:
        for k, w in enumerate(ids):
            out_scores[w] = sc[k]
            out_alns[w] = al[k]
    return out_scores, out_alns
