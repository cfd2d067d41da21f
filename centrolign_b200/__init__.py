"""centrolign_b200: B200-native (sm_100a) implementation of centrolign's data-parallel hot path.

Scope: the Stitcher's inter-anchor gap fill -- the piecewise-affine
partial-order-to-partial-order graph DP ``po_poa`` -- behind the reference's own call
signature, as hand-written CUDA kernels reached through a thin C ABI
(``include/centrolign_b200.h``).  PyTorch is used only as harness plumbing.
"""
from .batch import (AlignmentParameters, GraphSide, WindowBatch, batch_from_graph_pairs,  # noqa: F401
                    graph_from_edges, synth_windows)

__all__ = ["AlignmentParameters", "GraphSide", "WindowBatch", "batch_from_graph_pairs",
           "graph_from_edges", "synth_windows"]
from .popoa import ClbError, DeviceBatch, po_poa, po_poa_batch  # noqa: E402,F401

__all__ += ["ClbError", "DeviceBatch", "po_poa", "po_poa_batch"]
