#!/usr/bin/env python
"""clb_popoa_batch_multi on every visible GPU of one box: identical results to the single-device call, and the wall
time of the whole call (host flattening, H2D, kernels, D2H) for 1 .. N devices.  usage (GPU box): python tools/multi_device_check.py [windows]"""
import json, os, sys, time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from centrolign_b200.batch import AlignmentParameters, synth_windows  # noqa: E402
from centrolign_b200.popoa import po_poa_batch, po_poa_batch_multi  # noqa: E402

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
ndev = torch.cuda.device_count()
batch = synth_windows(nw, first_index=0, seed=20261017)
params = AlignmentParameters()
cells = float(batch.cells().sum())
ref_s, ref_a = po_poa_batch(batch, params, device=0)
ref_s = ref_s.copy(); ref_a = [a.copy() for a in ref_a]
out = {"windows": nw, "cells": cells, "devices_visible": ndev, "runs": []}
for n in sorted({1, 2, 4, 8, ndev}):
    if n > ndev:
        continue
    devs = list(range(n))
    po_poa_batch_multi(batch, params, devs)  # warm-up (contexts, pinned pools)
    t0 = time.perf_counter()
    s, a, parts = po_poa_batch_multi(batch, params, devs, return_parts=True)
    dt = time.perf_counter() - t0
    same = bool(np.array_equal(s, ref_s) and all(np.array_equal(x, y) for x, y in zip(a, ref_a)))
    loads = np.bincount(parts, weights=batch.cells().astype(np.float64), minlength=n)
    out["runs"].append({"devices": n, "seconds": dt, "gcups_e2e": cells / dt / 1e9, "identical_to_single_device": same,
                        "cell_share_min_max": [float(loads.min() / cells), float(loads.max() / cells)]})
    assert same
print(json.dumps(out))
