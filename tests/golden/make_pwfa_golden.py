"""Generate tests/golden/pwfa_golden.npz by running the UNMODIFIED reference's pwfa_po_poa
(oracle/_ref/libclref.so :: clref_pwfa_po_poa_succ, built by oracle/Makefile from /root/reference
sources) on the windows of tests/golden/popoa_golden.npz plus a few larger HOR-like windows, in
successor form, for several prune limits and with shuffled next() orders.
Run in the build container only:  python tests/golden/make_pwfa_golden.py

Stored per case: window index into the stored batch, parameter set, prune limit, the reference's score
and alignment.  tests/test_oracle.py (CPU) and tests/test_pwfa_gpu.py (GPU) replay them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from centrolign_b200.batch import (AlignmentParameters, concat_batches, successor_form,  # noqa: E402
                                   synth_windows)
from golden_io import load_golden  # noqa: E402

PRUNE_LIMITS = [0, 3, 50, 10 ** 6]  # Stitcher passes 2*wfa_pruning_dist = 50 (stitcher.hpp:339)


def main():
    from checkers import CpuChecker
    ref = CpuChecker("reference")
    base, params, pidx, _, _ = load_golden()
    hor = synth_windows(10, first_index=40, seed=11, len_min=150, len_max=2500, alt_len=61, alt_period=500)
    prod = AlignmentParameters()
    batch = concat_batches([base, hor])
    param_sets = list(params) + [prod]
    pidx = np.concatenate([pidx, np.full(hor.n_windows, len(param_sets) - 1, np.int32)])
    rng = np.random.default_rng(20261017)
    sb = concat_batches([successor_form(base), successor_form(hor)])
    sb_shuf = successor_form(batch, rng)
    cases, scores, alns = [], [], []
    for variant, b in enumerate((sb, sb_shuf)):
        for w in range(batch.n_windows):
            p = param_sets[pidx[w]]
            if 0 in p.gap_open:  # the reference's gcd divides by zero (alignment.hpp:1617-1628)
                continue
            for lim in (PRUNE_LIMITS if variant == 0 else PRUNE_LIMITS[1:3]):
                s, a = ref.pwfa_po_poa(b, w, p, lim)
                cases.append((variant, w, pidx[w], lim))
                scores.append(s)
                alns.append(a)
    aln_off = np.zeros(len(alns) + 1, np.int64)
    np.cumsum([len(a) for a in alns], out=aln_off[1:])
    out = {"param_sets": np.stack([p.packed() for p in param_sets]),
           "param_num_pw": np.asarray([p.num_pw for p in param_sets], np.int32),
           "cases": np.asarray(cases, np.int64), "score": np.asarray(scores, np.int64),
           "aln": np.concatenate(alns).astype(np.int32), "aln_off": aln_off}
    for vname, b in (("v0", sb), ("v1", sb_shuf)):
        for name, side in (("g1", b.g1), ("g2", b.g2)):
            for f in ("node_off", "label", "edge_off", "pred_off", "pred", "src_off", "src", "snk_off", "snk"):
                out[f"{vname}_{name}_{f}"] = getattr(side, f)
    path = os.path.join(HERE, "pwfa_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(cases)} cases over {batch.n_windows} windows, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
