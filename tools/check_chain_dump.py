#!/usr/bin/env python
"""Replay the chaining problems a caller sent to clb_chain_dp (written with CLB_DUMP_DIR=dir): every problem twice on the
GPU (determinism) and once through the C oracle (oracle/chain_oracle.c); chains, DP values and back-pointers compared.
    CLB_DUMP_DIR=/tmp/d oracle/_ref/centrolign_b200 ... ; python tools/check_chain_dump.py /tmp/d"""
import glob, os, sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from centrolign_b200.chain import chain_dp, read_chain_bin  # noqa: E402
from checkers import chain_oracle  # noqa: E402


def main():
    files = sorted(glob.glob(os.path.join(sys.argv[1], "chain_*.bin")), key=lambda p: int(p.split("_")[-1].split(".")[0]))
    os.environ.pop("CLB_DUMP_DIR", None)
    bad = nondet = 0
    for path in files:
        for kind, prob in read_chain_bin(path).items():
            c1, dp1, bp1, o1 = chain_dp(prob)
            c2, dp2, bp2, o2 = chain_dp(prob)
            co, dpo, bpo, oo = chain_oracle(prob)
            same_gpu = np.array_equal(c1, c2) and np.array_equal(dp1.view(np.uint32), dp2.view(np.uint32)) and np.array_equal(bp1, bp2)
            ok = np.array_equal(c1, co) and np.array_equal(dp1.view(np.uint32), dpo.view(np.uint32)) and np.array_equal(bp1, bpo)
            if not same_gpu:
                nondet += 1
            if not ok or not same_gpu:
                bad += 1
                if bad <= 15:
                    nd = int((dp1.view(np.uint32) != dpo.view(np.uint32)).sum())
                    nb = int((bp1 != bpo).sum())
                    print(f"{os.path.basename(path)} [{kind}] matches {prob.n_match} steps {prob.n_step} chains {prob.n_chain1}x{prob.n_chain2}: "
                          f"gpu==gpu {same_gpu}; vs oracle: chain {np.array_equal(c1, co)} ({len(c1)} vs {len(co)}), dp diffs {nd}, bp diffs {nb}")
    print(f"checked {len(files)} problems: {bad} differ from the oracle or between runs ({nondet} non-deterministic)")


if __name__ == "__main__":
    main()
