"""Chaining DP on the GPU next to the unmodified reference on the same problem.
    python tools/bench_chain.py [--seqs 2] [--length 20000] [--mode pair|msa] [--max-pairs 200000] [--hor-indels 0]
oracle/_ref/chain_fixture (oracle/chain_shim.cpp, built where /root/reference exists; it travels with the
snapshot) makes the problem with the reference's own match finder and runs the reference's sparse_chain_dp /
sparse_affine_chain_dp single-threaded on this box's CPU; clb_chain_dp then solves the same flat problems on the
GPU and the chains are compared match for match.  Prints one JSON line per problem kind."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from centrolign_b200.chain import ChainStats, chain_dp, read_chain_bin  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=2)
    ap.add_argument("--length", type=int, default=20000)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--hor-indels", type=int, default=0)
    ap.add_argument("--mode", default="pair")
    ap.add_argument("--max-pairs", type=int, default=200000)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    shim = os.path.join(ROOT, "oracle", "_ref", "chain_fixture")
    with tempfile.TemporaryDirectory() as tmp:
        fa, binp = os.path.join(tmp, "x.fa"), os.path.join(tmp, "x.bin")
        subprocess.run([sys.executable, os.path.join(ROOT, "integration", "make_hor_fasta.py"), fa, str(a.seqs), str(a.length),
                        str(a.seed), str(a.hor_indels)], check=True)
        t0 = time.time()
        res = subprocess.run([shim, fa, binp, a.mode, str(a.max_pairs), str(a.scale)], check=True, stdout=subprocess.PIPE, text=True)
        print("#", res.stdout.strip(), f"(shim wall {time.time() - t0:.1f} s)", file=sys.stderr)
        probs = read_chain_bin(binp)
    for kind in ("gapfree", "affine", "local"):
        prob = probs[kind]
        best = None
        for _ in range(a.reps + 1):
            st = ChainStats()
            t0 = time.perf_counter()
            chain, dp, bp, opt = chain_dp(prob, stats=st)
            wall = (time.perf_counter() - t0) * 1e3
            if best is None or wall < best[0]:
                best = (wall, st.kernel_ms, st.build_ms, st)
        wall, kms, bms, st = best
        print(json.dumps({
            "path": "chain_dp", "kind": kind, "num_pw": prob.num_pw, "matches": prob.n_match, "steps": prob.n_step,
            "chains": [prob.n_chain1, prob.n_chain2], "inserts": st.inserts, "queries": st.queries,
            "gpu_kernel_ms": kms, "gpu_build_ms": bms, "gpu_wall_ms": wall, "tree_bytes": st.tree_bytes,
            "reference_cpu_ms": prob.ref_ms, "speedup_kernel": prob.ref_ms / kms, "speedup_e2e": prob.ref_ms / wall,
            "chain_len": int(len(chain)), "equal_to_reference": bool(np.array_equal(chain, prob.expect_chain)),
            "us_per_step": kms * 1e3 / max(1, prob.n_step), "meta": prob.meta}))


if __name__ == "__main__":
    main()
