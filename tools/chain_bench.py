#!/usr/bin/env python
"""Chaining DP alone: the 100 k-match pairwise HOR problem of bench.py's other_paths.chain_dp (made by oracle/_ref/chain_fixture, i.e. by the
reference's own match finder) through clb_chain_dp, checked against the chain the reference found.  Usage:
    python tools/chain_bench.py [--reps 3] [--no-warmup] [--kinds gapfree,affine] [--bp 20000] [--matches 100000]
Environment knobs of chain_host.cu (CLB_CHAIN_GRID, CLB_CHAIN_CLUSTER, ...) apply.  One line per kind: kernel ms (best of reps), us per step."""
import argparse, os, subprocess, sys, tempfile, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-warmup", action="store_true")
    ap.add_argument("--kinds", default="gapfree,affine")
    ap.add_argument("--bp", type=int, default=20000)
    ap.add_argument("--matches", type=int, default=100000)
    ap.add_argument("--nseq", type=int, default=2)
    a = ap.parse_args()
    from centrolign_b200.chain import ChainStats, chain_dp, read_chain_bin
    shim = os.path.join(ROOT, "oracle", "_ref", "chain_fixture")
    with tempfile.TemporaryDirectory() as tmp:
        fa, binp = os.path.join(tmp, "x.fa"), os.path.join(tmp, "x.bin")
        subprocess.run([sys.executable, os.path.join(ROOT, "integration", "make_hor_fasta.py"), fa, str(a.nseq), str(a.bp), "7", "0"], check=True)
        subprocess.run([shim, fa, binp, "pair", str(a.matches), "1.0"], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
        probs = read_chain_bin(binp)
    for kind in a.kinds.split(","):
        prob = probs[kind]
        if not a.no_warmup:
            chain_dp(prob, device=0)
        best = None
        for _ in range(1 if a.no_warmup else a.reps):
            st = ChainStats()
            t0 = time.perf_counter()
            chain, dp, bp, opt = chain_dp(prob, device=0, stats=st)
            wall = (time.perf_counter() - t0) * 1e3
            ok = np.array_equal(chain, prob.expect_chain)
            if best is None or st.kernel_ms < best[0]:
                best = (st.kernel_ms, st.build_ms, wall, ok)
        print(f"{kind}: matches {prob.n_match} steps {prob.n_step} kernel {best[0]:.2f} ms ({best[0] * 1e3 / max(1, prob.n_step):.2f} us/step) "
              f"build {best[1]:.1f} ms wall {best[2]:.1f} ms reference {prob.ref_ms:.1f} ms equal_to_reference {best[3]}", flush=True)


if __name__ == "__main__":
    main()
