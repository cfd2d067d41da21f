"""Generate tests/golden/chain_golden.npz: chaining problems and the chains the UNMODIFIED reference finds for
them (oracle/_ref/chain_fixture, built from oracle/chain_shim.cpp against the reference's own objects).
Run in the build container only:  make -C integration && make -C oracle && python tests/golden/make_chain_golden.py

Cases: pairwise HOR arrays (two single-path graphs) and small MSA merges (2+1 and 2+2 sequences: multi-path
graphs), each as the gap-free problem (sparse_chain_dp), the global affine problem and the local affine problem
(sparse_affine_chain_dp with and without sources / sinks)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from centrolign_b200.chain import _FIELDS, read_chain_bin  # noqa: E402

CASES = [
    # name, make_hor_fasta args (n_seqs, length, seed, hor_indels), mode, max match pairs, scale
    ("pair3k", [2, 3000, 5, 0], "pair", 4000, 1.0),
    ("pair8k_indels", [2, 8000, 11, 2], "pair", 12000, 0.7),
    ("msa3_4k", [3, 4000, 9, 1], "msa", 6000, 1.0),
    ("msa4_3k", [4, 3000, 13, 1], "msa", 5000, 1.3),
    # long enough for the reference to really merge the two pairs (shorter arrays come out of its partitioner unaligned, i.e.
    # as two disjoint paths): nodes on several paths, ~4 tree entries per match
    ("msa4_24k_merged", [4, 24000, 7, 0], "msa", 12000, 1.0),
    # tiny problems, the size of the Anchorer's fill-in calls: they run from shared memory (chain_small_kernel)
    ("pair600", [2, 600, 3, 0], "pair", 300, 1.0),
    ("msa3_500", [3, 500, 4, 0], "msa", 250, 0.9),
]


def main():
    shim = os.path.join(ROOT, "oracle", "_ref", "chain_fixture")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, fa_args, mode, max_pairs, scale in CASES:
            fa, binp = os.path.join(tmp, name + ".fa"), os.path.join(tmp, name + ".bin")
            subprocess.run([sys.executable, os.path.join(ROOT, "integration", "make_hor_fasta.py"), fa] + [str(a) for a in fa_args],
                           check=True)
            res = subprocess.run([shim, fa, binp, mode, str(max_pairs), str(scale)], check=True, stdout=subprocess.PIPE, text=True)
            print(name, res.stdout.strip())
            for kind, prob in read_chain_bin(binp).items():
                pre = f"{name}/{kind}"
                out[pre + ".params"] = np.asarray([prob.num_pw, *prob.gap_open, *prob.gap_extend, prob.scale, prob.n_chain1,
                                                   prob.n_chain2, prob.ref_ms], np.float64)
                out[pre + ".min_score"] = np.asarray([prob.min_score], np.float32)
                out[pre + ".expect_chain"] = prob.expect_chain
                for k, _ in _FIELDS:
                    out[f"{pre}.{k}"] = prob.arrays[k]
    path = os.path.join(HERE, "chain_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(CASES) * 3} problems, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
