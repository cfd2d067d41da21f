#!/usr/bin/env python
"""End-to-end fixtures: run the UNMODIFIED reference CLI (oracle/_ref/centrolign_ref, built by
integration/Makefile from /root/reference sources) on seeded synthetic HOR arrays and record the md5 of
its CIGAR / GFA output in tests/golden/e2e.json.  tests/test_e2e_cli_gpu.py then runs the same CLI with
the Stitcher's po_poa redirected to the GPU (oracle/_ref/centrolign_b200) and demands identical bytes."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    # name, make_hor_fasta args (n_seqs, length, seed, hor_indels), centrolign options
    ("pair20k", [2, 20000, 7, 0], ["-a", "60000"]),
    ("pair60k_hor_indels", [2, 60000, 3, 2], ["-a", "100000"]),
    ("msa3_40k", [3, 40000, 9, 1], ["-a", "100000"]),
]


def make_fasta(path, args):
    subprocess.run([sys.executable, os.path.join(ROOT, "integration", "make_hor_fasta.py"), path] + [str(a) for a in args], check=True)


def md5(data: bytes) -> str:
    return hashlib.md5(data).hexdigest()


def main():
    ref = os.path.join(ROOT, "oracle", "_ref", "centrolign_ref")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, fa_args, opts in CASES:
            fa = os.path.join(tmp, name + ".fa")
            make_fasta(fa, fa_args)
            t0 = time.time()
            res = subprocess.run([ref, "-v", "0"] + opts + [fa], stdout=subprocess.PIPE, check=True)
            out[name] = {"fasta_args": fa_args, "options": opts, "fasta_md5": md5(open(fa, "rb").read()),
                         "output_md5": md5(res.stdout), "output_bytes": len(res.stdout),
                         "reference_seconds": round(time.time() - t0, 1), "head": res.stdout[:120].decode()}
            print(name, out[name]["output_md5"], out[name]["output_bytes"], out[name]["reference_seconds"])
    with open(os.path.join(ROOT, "tests", "golden", "e2e.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
