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
    # the wavefront route: min_wfa_size lowered through a config file so that every window above 100 cells with
    # similar side lengths goes to pwfa_po_poa (stitcher.hpp:327-339); the output differs from the default route's
    ("pair60k_wfa_route", [2, 60000, 3, 2], ["-a", "100000"], {"min_wfa_size": 100}),
    ("msa3_40k_wfa_route", [3, 40000, 9, 1], ["-a", "100000"], {"min_wfa_size": 100}),
    # cyclizing mode (-c, configs[4]): tandem-duplication detection with min_cyclizing_length lowered so that the
    # HOR indels of the small input qualify; exercises Stitcher::internal_stitch and the bond alignment path
    ("msa3_40k_cyclic", [3, 40000, 9, 3], ["-c", "-a", "100000"], {"min_cyclizing_length": 1000}),
    # round 2 -----------------------------------------------------------------------------------------------------
    # restrain_memory: memory_restraint_size 0 makes Core::align ask for the bit-packed chaining containers in every
    # merge (core.hpp:194, anchorer.hpp:1258-1270): the packed instantiation of the chaining DP on the GPU
    ("pair20k_restrain_memory", [2, 20000, 7, 0], ["-a", "60000"], {"memory_restraint_size": 0}),
    ("msa3_40k_restrain_memory", [3, 40000, 9, 1], ["-a", "100000"], {"memory_restraint_size": 0}),
    # default options (no anchor cap) with a Newick guide tree: scaled configs[2]
    ("tree4_3k_default", [4, 3000, 5, 1], [], {}, "((seq0,seq1),(seq2,seq3));"),
    # a cyclizing case small enough for every GPU test run (scaled configs[4])
    ("msa3_2k5_cyclic", [3, 2500, 9, 2], ["-c", "-y", "800"], {}),
    # configs[0] at full size with default options
    ("pair100k_default", [2, 100000, 1, 0], [], {}),
]


def run_cli(cli, opts, fa, overrides, tmp, env=None, tree=None):
    """Run the CLI; with config overrides, go through --generate-config / --config (src/main.cpp:137-193)."""
    if tree:
        nwk = os.path.join(tmp, "guide.nwk")
        with open(nwk, "w") as f:
            f.write(tree + "\n")
        opts = list(opts) + ["-T", nwk]
    if not overrides:
        return subprocess.run([cli, "-v", "0"] + opts + [fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    cfg = subprocess.run([cli] + opts + ["-G", fa], stdout=subprocess.PIPE, check=True).stdout.decode().splitlines()
    overrides = dict(overrides, logging_level=0)
    out = []
    for line in cfg:
        key = line.strip().split(":")[0]
        out.append(f" {key}: {overrides[key]}" if key in overrides and not line.strip().startswith("#") else line)
    path = os.path.join(tmp, "config.yaml")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    return subprocess.run([cli, "-C", path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)


def make_fasta(path, args):
    subprocess.run([sys.executable, os.path.join(ROOT, "integration", "make_hor_fasta.py"), path] + [str(a) for a in args], check=True)


def md5(data: bytes) -> str:
    return hashlib.md5(data).hexdigest()


def main():
    ref = os.path.join(ROOT, "oracle", "_ref", "centrolign_ref")
    path = os.path.join(ROOT, "tests", "golden", "e2e.json")
    out = json.load(open(path)) if os.path.exists(path) and "--all" not in sys.argv else {}
    with tempfile.TemporaryDirectory() as tmp:
        for case in CASES:
            name, fa_args, opts = case[:3]
            if name in out:  # the reference takes minutes per case: only new cases are generated unless --all is given
                continue
            overrides = case[3] if len(case) > 3 else {}
            tree = case[4] if len(case) > 4 else None
            fa = os.path.join(tmp, name + ".fa")
            make_fasta(fa, fa_args)
            t0 = time.time()
            res = run_cli(ref, opts, fa, overrides, tmp, tree=tree)
            assert res.returncode == 0, res.stderr.decode()[-2000:]
            out[name] = {"fasta_args": fa_args, "options": opts, "config_overrides": overrides, "tree": tree,
                         "fasta_md5": md5(open(fa, "rb").read()),
                         "output_md5": md5(res.stdout), "output_bytes": len(res.stdout),
                         "reference_seconds": round(time.time() - t0, 1), "head": res.stdout[:120].decode()}
            print(name, out[name]["output_md5"], out[name]["output_bytes"], out[name]["reference_seconds"])
    with open(os.path.join(ROOT, "tests", "golden", "e2e.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
