"""End-to-end wall time and output identity of the reference CLI, unmodified vs. built with the B200 shadow
headers (integration/Makefile: chaining, po_poa and pwfa_po_poa on the GPU), on synthetic HOR arrays.
    python tools/e2e_compare.py --seqs 2 --length 100000 [--hor-indels 0] [--options "-a 100000"]
Both binaries are oracle/_ref/centrolign_{ref,b200}; prints one JSON line."""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=2)
    ap.add_argument("--length", type=int, default=100000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--hor-indels", type=int, default=0)
    ap.add_argument("--options", default="")
    ap.add_argument("--skip-ref", action="store_true")
    a = ap.parse_args()
    out = {"config": f"{a.seqs} synthetic HOR arrays of {a.length} bp (seed {a.seed}, hor_indels {a.hor_indels})", "options": a.options}
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "x.fa")
        subprocess.run([sys.executable, os.path.join(ROOT, "integration", "make_hor_fasta.py"), fa, str(a.seqs), str(a.length),
                        str(a.seed), str(a.hor_indels)], check=True)
        env = dict(os.environ, CLB_COUNT_CALLS="1")
        for name in (["b200"] if a.skip_ref else ["b200", "ref"]):
            cli = os.path.join(ROOT, "oracle", "_ref", "centrolign_" + name)
            t0 = time.perf_counter()
            res = subprocess.run([cli, "-v", "0"] + a.options.split() + [fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
            dt = time.perf_counter() - t0
            out[name] = {"seconds": round(dt, 2), "returncode": res.returncode, "bytes": len(res.stdout),
                         "md5": hashlib.md5(res.stdout).hexdigest(),
                         "gpu_calls": [l for l in res.stderr.decode().splitlines() if l.startswith("[clb]")]}
            if res.returncode != 0:
                out[name]["stderr"] = res.stderr.decode()[-500:]
    if "ref" in out:
        out["identical"] = out["ref"]["md5"] == out["b200"]["md5"]
        out["speedup"] = round(out["ref"]["seconds"] / out["b200"]["seconds"], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
