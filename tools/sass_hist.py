#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the built library (evidence for which hardware paths the kernels use):
usage: tools/sass_hist.py [lib.so] > profiles/rNN_sass_histogram.md"""
import collections, os, re, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "centrolign_b200", "csrc", "libcentrolign_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
fn, hist = None, collections.OrderedDict()
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        fn = re.sub(r"\(.*", "", fn)
        hist[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m and fn:
        hist[fn][m.group(2)] += 1
watch = ["VIADDMNMX", "VIMNMX3", "VIMNMX", "IMAD", "SHFL", "LDS", "STS", "LDGSTS", "ATOMS", "ATOMG", "RED", "UTMALDG", "UTCMMA", "LDTM", "HMMA"]
print(f"# SASS opcode histogram of {os.path.basename(lib)} (cuobjdump -sass, sm_100a)\n")
print("DPX = VIADDMNMX / VIMNMX3 (fused add-max, three-input max); LDGSTS = cp.async; UTMALDG / UTCMMA / LDTM = TMA / tcgen05 / TMEM "
      "(none: integer max-plus DP has no GEMM form and its rows never exist in memory, DESIGN.md section 4).\n")
print("| kernel | instructions | " + " | ".join(watch) + " |")
print("|---|---|" + "---|" * len(watch))
for fn, h in hist.items():
    print(f"| `{fn}` | {sum(h.values())} | " + " | ".join(str(h.get(w, 0)) for w in watch) + " |")
