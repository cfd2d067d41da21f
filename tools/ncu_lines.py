#!/usr/bin/env python
"""Join an ncu report's per-SASS metrics with nvdisasm line info -> per-source-line summary.
usage: tools/ncu_lines.py report.ncu-rep [kernel-substring] [top-n]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "popoa_kernelILi3"
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "centrolign_b200", "csrc", "libcentrolign_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
seq = []
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], stdout=subprocess.PIPE, text=True).stdout
    fn, cur = None, None
    for l in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = int(m.group(2))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and fn and kern in fn:
            seq.append((cur, m.group(2).strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix, tx, smp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
data = rows[2:]
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for k in range(min(len(seq), len(data))):
    ln = seq[k][0] or -1
    e, t, s = int(data[k][ix]), int(data[k][tx]), int(data[k][smp])
    agg[ln][0] += e
    agg[ln][1] += t
    agg[ln][2] += s
    tot += e
src = open(os.path.join(root, "centrolign_b200", "csrc", "popoa_kernels.cu")).read().splitlines()
tots = sum(v[2] for v in agg.values())
print(f"sass {len(seq)} / ncu rows {len(data)}; total warp-inst {tot}")
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    text = src[ln - 1].strip()[:100] if 0 < ln <= len(src) else ""
    print("L%4d inst %5.1f%% thr %4.1f stall %5.1f%%  %s" % (ln, v[0] / tot * 100, v[1] / max(1, v[0]), v[2] / tots * 100, text))
