#!/usr/bin/env python
"""Per-source-line view of an ncu source-page CSV: joins the SASS rows of ONE kernel with nvdisasm -g line info of the same build.
usage: tools/ncu_lines.py source_page.csv nvdisasm_g.txt mangled_kernel_name [kernel_index] [topN]
  source_page.csv : ncu -i rep --page source --csv   (all captured launches, in order; kernel_index picks one)
  nvdisasm_g.txt  : nvdisasm -g the.cubin
Prints executed warp-instructions and stall samples per source line (innermost inlined location), with the main stall reasons."""
import collections, csv, re, sys

csv_path, dis_path, kname = sys.argv[1:4]
kidx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
rows = list(csv.reader(open(csv_path)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = starts[kidx]
e = starts[kidx + 1] if kidx + 1 < len(starts) else len(rows)
hdr, data = rows[s + 1], [r for r in rows[s + 2:e] if len(r) > 10]
ix, sx = hdr.index("Instructions Executed"), hdr.index("# Samples")
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
# line of every instruction, in order
lines, cur, inside = [], None, False
for ln in open(dis_path):
    if ln.startswith(".text."):
        inside = ln.strip().rstrip(":") == ".text." + kname
        continue
    if not inside:
        continue
    m = re.search(r'//## File "[^"]*", line (\d+)', ln)
    if m:
        cur = int(m.group(1))  # nested "inlined at" lines follow; keep the first (innermost)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
assert len(lines) >= len(data), (len(lines), len(data))
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = sum(int(r[ix]) for r in data)
tots = sum(int(r[sx]) for r in data)
for r, l in zip(data, lines):
    a = agg[l]
    a[0] += int(r[ix]); a[1] += int(r[sx])
    for c in cols:
        a[2][hdr[c][6:]] += int(r[c] or 0)
print(f"warp-instructions {tot:.4g}, samples {tots}")
st = collections.Counter()
for a in agg.values():
    st.update(a[2])
print({k: round(v / max(1, tots) * 100, 1) for k, v in st.most_common(10)})
for l, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"line {l}: instr {a[0] / tot * 100:5.1f}%  samples {a[1] / max(1, tots) * 100:5.1f}%  {dict(a[2].most_common(3))}")
