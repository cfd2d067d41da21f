#!/usr/bin/env python
"""Per-region view of an ncu source-page CSV (ncu -i rep --page source --csv > file):
usage: tools/ncu_hot.py file.csv [topN]  -- instruction share / stall-sample share by 0x800-byte code block, overall
stall reasons, and the top-N stalled instructions."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]; data = rows[2:]
ix = hdr.index("Instructions Executed"); tx = hdr.index("Thread Instructions Executed"); sx = hdr.index("# Samples")
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][0], 16)
tot = sum(int(r[ix]) for r in data); tots = sum(int(r[sx]) for r in data)
print(f"total warp-instr {tot:.4g}, samples {tots}")
agg = collections.OrderedDict()
for r in data:
    k = (int(r[0], 16) - base) // 0x800
    a = agg.setdefault(k, [0, 0, 0])
    a[0] += int(r[ix]); a[1] += int(r[tx]); a[2] += int(r[sx])
for k, v in agg.items():
    if v[0] / tot > 0.01 or v[2] / tots > 0.01:
        print(f"{k*0x800:#07x}  instr {v[0]/tot*100:5.1f}%  thr {v[1]/max(v[0],1):4.1f}  samples {v[2]/tots*100:5.1f}%")
st = collections.Counter()
for r in data:
    for c in cols: st[hdr[c]] += int(r[c] or 0)
print({k[6:]: round(v / tots * 100, 1) for k, v in st.most_common(12)})
for r in sorted(data, key=lambda r: -int(r[sx]))[:topn]:
    reasons = sorted(((int(r[c] or 0), hdr[c][6:]) for c in cols), reverse=True)[:2]
    print(f"{int(r[0],16)-base:#07x} {r[1].strip()[:58]:58s} smp {int(r[sx]):6d} exe {int(r[ix]):>10d} {reasons}")
