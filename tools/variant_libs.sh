#!/bin/bash
# Compare alternative builds of the library on the 8 000-window sample: usage (on a GPU box): bash tools/variant_libs.sh libA.so libB.so ...
B="timeout 300 python bench.py --windows ${WINDOWS:-8000} --no-cpu-baseline --no-e2e --no-other-paths"
f(){ tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['roofline']['frac'],4))"; }
for L in "$@"; do
  echo -n "$L configs[1] sample: "; CLB_LIBRARY=$PWD/centrolign_b200/csrc/$L $B 2>&1 | f
  echo -n "$L no bubbles:        "; CLB_LIBRARY=$PWD/centrolign_b200/csrc/$L $B --snp-rate 0 --alt-period 0 2>&1 | f
done
