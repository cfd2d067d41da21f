#!/bin/bash
# Workload-variant table of DESIGN.md section 4 (what bounds popoa_kernel): run on a B200 box, ~3 GPU-minutes.
#   gpurun --timeout 600 -- 'bash tools/variant_table.sh'
# Each line: variant, GCUPS resident, fraction of the measured INT32/DPX peak (8 000-window sample of configs[1]).
B="timeout 200 python bench.py --windows ${WINDOWS:-8000} --no-cpu-baseline --no-e2e --no-other-paths"
f(){ tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['roofline']['frac'],4))"; }
echo -n "configs[1] as specified:        "; $B 2>&1 | f
echo -n "traceback walk skipped:         "; CLB_DEBUG_FLAGS=1 $B 2>&1 | f   # needs a -DCLB_PROFILE build (CLB_LIBRARY=...), else identical to the first line
echo -n "lean step disabled:             "; CLB_DEBUG_FLAGS=2 $B 2>&1 | f
echo -n "tiling off (CLB_PANEL_ROWS=0):  "; CLB_PANEL_ROWS=0 $B 2>&1 | f
echo -n "no 171-node bubbles:            "; $B --alt-period 0 2>&1 | f
echo -n "no SNP bubbles:                 "; $B --snp-rate 0 2>&1 | f
echo -n "no bubbles at all:              "; $B --snp-rate 0 --alt-period 0 2>&1 | f
