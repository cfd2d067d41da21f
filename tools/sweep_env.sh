#!/bin/bash
# Sensitivity of the gap-fill kernel to its launch knobs on the 8 000-window sample (GPU box): bash tools/sweep_env.sh
B="timeout 300 python bench.py --windows ${WINDOWS:-8000} --no-cpu-baseline --no-e2e --no-other-paths"
f(){ tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1))"; }
echo -n "default:            "; $B 2>&1 | f
for p in 1024 4096 auto; do echo -n "CLB_PANEL_ROWS=$p: "; CLB_PANEL_ROWS=$p $B 2>&1 | f; done
for l in 32 128 256; do echo -n "CLB_START_LAG=$l:   "; CLB_START_LAG=$l $B 2>&1 | f; done
