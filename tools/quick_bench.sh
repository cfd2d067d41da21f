#!/bin/bash
# Short gap-fill measurement on a B200 box: parity tests of the po_poa path, then GCUPS on an 8 000-window sample
# of configs[1] (as specified / generic step forced / no bubbles).  usage: gpurun -- 'bash tools/quick_bench.sh [notest]'
B="timeout 300 python bench.py --windows ${WINDOWS:-8000} --no-cpu-baseline --no-e2e --no-other-paths"
f(){ tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['roofline']['frac'],4))"; }
if [ "$1" != "notest" ]; then python -m pytest tests/test_popoa_gpu.py tests/test_hostcpp.py -m gpu -x -q 2>&1 | tail -6; fi
echo -n "configs[1] sample:   "; $B 2>&1 | f
echo -n "generic step forced: "; CLB_DEBUG_FLAGS=2 $B 2>&1 | f
echo -n "no bubbles at all:   "; $B --snp-rate 0 --alt-period 0 2>&1 | f
