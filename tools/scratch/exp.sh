B="timeout 200 python bench.py --windows 8000 --no-cpu-baseline --no-e2e --no-other-paths"
f(){ tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d[\"value\"],1))"; }
echo -n "main: "; $B 2>&1 | f
echo -n "E (lean x1): "; CLB_LIBRARY=$PWD/centrolign_b200/csrc/libclb_E.so $B 2>&1 | f
echo -n "main lag32: "; CLB_START_LAG=32 $B 2>&1 | f
echo -n "main lag128: "; CLB_START_LAG=128 $B 2>&1 | f
echo -n "main linear: "; $B --snp-rate 0 --alt-period 0 2>&1 | f
echo -n "main notb: "; CLB_DEBUG_FLAGS=1 $B 2>&1 | f
