timeout 300 python -m pytest tests/test_popoa_gpu.py -x -q -m gpu 2>&1 | tail -2
B="timeout 200 python bench.py --windows 8000 --no-cpu-baseline --no-e2e --no-other-paths"
S="timeout 200 python bench.py --windows 1480 --len-min 3000 --len-max 4000 --no-cpu-baseline --no-e2e --no-other-paths"
f(){ tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d[\"value\"],1))"; }
echo -n "main auto/800: "; $B 2>&1 | f
echo -n "auto/200: "; CLB_LIBRARY=$PWD/centrolign_b200/csrc/libclb_p200.so $B 2>&1 | f
echo -n "fixed2048/800: "; CLB_PANEL_ROWS=2048 $B 2>&1 | f
echo -n "small: main auto/800: "; $S 2>&1 | f
echo -n "small: fixed2048/800: "; CLB_PANEL_ROWS=2048 $S 2>&1 | f
echo -n "small: auto/200: "; CLB_LIBRARY=$PWD/centrolign_b200/csrc/libclb_p200.so $S 2>&1 | f
