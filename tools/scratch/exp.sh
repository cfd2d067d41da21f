timeout 300 python -m pytest tests/test_popoa_gpu.py -x -q -m gpu 2>&1 | tail -3
B="timeout 200 python bench.py --windows 8000 --no-cpu-baseline --no-e2e --no-other-paths"
f(){ tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d[\"value\"],1), d[\"roofline\"][\"frac\"])"; }
for L in libcentrolign_b200.so libclb_w13.so libclb_w14.so; do echo "== $L"; CLB_LIBRARY=$PWD/centrolign_b200/csrc/$L $B 2>&1 | f; done
