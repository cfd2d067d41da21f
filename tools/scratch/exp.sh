CLB_PANEL_ROWS=256 timeout 200 python tools/scratch/dbg2.py 2>&1 | tail -14
echo "== tests panel default"; timeout 300 python -m pytest tests/test_popoa_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== tests panel 256"; CLB_PANEL_ROWS=256 timeout 300 python -m pytest tests/test_popoa_gpu.py -x -q -m gpu 2>&1 | tail -3
