timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 100 python bench.py --windows 8000 --no-cpu-baseline --no-e2e --no-other-paths 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench8000', round(d['value'],1))"
