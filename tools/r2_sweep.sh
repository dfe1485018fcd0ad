timeout 600 python -m pytest tests/test_headline_parity_gpu.py tests/test_fullsize_gpu.py -x -q -s -k lightgcn 2>&1 | tail -12
python bench.py --steps 20 --no_cpu_baseline --configs none 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['spmm'], indent=1))"
