timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_main_gpu.py -x -q 2>&1 | tail -12
python bench.py --steps 30 --no_cpu_baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lazy ms',d['ms_per_step'],'frac',d['roofline']['frac'],'launches',d['launches_per_step'])"
