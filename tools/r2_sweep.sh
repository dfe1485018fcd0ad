timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 30 --no_cpu_baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms',d['ms_per_step'],'frac',d['roofline']['frac'],'launches',d['launches_per_step'])"
python bench.py --steps 30 --no_cpu_baseline --mode rank 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rank ms',d['ms_per_step'])"
