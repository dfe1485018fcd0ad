timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for x in 0 1; do
GDMCF_PROJECTED_LOOP=$x python bench.py --steps 30 --no_cpu_baseline --configs none 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('projected $x ms',d['ms_per_step'],'frac',d['roofline']['frac'],'gemm_ms',d['roofline']['gemm_ms_per_step'],'flops',d['roofline']['flops_per_step'])"
GDMCF_PROJECTED_LOOP=$x python bench.py --steps 30 --no_cpu_baseline --configs none --mode rank 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('projected $x rank ms',d['ms_per_step'])"
done
