set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests_c.log
for x in 0 1; do
GDMCF_PDL=$x python bench.py --steps 30 --no_cpu_baseline 2>gpurun_out/r2_pdl_$x.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pdl $x ms',d['ms_per_step'],'frac',d['roofline']['frac'],'gemm_ms',d['roofline']['gemm_ms_per_step'],'spmm',d['spmm']['ms'])" >> gpurun_out/r2_pdl_sweep.txt
GDMCF_PDL=$x python bench.py --steps 30 --no_cpu_baseline --mode rank 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pdl $x rank ms',d['ms_per_step'])" >> gpurun_out/r2_pdl_sweep.txt
done
cat gpurun_out/r2_pdl_sweep.txt; tail -5 gpurun_out/r2_tests_c.log
