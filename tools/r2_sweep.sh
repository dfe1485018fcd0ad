set -x
python -m pytest tests/test_engine_gpu.py tests/test_models_gpu.py tests/test_main_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r2_tests_b.log
for x in 0 40 56 72; do
python bench.py --steps 30 --no_cpu_baseline --overlap_sms $x 2>gpurun_out/r2_ov_$x.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('overlap',d['config']['overlap_sms'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'])" >> gpurun_out/r2_overlap_sweep.txt
done
cat gpurun_out/r2_overlap_sweep.txt; tail -5 gpurun_out/r2_tests_b.log
