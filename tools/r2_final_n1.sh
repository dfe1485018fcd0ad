# round-2 final single-GPU pass: tests, default bench line, reference arm, one ncu --set full capture of the three heaviest contractions
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_r2_default_n1.json 2> gpurun_out/bench_r2_default_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2_reference_arm.json 2> gpurun_out/bench_r2_reference_arm.err; echo ref rc=$?
python tools/r2_ncu_cases.py > gpurun_out/r2_ncu_cases_plain.txt 2>&1 && ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_bf16_tn_2cta -c 6 -f -o gpurun_out/r2_gemm python tools/r2_ncu_cases.py > gpurun_out/r2_ncu_cases_ncu.txt 2>&1; echo ncu rc=$?
cat gpurun_out/r2_ncu_cases_plain.txt | tail -3
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2_default_n1.json').read())
print('ms',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'cpu',d['cpu_baseline']['value'],'parity',d['parity'])
for c in d['configs']: print(c['workload'],c['mode'],c['steps'],round(c['ms_per_step'],3),round(c['users_per_s']))
print('spmm',d['spmm']['ms'],d['spmm'].get('bf16_mode'))
r=json.loads(open('gpurun_out/bench_r2_reference_arm.json').read()); print('ref',r['value'],r['ms_per_step'])
"
