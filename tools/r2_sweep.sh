( time python bench.py > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err ) 2> gpurun_out/r2_bench_b.time
tail -3 gpurun_out/r2_bench_b.time; tail -3 gpurun_out/r2_bench_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_b.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'spmm',d['spmm']['frac'])
print('parity',{k:v for k,v in d['parity'].items() if k not in('shape','oracle')})
for c in d['configs']: print(c)
PY
