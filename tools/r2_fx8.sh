# N ranks: factor exchange on (default) then off, no tests
N=${1:-8}
for extra in "" "--no_bf16_gather" "--no_factor_exchange --no_bf16_gather"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --configs none --no_cpu_baseline $extra 2>gpurun_out/r2_fx_n$N.err | tee gpurun_out/r2_fx_n$N$extra.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n',d['n_gpus'],'$extra','ms',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])"
done
