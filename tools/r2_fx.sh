# factor exchange vs reduce-scatter of the item table's gradient at N ranks: tests, then both benches
N=${1:-2}
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -k torchrun 2>&1 | tail -5
for extra in "" "--no_bf16_gather" "--no_factor_exchange --no_bf16_gather"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --configs none --no_cpu_baseline $extra 2>gpurun_out/r2_fx_n$N.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n',d['n_gpus'],'$extra','ms',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])"
tail -2 gpurun_out/r2_fx_n$N.err
done
