timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -k torchrun 2>&1 | tail -5
for n in 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 30 --warmup 5 2>gpurun_out/r2_n$n.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n',d['n_gpus'],'ms',d['ms_per_step'],'value',d['value'],'frac',d['roofline']['frac'])"
done
tail -3 gpurun_out/r2_n2.err
