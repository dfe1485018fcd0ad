# round-2 final multi-GPU pass at N ranks: (N = 2: the torchrun tests first) then the bench line
N=$1
if [ "$N" = "2" ]; then timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_main_gpu.py -q -k torchrun 2>&1 | tail -4; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --configs none 2>gpurun_out/bench_r2_n$N.err > gpurun_out/bench_r2_n$N.json; echo rc=$?
python -c "import json; d=json.loads(open('gpurun_out/bench_r2_n$N.json').read()); print('n',d['n_gpus'],'steps',d['steps'],'ms',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'clocks',d['clocks'])"
