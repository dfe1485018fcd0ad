N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --configs none 2>gpurun_out/r2_scale_n$N.err > gpurun_out/r2_scale_n$N.json
python -c "import json; d=json.loads(open('gpurun_out/r2_scale_n$N.json').read()); print('n',d['n_gpus'],'ms',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])"
