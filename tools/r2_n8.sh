for sms in 32 16 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 --configs none --nccl_sms $sms 2>gpurun_out/r2_n8_$sms.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n',d['n_gpus'],'nccl_sms',d['config']['nccl_sms'],'ms',d['ms_per_step'],'value',d['value'],'frac',d['roofline']['frac'])" >> gpurun_out/r2_n8_sweep.txt
done
cat gpurun_out/r2_n8_sweep.txt
