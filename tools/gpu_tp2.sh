#!/bin/bash
# 2 GPUs, seconds: push all-reduce vs ncclAllReduce on an 8-layer cut of the 70B shape (same bench line, ZB_TP_NCCL_ONLY=1 for the second)
mkdir -p gpurun_out/r2tp2c
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ZB_BENCH_LIMIT_S=60 timeout 80 $TR --master-port 29641 bench.py --gpus 2 --steps 20 --warmup 5 --layers 8 > gpurun_out/r2tp2c/push.json 2> gpurun_out/r2tp2c/push.err
ZB_TP_NCCL_ONLY=1 ZB_BENCH_LIMIT_S=60 timeout 80 $TR --master-port 29642 bench.py --gpus 2 --steps 20 --warmup 5 --layers 8 > gpurun_out/r2tp2c/nccl.json 2> gpurun_out/r2tp2c/nccl.err
python -c "
import json
for f in ('push','nccl'):
    d=json.loads([l for l in open('gpurun_out/r2tp2c/'+f+'.json') if l.startswith('{')][-1]); print(f, d.get('value'), d.get('ms_per_step'), d.get('allreduce_us_per_step'), d.get('exchange'))"
