#!/bin/bash
# 2 GPUs, seconds: the N=2 bench line on an 8-layer cut of the 70B shape (ZB_TP_NCCL_ONLY=1 for the NCCL A/B)
mkdir -p gpurun_out/r2tp2d
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ZB_BENCH_LIMIT_S=45 timeout 60 $TR --master-port 29651 bench.py --gpus 2 --steps 20 --warmup 5 --layers 8 > gpurun_out/r2tp2d/push.json 2> gpurun_out/r2tp2d/push.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2tp2d/push.json') if l.startswith('{')][-1]); print('push', d.get('value'), d.get('ms_per_step'), d.get('allreduce_us_per_step'), d.get('exchange'))"
