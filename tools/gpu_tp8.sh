#!/bin/bash
# 8 GPUs, seconds: the driver's N=8 bench line on an 8-layer cut of the 70B shape (sharding / exchange / plumbing check)
mkdir -p gpurun_out/r2tp8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
ZB_BENCH_LIMIT_S=70 timeout 90 $TR --master-port 29631 bench.py --gpus 8 --steps 20 --warmup 5 --layers 8 > gpurun_out/r2tp8/n8_l8.json 2> gpurun_out/r2tp8/n8_l8.err
grep "bench +" gpurun_out/r2tp8/n8_l8.err | grep "rank 0" | tail -8; head -c 700 gpurun_out/r2tp8/n8_l8.json; echo
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2tp8/n8_l8.json') if l.startswith('{')][-1]); print('tp8 l8', d.get('value'), d.get('ms_per_step'), d.get('allreduce_us_per_step'), d.get('exchange'))"
