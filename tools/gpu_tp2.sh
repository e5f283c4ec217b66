#!/bin/bash
# 2 GPUs: the driver's N=2 bench line -- first on an 8-layer cut of the 70B shape (plumbing check, seconds), then at full depth
# with the bench's own watchdog at 4 minutes; progress markers land in the .err files
mkdir -p gpurun_out/r2tp2b
free -g | head -2 > gpurun_out/r2tp2b/box.txt; nproc >> gpurun_out/r2tp2b/box.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ZB_BENCH_LIMIT_S=100 timeout 130 $TR --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 --layers 8 > gpurun_out/r2tp2b/n2_l8.json 2> gpurun_out/r2tp2b/n2_l8.err
grep "bench +" gpurun_out/r2tp2b/n2_l8.err | tail -12; head -c 600 gpurun_out/r2tp2b/n2_l8.json; echo
ZB_BENCH_LIMIT_S=230 timeout 260 $TR --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2tp2b/n2_full.json 2> gpurun_out/r2tp2b/n2_full.err
grep "bench +" gpurun_out/r2tp2b/n2_full.err | tail -14; head -c 1500 gpurun_out/r2tp2b/n2_full.json; echo
