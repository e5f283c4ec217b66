#!/bin/bash
TAG=${1:-tp}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py > $OUT/tp_check.log 2>&1
echo "rc=$?" >> $OUT/tp_check.log; tail -30 $OUT/tp_check.log
