#!/bin/bash
TAG=${1:-tp}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
nvidia-smi topo -m > $OUT/topo.txt 2>&1
KINDS="${3:-llama_tp_q4_k_m mixtral_tp_q4_k_m llama_tp_q8_0}"
for mode in fused nccl; do
  if [ $mode = nccl ]; then export ZB_TP_NCCL_ONLY=1; else unset ZB_TP_NCCL_ONLY; fi
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py $KINDS > $OUT/tp_check_$mode.log 2>&1
  echo "rc=$?" >> $OUT/tp_check_$mode.log; echo "== $mode"; tail -4 $OUT/tp_check_$mode.log
done
