#!/bin/bash
# 8-GPU box: tensor-parallel parity at 2/4/8 ranks, TP timing on the 70B shape (reduced layers), replica scaling of the headline bench.
TAG=${1:-tp8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
nvidia-smi topo -m > $OUT/topo.txt 2>&1
run() { # n, extra env, out, args...
  local n=$1; shift; local envs=$1; shift; local out=$1; shift
  env $envs timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 "$@" > $OUT/$out.log 2>&1
  echo "rc=$?" >> $OUT/$out.log; grep -E '^\{' $OUT/$out.log | tail -1 | cut -c1-700
}
echo "== parity"
run 2 X=1 tp_check_2 tools/tp_check.py llama_tp_q4_k_m mixtral_tp_q4_k_m llama_tp_q8_0
run 4 X=1 tp_check_4 tools/tp_check.py llama_tp8_q4_k_m
run 8 X=1 tp_check_8 tools/tp_check.py llama_tp8_q4_k_m
echo "== c4 (70B shape, 4 of 80 layers) TP timing"
python -c "import bench; bench.model_path('c4', layers=4)" > $OUT/gen_c4.log 2>&1
for n in 2 4 8; do
  run $n X=1 bench_c4_tp${n}_fused bench.py --tp --workload c4 --layers 4 --steps 64 --warmup 4 --gpus $n
  run $n ZB_TP_NCCL_ONLY=1 bench_c4_tp${n}_nccl bench.py --tp --workload c4 --layers 4 --steps 64 --warmup 4 --gpus $n
done
echo "== replica scaling of the headline bench (c2)"
for n in 2 8; do
  run $n X=1 bench_c2_rep$n bench.py --gpus $n --steps 64 --warmup 4 --no-cpu
done
