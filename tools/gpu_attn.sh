#!/bin/bash
TAG=${1:-attn}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_stream.py -m gpu -q --maxfail=5 -p no:cacheprovider -k "attn or engine or greedy or logits" > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
for cfg in "4 32" "8 32" "8 64" "16 64" "16 160"; do set -- $cfg
  ( ZB_ATTN_WARPS=$1 ZB_ATTN_CHUNK=$2 timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_w$1_c$2.json 2> $OUT/bench_w$1_c$2.err
  echo "warps $1 chunk $2: $(cut -c1-110 $OUT/bench_w$1_c$2.json)"; tail -1 $OUT/bench_w$1_c$2.err
done
( ZB_ATTN_WARPS=8 ZB_ATTN_CHUNK=32 timeout -s KILL 600 python -m pytest tests/test_gpu_engine.py -m gpu -q --maxfail=5 -p no:cacheprovider ) > $OUT/pytest_w8.log 2>&1; tail -3 $OUT/pytest_w8.log
