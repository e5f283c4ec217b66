#!/bin/bash
TAG=${1:-mma}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
ZB_GEMV_MMA=1 timeout -s KILL 600 python -m pytest tests/test_gpu_stream.py -m gpu -q --maxfail=8 -p no:cacheprovider -k "Q4_K or Q5_K" > $OUT/pytest_mma.log 2>&1; tail -12 $OUT/pytest_mma.log
ZB_GEMV_MMA=1 timeout 300 python tools/gemv_bench.py --pdl --only c2 --json $OUT/gemv_mma.json > $OUT/gemv_mma.log 2>&1; cat $OUT/gemv_mma.log
ZB_GEMV_MMA=1 timeout 300 python tools/gemv_bench.py --pdl --only c3 --json $OUT/gemv_mma3.json > $OUT/gemv_mma3.log 2>&1; cat $OUT/gemv_mma3.log
