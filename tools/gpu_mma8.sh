#!/bin/bash
TAG=${1:-mma8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 600 python -m pytest tests/test_gpu_mma.py -m gpu -q --maxfail=12 -p no:cacheprovider > $OUT/pytest_mma.log 2>&1; tail -12 $OUT/pytest_mma.log
ZB_MMA_I8=0 timeout -s KILL 600 python -m pytest tests/test_gpu_mma.py -m gpu -q --maxfail=12 -p no:cacheprovider > $OUT/pytest_mma_f16.log 2>&1; tail -4 $OUT/pytest_mma_f16.log
timeout 300 python tools/gemv_bench.py --pdl --mma --only c2 --json $OUT/gemv_i8.json > $OUT/gemv_i8.log 2>&1; cat $OUT/gemv_i8.log | tail -8
timeout -s KILL 900 python -m pytest tests/test_gpu_engine.py -m gpu -q --maxfail=5 -p no:cacheprovider > $OUT/pytest_engine.log 2>&1; tail -5 $OUT/pytest_engine.log
( timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cut -c1-200 $OUT/bench_c2.json; tail -2 $OUT/bench_c2.err
