#!/bin/bash
TAG=${1:-mma12}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 600 python -m pytest tests/test_gpu_mma.py -m gpu -q --maxfail=12 -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
timeout 300 python tools/gemv_trace.py > $OUT/trace.log 2>&1; grep -E "==|L3|whole" $OUT/trace.log
timeout 300 python tools/gemv_bench.py --pdl --mma --only c2 --json $OUT/gemv.json > $OUT/gemv.log 2>&1; cat $OUT/gemv.log | tail -7
ZB_MMA_CTAS=74 timeout 300 python tools/gemv_bench.py --pdl --mma --only c2 > $OUT/gemv_74.log 2>&1; cat $OUT/gemv_74.log | tail -7
( timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cut -c1-160 $OUT/bench_c2.json; tail -2 $OUT/bench_c2.err
