#!/bin/bash
TAG=${1:-mma11}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 600 python -m pytest tests/test_gpu_mma.py tests/test_gpu_engine.py -m gpu -q --maxfail=12 -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
timeout 300 python tools/gemv_trace.py > $OUT/trace.log 2>&1; grep -E "==|L3|whole" $OUT/trace.log
timeout 300 python tools/gemv_bench.py --pdl --mma --only c2 --json $OUT/gemv.json > $OUT/gemv.log 2>&1; cat $OUT/gemv.log | tail -7
( timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cut -c1-160 $OUT/bench_c2.json; tail -2 $OUT/bench_c2.err
( timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --workload c1 ) > $OUT/bench_c1.json 2> $OUT/bench_c1.err; cut -c1-160 $OUT/bench_c1.json; tail -2 $OUT/bench_c1.err
( ZB_MMA_Q4_0_ALL=1 timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --workload c1 ) > $OUT/bench_c1_all.json 2> $OUT/bench_c1_all.err; cut -c1-160 $OUT/bench_c1_all.json; tail -2 $OUT/bench_c1_all.err
