#!/bin/bash
# Tensor-core GEMV: parity (Q4_K, Q6_K, Q4_0), microbench C1+C2 shapes, engine parity, C2 and C1 bench lines.
TAG=${1:-mma6}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 600 python -m pytest tests/test_gpu_mma.py -m gpu -q --maxfail=12 -p no:cacheprovider > $OUT/pytest_mma.log 2>&1; tail -12 $OUT/pytest_mma.log
timeout 300 python tools/gemv_bench.py --pdl --mma --json $OUT/gemv_mma.json > $OUT/gemv_mma.log 2>&1; cat $OUT/gemv_mma.log | tail -14
timeout -s KILL 900 python -m pytest tests/test_gpu_engine.py -m gpu -q --maxfail=5 -p no:cacheprovider > $OUT/pytest_engine.log 2>&1; tail -5 $OUT/pytest_engine.log
( timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cut -c1-200 $OUT/bench_c2.json; tail -2 $OUT/bench_c2.err
( timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --workload c1 ) > $OUT/bench_c1.json 2> $OUT/bench_c1.err; cut -c1-200 $OUT/bench_c1.json; tail -2 $OUT/bench_c1.err
( ZB_GEMV_TC=0 timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --workload c1 ) > $OUT/bench_c1_simt.json 2> $OUT/bench_c1_simt.err; cut -c1-200 $OUT/bench_c1_simt.json; tail -2 $OUT/bench_c1_simt.err
