#!/bin/bash
# Tensor-core GEMV bring-up: parity tests, microbench (CUDA-core vs tensor-core), engine parity, C2 bench A/B.
TAG=${1:-mma2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 600 python -m pytest tests/test_gpu_mma.py -m gpu -q --maxfail=12 -p no:cacheprovider > $OUT/pytest_mma.log 2>&1; tail -25 $OUT/pytest_mma.log
timeout 300 python tools/gemv_bench.py --pdl --only c2 --mma --json $OUT/gemv_mma.json > $OUT/gemv_mma.log 2>&1; cat $OUT/gemv_mma.log | tail -8
timeout 300 python tools/gemv_bench.py --pdl --only c2 --json $OUT/gemv_simt.json > $OUT/gemv_simt.log 2>&1; cat $OUT/gemv_simt.log | tail -8
timeout -s KILL 900 python -m pytest tests/test_gpu_engine.py -m gpu -q --maxfail=5 -p no:cacheprovider > $OUT/pytest_engine.log 2>&1; tail -8 $OUT/pytest_engine.log
( timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cut -c1-400 $OUT/bench_c2.json; tail -2 $OUT/bench_c2.err
( ZB_GEMV_TC=0 timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2_simt.json 2> $OUT/bench_c2_simt.err; cut -c1-400 $OUT/bench_c2_simt.json; tail -2 $OUT/bench_c2_simt.err
