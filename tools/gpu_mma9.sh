#!/bin/bash
TAG=${1:-mma9}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 600 python -m pytest tests/test_gpu_mma.py -m gpu -q --maxfail=12 -p no:cacheprovider > $OUT/pytest_mma.log 2>&1; tail -8 $OUT/pytest_mma.log
timeout 300 python tools/gemv_bench.py --pdl --mma --json $OUT/gemv_i8.json > $OUT/gemv_i8.log 2>&1; cat $OUT/gemv_i8.log | tail -12
timeout -s KILL 900 python -m pytest tests/test_gpu_engine.py -m gpu -q --maxfail=5 -p no:cacheprovider > $OUT/pytest_engine.log 2>&1; tail -5 $OUT/pytest_engine.log
for c in 32 64 160; do ( ZB_ATTN_CHUNK=$c timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2_chunk$c.json 2> $OUT/bench_c2_chunk$c.err; echo "chunk $c: $(cut -c1-120 $OUT/bench_c2_chunk$c.json)"; tail -2 $OUT/bench_c2_chunk$c.err; done
( timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --workload c1 ) > $OUT/bench_c1.json 2> $OUT/bench_c1.err; cut -c1-120 $OUT/bench_c1.json; tail -2 $OUT/bench_c1.err
( ZB_ATTN_CHUNK=96 timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --workload c1 ) > $OUT/bench_c1_chunk96.json 2> $OUT/bench_c1_chunk96.err; cut -c1-120 $OUT/bench_c1_chunk96.json; tail -2 $OUT/bench_c1_chunk96.err
