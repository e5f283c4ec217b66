#!/bin/bash
TAG=${1:-mma10}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 600 python -m pytest tests/test_gpu_mma.py tests/test_gpu_engine.py tests/test_gpu_stream.py -m gpu -q --maxfail=12 -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -8 $OUT/pytest.log
timeout 300 python tools/gemv_bench.py --pdl --mma --only c3 --json $OUT/gemv_c3.json > $OUT/gemv_c3.log 2>&1; cat $OUT/gemv_c3.log | tail -4
timeout 300 python tools/gemv_bench.py --pdl --only c3 > $OUT/gemv_c3_simt.log 2>&1; cat $OUT/gemv_c3_simt.log | tail -4
( timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cut -c1-160 $OUT/bench_c2.json; tail -2 $OUT/bench_c2.err
( ZB_ATTN_SHORT=0 timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2_long.json 2> $OUT/bench_c2_long.err; cut -c1-160 $OUT/bench_c2_long.json; tail -2 $OUT/bench_c2_long.err
( timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --workload c1 ) > $OUT/bench_c1.json 2> $OUT/bench_c1.err; cut -c1-160 $OUT/bench_c1.json; tail -2 $OUT/bench_c1.err
date +%s > $OUT/t0; ( timeout 400 python bench.py --steps 64 --warmup 8 --no-cpu --workload c3 ) > $OUT/bench_c3.json 2> $OUT/bench_c3.err; cut -c1-160 $OUT/bench_c3.json; tail -2 $OUT/bench_c3.err; date +%s > $OUT/t1; echo "c3 wall: $(( $(cat $OUT/t1) - $(cat $OUT/t0) )) s"
