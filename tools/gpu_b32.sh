#!/bin/bash
TAG=${1:-b32}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_batch.py -m gpu -q --maxfail=5 -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
( timeout 900 python bench.py --steps 64 --warmup 4 --workload c3 --batch 32 ) > $OUT/bench_c3_b32.json 2> $OUT/bench_c3_b32.err; cut -c1-420 $OUT/bench_c3_b32.json; tail -2 $OUT/bench_c3_b32.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 330 --csv --log-file $OUT/launches_c3_b32.csv \
    python bench.py --steps 2 --warmup 3 --workload c3 --batch 32 > $OUT/ncu_launches.log 2>&1
