#!/bin/bash
TAG=$1; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 24 -c 4 -o $OUT/prof_gemm_tc \
    python bench.py --steps 2 --warmup 3 --workload c3 --layers 4 --batch 32 > $OUT/ncu_tc.log 2>&1
tail -3 $OUT/ncu_tc.log
