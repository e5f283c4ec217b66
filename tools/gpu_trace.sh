#!/bin/bash
TAG=${1:-trace}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python tools/gemv_trace.py > $OUT/trace.log 2>&1; cat $OUT/trace.log | tail -60
timeout -s KILL 600 python -m pytest tests/test_gpu_mma.py -m gpu -q --maxfail=12 -p no:cacheprovider -k "Q6_K or q6k" > $OUT/pytest_q6k.log 2>&1; tail -15 $OUT/pytest_q6k.log
timeout 300 python tools/gemv_bench.py --pdl --only c2 --mma --json $OUT/gemv_mma.json > $OUT/gemv_mma.log 2>&1; cat $OUT/gemv_mma.log | tail -8
