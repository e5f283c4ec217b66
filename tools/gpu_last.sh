#!/bin/bash
TAG=${1:-last}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
( timeout -s KILL 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider 2>&1 | tail -8 ) > $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 200 python tools/moe_gemv_bench.py > $OUT/moe_gemv.log 2>&1; cat $OUT/moe_gemv.log | tail -5
