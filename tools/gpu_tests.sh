#!/bin/bash
# GPU parity pass only: bash tools/gpu_tests.sh <tag> [pytest args]
TAG=${1:-t}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider "$@" > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -60 $OUT/pytest_gpu.log
