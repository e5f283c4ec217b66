#!/bin/bash
TAG=${1:-tc}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout -s KILL 600 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --maxfail=40 -p no:cacheprovider -x > $OUT/pytest_tc.log 2>&1; tail -40 $OUT/pytest_tc.log
