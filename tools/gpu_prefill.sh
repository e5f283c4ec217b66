#!/bin/bash
# chunked prefill: parity tests, then the 4k-token C3 prompt timing
TAG=${1:-pf}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 600 python -m pytest tests/test_gpu_prefill.py -m gpu -q --maxfail=20 -p no:cacheprovider > $OUT/pytest_prefill.log 2>&1; tail -30 $OUT/pytest_prefill.log
timeout -s KILL 600 python bench.py --workload c3 --layers ${LAYERS:-8} --prefill 4096 --steps 3 --warmup 1 > $OUT/bench_prefill.json 2> $OUT/bench_prefill.err; cat $OUT/bench_prefill.json; tail -5 $OUT/bench_prefill.err
