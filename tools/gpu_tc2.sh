#!/bin/bash
# tcgen05 GEMM consumers: unit tests, batched decode, chunked prefill, then the C3 timings
TAG=${1:-tc2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_batch.py tests/test_gpu_prefill.py -m gpu -q --maxfail=20 -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -30 $OUT/pytest.log
timeout -s KILL 600 python bench.py --workload c3 --layers ${LAYERS:-8} --prefill 4096 --steps 3 --warmup 1 > $OUT/bench_prefill.json 2> $OUT/bench_prefill.err; cat $OUT/bench_prefill.json; tail -5 $OUT/bench_prefill.err
timeout -s KILL 600 python bench.py --workload c3 --layers ${LAYERS:-8} --batch 32 --steps 32 --warmup 4 --no-cpu > $OUT/bench_b32.json 2> $OUT/bench_b32.err; cat $OUT/bench_b32.json; tail -5 $OUT/bench_b32.err
