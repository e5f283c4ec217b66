#!/bin/bash
# Tensor-core GEMV iteration: parity, phase trace, microbench, engine parity + bench, one ncu full capture of the gate_up GEMV.
TAG=${1:-mma3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 600 python -m pytest tests/test_gpu_mma.py -m gpu -q --maxfail=12 -p no:cacheprovider > $OUT/pytest_mma.log 2>&1; tail -12 $OUT/pytest_mma.log
timeout 300 python tools/gemv_trace.py > $OUT/trace.log 2>&1; grep -E "==|L3|whole" $OUT/trace.log
timeout 300 python tools/gemv_bench.py --pdl --only c2 --mma --json $OUT/gemv_mma.json > $OUT/gemv_mma.log 2>&1; cat $OUT/gemv_mma.log | tail -8
timeout -s KILL 900 python -m pytest tests/test_gpu_engine.py -m gpu -q --maxfail=5 -p no:cacheprovider > $OUT/pytest_engine.log 2>&1; tail -5 $OUT/pytest_engine.log
( timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cut -c1-200 $OUT/bench_c2.json; tail -2 $OUT/bench_c2.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemv_mma_kernelILi12E -s 6 -c 2 -o $OUT/prof_mma_q4k \
    python tools/gemv_bench.py --pdl --only c2.gate_up --mma > $OUT/ncu_full.log 2>&1; tail -3 $OUT/ncu_full.log
ls -la $OUT
