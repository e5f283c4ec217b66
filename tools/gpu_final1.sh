#!/bin/bash
# Round-end style pass: full GPU parity suite, bench lines, ncu launch list + full captures of the tensor-core GEMV.
set -u
TAG=${1:-final1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
( timeout -s KILL 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider 2>&1 | tail -15 ) > $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 300 python tools/gemv_trace.py > $OUT/trace.log 2>&1; grep -E "==|L3|whole" $OUT/trace.log
timeout 300 python tools/gemv_bench.py --pdl --mma --json $OUT/gemv_mma.json > $OUT/gemv_mma.log 2>&1; cat $OUT/gemv_mma.log | tail -14
( timeout 600 python bench.py --steps 128 --warmup 8 ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cut -c1-160 $OUT/bench_c2.json; tail -2 $OUT/bench_c2.err
( timeout 600 python bench.py --impl reference --steps 16 --warmup 2 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-200 $OUT/bench_ref.json; tail -2 $OUT/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 340 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemv_mma_kernelILi12E -s 40 -c 4 -o $OUT/prof_mma_q4k \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full_q4k.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemv_mma_kernelILi14E -s 30 -c 2 -o $OUT/prof_mma_q6k \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full_q6k.log 2>&1
ls $OUT
