#!/bin/bash
# One GPU-box pass: bench lines (c2 headline, c1 north-star target, PDL off A/B), ncu launch list + full captures.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [tests]
set -u
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
if [ "${2:-}" = "tests" ]; then
  ( timeout -s KILL 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider 2>&1 | tail -40 ) > $OUT/pytest_gpu.log
fi
( timeout 600 python bench.py --steps 128 --warmup 8 ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err
( timeout 600 python bench.py --steps 128 --warmup 8 --workload c1 ) > $OUT/bench_c1.json 2> $OUT/bench_c1.err
( ZB_NO_PDL=1 timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2_nopdl.json 2> $OUT/bench_c2_nopdl.err
( ZB_NO_PDL=1 timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --workload c1 ) > $OUT/bench_c1_nopdl.json 2> $OUT/bench_c1_nopdl.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemv_stream_kernelILi12E -s 30 -c 4 -o $OUT/prof_q4k \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemv_stream_kernelILi14E -s 28 -c 2 -o $OUT/prof_q6k \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full6.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:decode_attn_kernel -s 10 -c 1 -o $OUT/prof_attn \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_attn.log 2>&1
[ -f $OUT/pytest_gpu.log ] && tail -3 $OUT/pytest_gpu.log
cat $OUT/bench_c2.json $OUT/bench_c1.json $OUT/bench_c2_nopdl.json $OUT/bench_c1_nopdl.json; tail -2 $OUT/bench_c2.err
