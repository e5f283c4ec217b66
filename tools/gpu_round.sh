#!/bin/bash
# One GPU-box pass: parity tests, bench lines (c2 headline, c1 north-star target), ncu launch list + one full capture.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/pytest_gpu.log
( timeout 600 python bench.py --steps 128 --warmup 8 ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err
( timeout 600 python bench.py --steps 128 --warmup 8 --workload c1 ) > $OUT/bench_c1.json 2> $OUT/bench_c1.err
( timeout 600 python bench.py --impl reference --steps 16 --warmup 2 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:FmtQ4_K -s 60 -c 2 -o $OUT/prof_q4k \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1
tail -3 $OUT/pytest_gpu.log; cat $OUT/bench_c2.json $OUT/bench_c1.json $OUT/bench_ref.json; tail -2 $OUT/bench_c2.err
