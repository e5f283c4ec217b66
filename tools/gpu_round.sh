#!/bin/bash
# One GPU-box pass: full parity suite, bench lines (c2 headline, c1 north-star target, c3 B=32), ncu launch list + full captures.
set -u
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
( timeout -s KILL 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider 2>&1 | tail -15 ) > $OUT/pytest_gpu.log
( timeout 600 python bench.py --steps 128 --warmup 8 ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err
( timeout 600 python bench.py --steps 128 --warmup 8 --workload c1 ) > $OUT/bench_c1.json 2> $OUT/bench_c1.err
( timeout 600 python bench.py --impl reference --steps 16 --warmup 2 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
( timeout 900 python bench.py --steps 64 --warmup 4 --workload c3 --batch 32 ) > $OUT/bench_c3_b32.json 2> $OUT/bench_c3_b32.err
( timeout 900 python bench.py --steps 64 --warmup 4 --workload c3 --no-cpu ) > $OUT/bench_c3_b1.json 2> $OUT/bench_c3_b1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 340 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemv_stream_kernelILi12E -s 40 -c 4 -o $OUT/prof_q4k \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_stream_kernel -s 70 -c 1 -o $OUT/prof_head \
    python tools/gemv_bench.py --only c2.head > $OUT/ncu_head.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 20 -c 2 -o $OUT/prof_gemm_tc \
    python bench.py --steps 2 --warmup 3 --workload c3 --batch 32 > $OUT/ncu_tc.log 2>&1
tail -3 $OUT/pytest_gpu.log
for f in bench_c2 bench_c1 bench_ref bench_c3_b32 bench_c3_b1; do echo "== $f"; cut -c1-900 $OUT/$f.json; tail -2 $OUT/$f.err; done
