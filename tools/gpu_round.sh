#!/bin/bash
# One GPU-box pass: full parity suite, bench lines (c2 headline, c1 north-star target, c3 B=32 / B=1), PDL A/B, ncu launch list + full captures.
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
( ZB_NO_PDL=1 timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2_nopdl.json 2> $OUT/bench_c2_nopdl.err
( timeout 900 python bench.py --steps 64 --warmup 4 --workload c3 --batch 32 ) > $OUT/bench_c3_b32.json 2> $OUT/bench_c3_b32.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 340 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemv_stream_kernelILi12E -s 40 -c 4 -o $OUT/prof_q4k \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1
tail -3 $OUT/pytest_gpu.log
for f in bench_c2 bench_c1 bench_ref bench_c2_nopdl bench_c3_b32; do echo "== $f"; cut -c1-260 $OUT/$f.json; tail -2 $OUT/$f.err; done
python -c "
import json
d=json.load(open('$OUT/bench_c2.json')); print(json.dumps(d['roofline'])[:1200]); print(d['clocks'])
d=json.load(open('$OUT/bench_c1.json')); print(json.dumps(d['roofline'])[:600])
"
