#!/bin/bash
TAG=${1:-pfp}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
python bench.py --workload c3 --layers 2 --prefill 4096 --steps 1 --warmup 1 > $OUT/warm.json 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_prefill.csv python bench.py --workload c3 --layers 2 --prefill 4096 --steps 1 --warmup 1 > $OUT/ncu.log 2>&1
python tools/launch_shares.py $OUT/launches_prefill.csv > $OUT/summary_prefill.txt 2>&1; cat $OUT/summary_prefill.txt | head -40
