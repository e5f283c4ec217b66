#!/bin/bash
# ncu full capture of one micro-benchmark shape:  bash tools/gpu_prof.sh <tag> <shape> [env...]
TAG=$1; SHAPE=$2; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_stream_kernel -s 70 -c 1 -o $OUT/prof_$SHAPE \
   python tools/gemv_bench.py --only $SHAPE > $OUT/ncu_$SHAPE.log 2>&1
tail -3 $OUT/ncu_$SHAPE.log
