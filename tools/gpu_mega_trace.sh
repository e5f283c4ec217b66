#!/bin/bash
# persistent kernel iteration: parity, raw phase trace, short bench
OUT=gpurun_out/${1:-r2mega7}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_mega.py -x -q 2>&1 | tail -15 > $OUT/tests.log
cat $OUT/tests.log
ZB_MEGA_TRACE_DUMP=$OUT/trace_c2.npz timeout 300 python tools/mega_trace.py c2 > $OUT/trace_c2.txt 2>&1
cat $OUT/trace_c2.txt
timeout 300 python bench.py --steps 64 --warmup 8 --no-cpu > $OUT/bench_mega.json 2> $OUT/bench_mega.err
head -c 300 $OUT/bench_mega.json; tail -5 $OUT/bench_mega.err
