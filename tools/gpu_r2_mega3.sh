#!/bin/bash
# round 2: persistent kernel after the prologue / carve-out / barrier changes: parity, phase trace, bench with and without
mkdir -p gpurun_out/r2mega5
timeout 900 python -m pytest tests/test_gpu_mega.py -x -q 2>&1 | tail -25 > gpurun_out/r2mega5/tests.log
cat gpurun_out/r2mega5/tests.log
timeout 300 python tools/mega_trace.py c2 > gpurun_out/r2mega5/trace_c2.txt 2>&1
cat gpurun_out/r2mega5/trace_c2.txt
timeout 300 python bench.py --steps 64 --warmup 8 --no-cpu > gpurun_out/r2mega5/bench_mega.json 2> gpurun_out/r2mega5/bench_mega.err
head -c 400 gpurun_out/r2mega5/bench_mega.json; tail -5 gpurun_out/r2mega5/bench_mega.err
