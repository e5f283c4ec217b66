#!/bin/bash
# round 2: persistent whole-token kernel -- parity tests, then a short C2 bench with and without it
mkdir -p gpurun_out/r2mega
timeout 600 python -m pytest tests/test_gpu_mega.py -x -q 2>&1 | tail -25 > gpurun_out/r2mega/tests.log
cat gpurun_out/r2mega/tests.log
timeout 300 python bench.py --steps 64 --warmup 8 --no-cpu > gpurun_out/r2mega/bench_mega.json 2> gpurun_out/r2mega/bench_mega.err
tail -c 1500 gpurun_out/r2mega/bench_mega.json; tail -5 gpurun_out/r2mega/bench_mega.err
ZB_MEGA=0 timeout 300 python bench.py --steps 64 --warmup 8 --no-cpu > gpurun_out/r2mega/bench_graph.json 2> gpurun_out/r2mega/bench_graph.err
head -c 300 gpurun_out/r2mega/bench_graph.json
