#!/bin/bash
# rehearsal of the driver's N=1 sequence: reference arm, then our arm (main 70B-shape line + sub-records)
mkdir -p gpurun_out/r2bench
( time timeout 800 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2bench/ref_n1.json 2> gpurun_out/r2bench/ref_n1.err
tail -c 600 gpurun_out/r2bench/ref_n1.json; tail -4 gpurun_out/r2bench/ref_n1.err
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2bench/ours_n1.json 2> gpurun_out/r2bench/ours_n1.err
tail -c 3000 gpurun_out/r2bench/ours_n1.json; tail -8 gpurun_out/r2bench/ours_n1.err
