#!/bin/bash
mkdir -p gpurun_out/r2c4
timeout 600 python tools/gemv_bench.py --only c4 --pdl --mma > gpurun_out/r2c4/gemv_mma_c4.log 2>&1
timeout 600 python tools/gemv_bench.py --only c4 --pdl > gpurun_out/r2c4/gemv_stream_c4.log 2>&1
cat gpurun_out/r2c4/gemv_mma_c4.log; echo; cat gpurun_out/r2c4/gemv_stream_c4.log
