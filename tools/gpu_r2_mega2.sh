#!/bin/bash
mkdir -p gpurun_out/r2mega
timeout 900 python -m pytest tests/test_gpu_mega.py -q 2>&1 | tail -25 > gpurun_out/r2mega/tests2.log
cat gpurun_out/r2mega/tests2.log
timeout 300 python tools/mega_trace.py c2 > gpurun_out/r2mega/trace_c2.txt 2>&1
cat gpurun_out/r2mega/trace_c2.txt
