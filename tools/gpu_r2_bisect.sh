#!/bin/bash
mkdir -p gpurun_out/r2bisect
export ZB_MEGA=0
( timeout 300 python tools/c2_time.py ) >> gpurun_out/r2bisect/times4.log 2>&1
cat gpurun_out/r2bisect/times4.log
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_mma.py -q -x 2>&1 | tail -4
