#!/bin/bash
TAG=${1:-b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_stream.py::test_decode_attention_stage tests/test_gpu_engine.py -m gpu -q --maxfail=10 -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -40 $OUT/pytest.log
