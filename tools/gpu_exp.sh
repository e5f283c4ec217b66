#!/bin/bash
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout -s KILL 600 python -m pytest tests/test_gpu_stream.py -m gpu -q --maxfail=5 -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
for cfg in "0 0" "2 0" "2 4" "3 2" "1 4" "2 2" "3 1"; do
  set -- $cfg
  echo "== CPS=$1 R=$2"
  ZB_GEMV_CPS=$1 ZB_GEMV_R=$2 timeout 300 python tools/gemv_bench.py --pdl --json $OUT/gemv_cps$1_r$2.json 2>&1 | tee $OUT/gemv_cps$1_r$2.log
done
