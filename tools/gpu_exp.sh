#!/bin/bash
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 600 python -m pytest tests/test_gpu_stream.py tests/test_gpu_engine.py -m gpu -q --maxfail=5 -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
for gc in 3 2 1; do
  echo "== GRID_CPS=$gc"
  ( ZB_GEMV_GRID_CPS=$gc timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2_g$gc.json 2> $OUT/bench_c2_g$gc.err; cut -c1-140 $OUT/bench_c2_g$gc.json
  ( ZB_GEMV_GRID_CPS=$gc timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --workload c1 ) > $OUT/bench_c1_g$gc.json 2> $OUT/bench_c1_g$gc.err; cut -c1-140 $OUT/bench_c1_g$gc.json
done
ZB_GEMV_GRID_CPS=2 timeout 300 python tools/gemv_bench.py --pdl --json $OUT/gemv_g2.json > $OUT/gemv_g2.log 2>&1; cat $OUT/gemv_g2.log
