#!/bin/bash
# quick iteration: stream-kernel parity + GEMV micro-benchmark (+ optional engine bench / ncu)
TAG=${1:-it}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_engine.py -m gpu -q --maxfail=10 -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -15 $OUT/pytest.log
timeout 600 python tools/gemv_bench.py --json $OUT/gemv_bench.json 2>&1 | tee $OUT/gemv_bench.log
timeout 600 python tools/gemv_bench.py --pdl --json $OUT/gemv_bench_pdl.json 2>&1 | tee $OUT/gemv_bench_pdl.log
if [ "${2:-}" = "bench" ]; then
  ( timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cat $OUT/bench_c2.json | cut -c1-400
  ( timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu --workload c1 ) > $OUT/bench_c1.json 2> $OUT/bench_c1.err; cat $OUT/bench_c1.json | cut -c1-400
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 330 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemv_stream_kernelILi14E -s 85 -c 2 -o $OUT/prof_head \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_head.log 2>&1
fi
