#!/bin/bash
TAG=${1:-moe}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
timeout -s KILL 600 python -m pytest tests/test_gpu_mma.py tests/test_gpu_engine.py -m gpu -q --maxfail=12 -p no:cacheprovider > $OUT/pytest.log 2>&1; tail -6 $OUT/pytest.log
python - > $OUT/moe_time.log 2>&1 <<'PY'
import os, sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import modelzoo as Z
from zerfoo_b200 import engine
for tc in ("1", "0"):
    os.environ["ZB_GEMV_TC"] = tc
    g = engine.load_file(Z.path("mixtral_q4_k_m"))
    first = g.prefill(Z.PROMPT)
    toks, ms = g.decode_n(first, 64)
    toks, ms = g.decode_n(toks[-1], 64)
    print("mixtral miniature, ZB_GEMV_TC=%s: %.3f ms/step, %d launches/step" % (tc, ms / 64, g.refresh_info().launches_per_step))
    g.close()
PY
cat $OUT/moe_time.log | tail -3
( timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cut -c1-160 $OUT/bench_c2.json; tail -2 $OUT/bench_c2.err
