#!/bin/bash
# last sanity of the round on one B200: engine + tensor-core GEMV tests on the final library, then how long the 70B shape takes to load now
mkdir -p gpurun_out/r2last
timeout 100 python -m pytest tests/test_gpu_engine.py tests/test_gpu_mma.py -x -q 2>&1 | tail -3 > gpurun_out/r2last/tests.log; cat gpurun_out/r2last/tests.log
cat > /tmp/probe.py <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
from zerfoo_b200 import engine
import bench
t = time.time(); p = bench.model_path("c4", fast=True); print("gen s", round(time.time() - t, 1), flush=True)
t = time.time(); g = engine.load_file(p, max_seq=512); print("load s", round(time.time() - t, 1), flush=True)
first = g.prefill(bench.PROMPT)
toks, ms = g.decode_n(first, 8)
toks, ms = g.decode_n(toks[-1], 20)
print("c4 ms/step", ms / 20, "tok/s", 20000 / ms, flush=True)
g.close()
PY
timeout 150 python /tmp/probe.py > gpurun_out/r2last/c4_load.log 2>&1; cat gpurun_out/r2last/c4_load.log
