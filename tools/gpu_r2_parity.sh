#!/bin/bash
# new parity tests of round 2 + the 70B-shape probe after the column-slab down projection
mkdir -p gpurun_out/r2par
timeout 1200 python -m pytest tests/test_gpu_engine.py -x -q -k "wide_ffn or mixtral" 2>&1 | tail -15 > gpurun_out/r2par/wide.log; cat gpurun_out/r2par/wide.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "gguf_py" 2>&1 | tail -5 > gpurun_out/r2par/gguf_py.log; cat gpurun_out/r2par/gguf_py.log
timeout 1500 python -m pytest tests/test_gpu_full_depth.py -x -q 2>&1 | tail -25 > gpurun_out/r2par/full_depth.log; cat gpurun_out/r2par/full_depth.log
cat > /tmp/probe.py <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
from zerfoo_b200 import gguf as G, engine
import bench
wl = sys.argv[1]
p = f"/tmp/zb200_models/probe_{wl}.gguf"
os.makedirs("/tmp/zb200_models", exist_ok=True)
t = time.time(); G.write_synthetic_gguf(p, G.preset(wl), fast=True); print(wl, "gen s", round(time.time() - t, 1), "GB", round(os.path.getsize(p) / 1e9, 2), flush=True)
t = time.time(); g = engine.load_file(p, max_seq=512); print(wl, "load s", round(time.time() - t, 1), flush=True)
info = g.refresh_info()
first = g.prefill(bench.PROMPT)
toks, ms = g.decode_n(first, 8)
toks, ms = g.decode_n(toks[-1], 32)
print(wl, "ms/step", ms / 32, "tok/s", 32000 / ms, "launches", info.launches_per_step, "hbm frac", info.weight_bytes_per_token / (ms / 32 / 1000) / 1e9 / 6455.3, flush=True)
g.close()
os.remove(p)
PY
timeout 900 python /tmp/probe.py c4 > gpurun_out/r2par/probe_c4.log 2>&1; cat gpurun_out/r2par/probe_c4.log
