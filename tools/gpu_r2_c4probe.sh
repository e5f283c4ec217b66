#!/bin/bash
# calibration: how long do generation (fast random blocks), load and a decode step of the 70B-shape Q4_K_M model take on one B200?
mkdir -p gpurun_out/r2c4
nproc > gpurun_out/r2c4/box.txt; free -g >> gpurun_out/r2c4/box.txt; df -h /tmp >> gpurun_out/r2c4/box.txt
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
timeout 300 python /tmp/probe.py c3 >> gpurun_out/r2c4/probe.log 2>&1
timeout 900 python /tmp/probe.py c4 >> gpurun_out/r2c4/probe.log 2>&1
cat gpurun_out/r2c4/box.txt gpurun_out/r2c4/probe.log
