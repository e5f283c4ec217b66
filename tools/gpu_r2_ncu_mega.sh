#!/bin/bash
# one ncu --set full capture of the persistent kernel (source-level stall reasons)
mkdir -p gpurun_out/r2mega
cat > /tmp/run_mega.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
from zerfoo_b200 import engine
g = engine.load_file(bench.model_path("c2"), max_seq=512)
first = g.prefill(bench.PROMPT)
toks, ms = g.decode_n(first, 8)
print("ms/step", ms / 8)
g.close()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 25 -c 1 -f -o gpurun_out/r2mega/mega_v1 python /tmp/run_mega.py > gpurun_out/r2mega/ncu.log 2>&1
tail -5 gpurun_out/r2mega/ncu.log
ls -la gpurun_out/r2mega/
