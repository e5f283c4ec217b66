#!/bin/bash
# raw phase trace + one ncu --set full capture (source-level stalls) of the persistent kernel
mkdir -p gpurun_out/r2mega6
ZB_MEGA_TRACE_DUMP=gpurun_out/r2mega6/trace_c2.npz timeout 300 python tools/mega_trace.py c2 > gpurun_out/r2mega6/trace_c2.txt 2>&1
tail -3 gpurun_out/r2mega6/trace_c2.txt
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 25 -c 1 -f -o gpurun_out/r2mega6/mega_v3 python /tmp/run_mega.py > gpurun_out/r2mega6/ncu.log 2>&1
tail -3 gpurun_out/r2mega6/ncu.log
ncu -i gpurun_out/r2mega6/mega_v3.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r2mega6/srcboth.csv 2>/dev/null
ncu -i gpurun_out/r2mega6/mega_v3.ncu-rep --page raw --csv > gpurun_out/r2mega6/raw.csv 2>/dev/null
rm -f gpurun_out/r2mega6/mega_v3.ncu-rep
ls -la gpurun_out/r2mega6/
