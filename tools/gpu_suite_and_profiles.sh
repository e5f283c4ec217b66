#!/bin/bash
# full GPU suite + round-2 ncu evidence: launch lists of the C2 and C4 (70B-shape) decode steps, one --set full capture of the
# dominant kernel (Q4_K tensor-core GEMV) inside the C4 step
mkdir -p gpurun_out/r2suite
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2suite/tests.log; cat gpurun_out/r2suite/tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3500 -c 400 --csv --log-file gpurun_out/r2suite/launches_c2.csv \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu --no-also > gpurun_out/r2suite/ncu_c2.log 2>&1
tail -2 gpurun_out/r2suite/ncu_c2.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 9500 -c 900 --csv --log-file gpurun_out/r2suite/launches_c4.csv \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --no-also > gpurun_out/r2suite/ncu_c4.log 2>&1
tail -2 gpurun_out/r2suite/ncu_c4.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_mma --launch-skip 4600 -c 1 -f -o gpurun_out/r2suite/c4_q4k \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --no-also > gpurun_out/r2suite/ncu_c4_full.log 2>&1
tail -3 gpurun_out/r2suite/ncu_c4_full.log
ncu -i gpurun_out/r2suite/c4_q4k.ncu-rep --page raw --csv > gpurun_out/r2suite/c4_q4k_raw.csv 2>/dev/null
ncu -i gpurun_out/r2suite/c4_q4k.ncu-rep --page source --print-source sass --csv > gpurun_out/r2suite/c4_q4k_sass.csv 2>/dev/null
rm -f gpurun_out/r2suite/c4_q4k.ncu-rep
ls -la gpurun_out/r2suite
