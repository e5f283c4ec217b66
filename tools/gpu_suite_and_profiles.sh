#!/bin/bash
# final pass of a round on one B200: the whole -m gpu suite, smoke(), and a short bench line (C2 main, no sub-records)
mkdir -p gpurun_out/r2final
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2final/tests.log; cat gpurun_out/r2final/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2final/smoke.log 2>&1; tail -2 gpurun_out/r2final/smoke.log
timeout 300 python bench.py --workload c2 --steps 64 --warmup 8 --no-cpu --no-also > gpurun_out/r2final/bench_c2.json 2> gpurun_out/r2final/bench_c2.err
python -c "
import json; d=json.load(open('gpurun_out/r2final/bench_c2.json')); print('c2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_hbm_frac'])"
