#!/bin/bash
# new tests (fp16 KV, sliding window, mega opt-in) + C2 A/B: random-block weights vs quantized-float weights
mkdir -p gpurun_out/r2check2
timeout 1500 python -m pytest tests/test_gpu_mega.py tests/test_gpu_prefill.py tests/test_gpu_batch.py tests/test_gpu_engine.py tests/test_gpu_tp.py tests/test_gpu_stream.py tests/test_gpu_mma.py -q 2>&1 | tail -25 > gpurun_out/r2check2/tests.log; cat gpurun_out/r2check2/tests.log
for i in 1 2; do
timeout 300 python bench.py --workload c2 --steps 64 --warmup 8 --no-cpu --no-also > gpurun_out/r2check2/c2_fast_$i.json 2>gpurun_out/r2check2/err.log
python -c "
import json;d=json.load(open('gpurun_out/r2check2/c2_fast_$i.json'));print('fast', d['value'], d['ms_per_step'], d['roofline']['us_per_launch'])"
ZB_BENCH_QUANTIZED_WEIGHTS=1 timeout 300 python bench.py --workload c2 --steps 64 --warmup 8 --no-cpu --no-also > gpurun_out/r2check2/c2_quant_$i.json 2>>gpurun_out/r2check2/err.log
python -c "
import json;d=json.load(open('gpurun_out/r2check2/c2_quant_$i.json'));print('quant', d['value'], d['ms_per_step'], d['roofline']['us_per_launch'])"
done
tail -3 gpurun_out/r2check2/err.log
