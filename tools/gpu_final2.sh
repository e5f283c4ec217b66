#!/bin/bash
# Round-end pass: full GPU parity suite, sanitizer on the tensor-core GEMV, bench lines, ncu launch list + full captures, C3 batch/prefill.
set -u
TAG=${1:-final2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export ZB_BENCH_MODEL_DIR=/tmp/zb200_models
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
( timeout -s KILL 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider 2>&1 | tail -15 ) > $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
( timeout -s KILL 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mma.py -m gpu -q -x -p no:cacheprovider -k "40x512 or 257x2048 or prologues or 1030x512" 2>&1 | tail -6 ) > $OUT/sanitizer_memcheck.log; tail -3 $OUT/sanitizer_memcheck.log
( timeout -s KILL 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_mma.py -m gpu -q -x -p no:cacheprovider -k "40x512 or 1030x512" 2>&1 | tail -6 ) > $OUT/sanitizer_racecheck.log; tail -3 $OUT/sanitizer_racecheck.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
( timeout 600 python bench.py --steps 128 --warmup 8 ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cut -c1-160 $OUT/bench_c2.json; tail -2 $OUT/bench_c2.err
( timeout 600 python bench.py --steps 128 --warmup 8 --workload c1 --no-cpu ) > $OUT/bench_c1.json 2> $OUT/bench_c1.err; cut -c1-160 $OUT/bench_c1.json; tail -2 $OUT/bench_c1.err
( ZB_GEMV_TC=0 timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu ) > $OUT/bench_c2_cudacore.json 2> $OUT/bench_c2_cudacore.err; cut -c1-160 $OUT/bench_c2_cudacore.json; tail -2 $OUT/bench_c2_cudacore.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 340 --csv --log-file $OUT/launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemv_mma_kernelILi12E -s 40 -c 4 -o $OUT/prof_mma_q4k \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full_q4k.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemv_mma_kernel -c 3 -o $OUT/prof_mma_large \
    python tools/gemv_bench.py --mma --only c4.gate_up > $OUT/ncu_full_large.log 2>&1
( timeout 500 python bench.py --steps 32 --warmup 4 --workload c3 --batch 32 ) > $OUT/bench_c3_b32.json 2> $OUT/bench_c3_b32.err; cut -c1-200 $OUT/bench_c3_b32.json; tail -2 $OUT/bench_c3_b32.err
( timeout 500 python bench.py --workload c3 --prefill 4096 ) > $OUT/bench_c3_prefill.json 2> $OUT/bench_c3_prefill.err; cut -c1-300 $OUT/bench_c3_prefill.json; tail -2 $OUT/bench_c3_prefill.err
ls $OUT
