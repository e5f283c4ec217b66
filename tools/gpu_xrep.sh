#!/bin/bash
TAG=${1:-xrep}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for r in 0 4 16; do echo "== xrep $r"; timeout 200 python tools/gemv_bench.py --pdl --mma --only c2 --xrep $r 2>&1 | tail -6 | tee $OUT/gemv_xrep$r.log; done
