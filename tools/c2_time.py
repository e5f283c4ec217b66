#!/usr/bin/env python
"""Time the C2 CUDA-graph decode step of whatever checkout this script's cwd holds (regression bisect across worktrees)."""
import os, sys
sys.path.insert(0, os.getcwd())
import bench
from zerfoo_b200 import engine
kw = {}
try:
    p = bench.model_path("c2", fast=True)
except TypeError:
    p = bench.model_path("c2")
try:
    g = engine.load_file(p, max_seq=512, mega=False)
except TypeError:
    g = engine.load_file(p, max_seq=512)
first = g.prefill(bench.PROMPT)
toks, _ = g.decode_n(first, 8)
best = 1e9
for _ in range(3):
    toks, ms = g.decode_n(toks[-1], 64)
    best = min(best, ms / 64)
print(os.path.basename(os.getcwd()), "launches", g.refresh_info().launches_per_step, "ms/step", round(best, 4), "tok/s", round(1000 / best, 1), "pos", g.position)
g.close()
