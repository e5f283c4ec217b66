#!/usr/bin/env python
"""Phase timeline of the tensor-core GEMV (GPU box only): ZB_MMA_TRACE=1 makes every CTA stamp %globaltimer at
0 start | 1 ring filled | 2 after griddepcontrol.wait | 3 x built | 4 fragments built | 5 main loop done | 6 CTA sync | 7 end.
Prints, per launch of a PDL-chained graph, the median / max over CTAs of each stamp relative to the launch's first CTA start,
and the gap to the previous launch.   python tools/gemv_trace.py"""
import ctypes as C, os, sys
os.environ["ZB_MMA_TRACE"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zerfoo_b200 import gguf as G, kernels as K, lib

SHAPES = [("c2.o", G.Q4_K, 3072, 3072), ("c2.gate_up", G.Q4_K, 16384, 3072), ("c2.down", G.Q4_K, 3072, 8192), ("c2.down6", G.Q6_K, 3072, 8192)]
L = lib.load()
rng = np.random.default_rng(0)
used = 0
for label, qt, m, k in SHAPES:
    if L.zb_mma_check(qt, m, k) != 0:
        continue
    rb = k // 256 * G.BLOCK_BYTES[qt]
    raw = rng.integers(0, 256, size=m * rb, dtype=np.uint8)
    n = 8
    ws = [K.MmaWeight(qt, raw, m, k) for _ in range(n)]
    x = torch.randn(k, device="cuda"); y = torch.empty(m, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(n): K.gemv_mma(ws[i], x, y=y, pdl=True)
    st.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        for i in range(n): K.gemv_mma(ws[i], x, y=y, pdl=True)
    gr.replay(); gr.replay(); torch.cuda.synchronize()
    buf = np.zeros(64 * 148 * 8, np.uint64)
    got = L.zb_mma_trace_read(buf.ctypes.data_as(C.c_void_p), 64)
    tr = buf.reshape(64, 148, 8)[used + n: used + 2 * n].astype(np.float64)
    used += 2 * n
    print(f"== {label} {G.TYPE_NAMES[qt]} {m}x{k} ({got} launches traced); ns relative to this launch's first CTA start: median/max over CTAs")
    prev_end = None
    for i in range(n):
        t = tr[i]
        live = t[:, 0] > 0
        t = t[live]
        t0 = t[:, 0].min()
        cols = " ".join(f"{int(np.median(t[:, j] - t0)):6d}/{int((t[:, j] - t0).max()):6d}" for j in range(8))
        gap = "" if prev_end is None else f" start-prev_end {int(t0 - prev_end):6d}"
        print(f"  L{i} ctas {int(live.sum()):3d}: {cols}{gap}")
        prev_end = t[:, 7].max()
    print(f"  whole graph: {(tr[n-1][:, 7].max() - tr[0][tr[0][:, 0] > 0][:, 0].min()) / n:.0f} ns per launch")
    del ws
