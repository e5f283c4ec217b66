#!/usr/bin/env python
"""Expert-indirect GEMV at Mixtral 8x7B size (C5 shapes, top-2 of 8 experts): tensor-core kernel vs CUDA-core kernel (GPU box only).
Experts are rotated so that every launch streams from HBM; launches are captured in a PDL-chained CUDA graph."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zerfoo_b200 import gguf as G, kernels as K

peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
rng = np.random.default_rng(0)
E = 8
for label, m, k, pairs in (("c5.expert_gate_up", 28672, 4096, True), ("c5.expert_down", 4096, 14336, False)):
    rb = k // 256 * 144
    raw = rng.integers(0, 256, size=E * m * rb, dtype=np.uint8)
    raw.reshape(-1, 144)[:, :4] = np.frombuffer(np.array([0.01, 0.005], np.float16).tobytes(), np.uint8)
    x = torch.randn(2 * k, device="cuda")
    sels = [torch.tensor([(2 * i) % E, (2 * i + 1) % E], dtype=torch.int32, device="cuda") for i in range(4)]
    nout = m // 2 if pairs else m
    y = torch.empty(2 * nout, device="cuda")
    for name, W, run in (("tensor-core", K.MmaWeight(G.Q4_K, raw, E * m, k, experts=E), K.gemv_mma),
                         ("CUDA-core", K.StreamWeight(G.Q4_K, raw, E * m, k, experts=E), K.gemv_stream)):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for s in sels: run(W, x, sel=s, a_slot_stride=k, swiglu_pairs=pairs, y=y, pdl=True)
        st.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for i in range(16): run(W, x, sel=sels[i % 4], a_slot_stride=k, swiglu_pairs=pairs, y=y, pdl=True)
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): gr.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / 48
        alg = 2 * (m * rb + 4 * k + 4 * m)
        print(f"{label:18s} 2 x {m}x{k} Q4_K  {name:12s} {us:8.2f} us  {alg / us / 1e3:8.1f} GB/s  {alg / us / 1e3 / peak * 100:5.1f}% of measured peak", flush=True)
        del W
