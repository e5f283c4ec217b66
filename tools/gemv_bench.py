#!/usr/bin/env python
"""Stand-alone timing of the streamed GEMV on the BASELINE matrix shapes (GPU box only).
Weights are rotated through enough copies to exceed the 126 MB L2, CUDA events on the launch stream.
    python tools/gemv_bench.py [--pdl] [--json out.json]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zerfoo_b200 import gguf as G, kernels as K

SHAPES = [  # (label, qtype, rows, K)
    ("c2.qkv", G.Q4_K, 5120, 3072), ("c2.o", G.Q4_K, 3072, 3072), ("c2.gate_up", G.Q4_K, 16384, 3072), ("c2.down", G.Q4_K, 3072, 8192),
    ("c2.down6", G.Q6_K, 3072, 8192), ("c2.head", G.Q6_K, 128256, 3072),
    ("c1.qkv", G.Q4_0, 1536, 1152), ("c1.gate_up", G.Q4_0, 13824, 1152), ("c1.down", G.Q4_0, 1152, 6912), ("c1.head", G.Q4_0, 262144, 1152),
    ("c4.gate_up", G.Q4_K, 57344, 8192), ("c4.qkv", G.Q4_K, 10240, 8192), ("c4.o", G.Q4_K, 8192, 8192), ("c4.down", G.Q4_K, 8192, 28672),
    ("c4.down6", G.Q6_K, 8192, 28672), ("c4.head", G.Q6_K, 128256, 8192), ("c4tp8.down", G.Q4_K, 8192, 3584), ("c4tp8.gate_up", G.Q4_K, 7168, 8192),
    ("c4tp8.qkv", G.Q4_K, 1280, 8192), ("c4tp8.o", G.Q4_K, 8192, 1024),
    ("c3.gate_up", G.Q5_K, 28672, 4096), ("c3.down", G.Q5_K, 4096, 14336), ("q8.head", G.Q8_0, 32000, 4096),
]

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--pdl", action="store_true"); ap.add_argument("--json"); ap.add_argument("--only"); ap.add_argument("--xrep", type=int, default=0, help="read x from this many identical copies (hot-spot probe)"); ap.add_argument("--mma", action="store_true", help="tensor-core GEMV (gemv_mma.cu) where the format supports it")
    a = ap.parse_args()
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
    rng = np.random.default_rng(0); out = []
    for label, qt, m, k in SHAPES:
        if a.only and a.only not in label: continue
        rb = k // G.BLOCK_ELEMS[qt] * G.BLOCK_BYTES[qt]
        raw = rng.integers(0, 256, size=m * rb, dtype=np.uint8)
        if qt in (G.Q4_0, G.Q8_0): raw.reshape(-1, G.BLOCK_BYTES[qt])[:, :2] = np.frombuffer(np.float16(0.01).tobytes(), np.uint8)
        if qt in (G.Q4_K, G.Q5_K): raw.reshape(-1, G.BLOCK_BYTES[qt])[:, :4] = np.frombuffer(np.array([0.01, 0.005], np.float16).tobytes(), np.uint8)
        if qt == G.Q6_K: raw.reshape(-1, 210)[:, 208:] = np.frombuffer(np.float16(0.01).tobytes(), np.uint8)
        nbytes = m * rb
        copies = max(2, int(300e6 // nbytes) + 1)
        copies = min(copies, 64)
        L = K._lib.load()
        use_mma = a.mma and L.zb_mma_check(qt, m, k) == 0
        if a.mma and not use_mma: continue
        ws = [(K.MmaWeight if use_mma else K.StreamWeight)(qt, raw, m, k) for _ in range(copies)]
        run = K.gemv_mma if use_mma else K.gemv_stream
        x = torch.randn(k, device="cuda"); y = torch.empty(m, device="cuda")
        kw = {}
        if a.xrep > 1 and use_mma:
            x = x.repeat(a.xrep).contiguous(); kw = dict(a_replicas=a.xrep, a_replica_stride=k)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for i in range(copies): run(ws[i], x, y=y, pdl=a.pdl, **kw)
        st.synchronize()
        iters = max(copies * 2, 40)
        # capture the launches in a CUDA graph so the CPU launch rate (ctypes, ~10 us/call) is out of the measurement
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for i in range(iters): run(ws[i % copies], x, y=y, pdl=a.pdl, **kw)
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): gr.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / (3 * iters)
        alg = nbytes + 4 * k + 4 * m
        gbs = alg / us / 1e3
        out.append({"shape": label, "type": G.TYPE_NAMES[qt], "rows": m, "K": k, "MB": nbytes / 1e6, "us": us, "GBps": gbs, "frac_measured_peak": gbs / peak})
        print(f"{label:12s} {G.TYPE_NAMES[qt]:5s} {m:7d}x{k:<6d} {nbytes/1e6:8.1f} MB  {us:8.2f} us  {gbs:8.1f} GB/s  {gbs/peak*100:5.1f}% of measured peak", flush=True)
        del ws
    if a.json: json.dump(out, open(a.json, "w"), indent=1)

if __name__ == "__main__":
    main()
