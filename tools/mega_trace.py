#!/usr/bin/env python
"""Phase timeline of the persistent whole-token kernel (ZB_MEGA_TRACE=1): per op class, median / max over CTAs and ops of
  prologue (x load + norm + digit fragments), main loop (warp 0 / whole CTA), epilogue (partial-sum exchange + stores), barrier wait.
SM-clock cycles of thread 0 of every CTA -> microseconds at the clock given (default 1.965 GHz).

    ZB_MEGA_TRACE=1 python tools/mega_trace.py [workload] [ghz]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ZB_MEGA_TRACE"] = "1"
import bench  # noqa: E402
from zerfoo_b200 import engine, gguf as G, lib  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    ghz = float(sys.argv[2]) if len(sys.argv) > 2 else 1.965
    path = bench.model_path(wl)
    g = engine.load_file(path, max_seq=512, mega=True)
    first = g.prefill(bench.PROMPT)
    toks, ms = g.decode_n(first, 32)
    print(f"{wl}: {ms / 32 * 1000:.1f} us per step (32 steps), position {g.position}")
    L = lib.load()
    max_ops = 4096
    kinds = (C.c_int * max_ops)()
    ctas = C.c_int()
    n = L.zb_engine_mega_trace(g._h, None, kinds, max_ops, C.byref(ctas))
    buf = np.zeros((n, ctas.value, 8), dtype=np.int64)
    L.zb_engine_mega_trace(g._h, buf.ctypes.data_as(C.c_void_p), kinds, max_ops, C.byref(ctas))
    dump = os.environ.get("ZB_MEGA_TRACE_DUMP")
    if dump:
        np.savez_compressed(dump, t=buf, kinds=np.array(list(kinds[:n]), dtype=np.int32), ghz=ghz)
    us = lambda cyc: cyc / (ghz * 1e3)
    t = buf.astype(np.float64)
    t[buf == 0] = np.nan
    classes = {}
    for i in range(n):
        classes.setdefault(kinds[i], []).append(i)
    step0 = np.nanmin(t[0, :, 0])
    print(f"whole launch (first op start -> last stamp), CTA median: {us(np.nanmedian(np.nanmax(t[:, :, :6].reshape(-1, ctas.value, 6)[-1], axis=1) - t[0, :, 0])):.1f} us")
    print(f"{'op':>28} {'n':>4} {'prolog':>8} {'loop w0':>8} {'loop cta':>8} {'epilog':>8} {'barrier':>8} {'total':>8}   (us, median over CTAs and ops; max in brackets)")
    tot_all = 0.0
    for k, idx in sorted(classes.items()):
        tt = t[idx]
        def d(a, b):
            x = tt[:, :, b] - tt[:, :, a]
            return us(np.nanmedian(x)), us(np.nanmax(x)) if np.isfinite(x).any() else (float('nan'), float('nan'))
        if k >= 100:
            name = f"gemv {G.TYPE_NAMES.get((k - 100) % 1000, '?')} K={256 * (k // 1000)}"
        else:
            name = {0: "embed", 2: "attention", 3: "final"}.get(k, str(k))
        cols = [d(0, 1), d(1, 2), d(1, 3), d(3, 4), d(4, 5), d(0, 5)] if k >= 100 else [(np.nan, np.nan)] * 3 + [d(0, 4), d(4, 5), d(0, 5)]
        tot = cols[-1][0] * len(idx)
        tot_all += tot
        print(f"{name:>28} {len(idx):>4} " + " ".join(f"{c[0]:5.2f}[{c[1]:4.1f}]" if np.isfinite(c[0]) else "      -     " for c in cols) + f"   sum {tot:7.1f}")
    print(f"sum of medians: {tot_all:.1f} us")
    g.close()


if __name__ == "__main__":
    main()
