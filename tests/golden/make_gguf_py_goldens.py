#!/usr/bin/env python
"""Pins the block formats on an implementation that is neither the reference's nor ours: gguf-py (the ggml project's own
Python package, `gguf` 0.19.0 in this image), whose `gguf.quants.dequantize` is the canonical statement of the formats the
reference's kernels decode (internal/cuda/kernels/gemv_q5k.cu:15-23, gemv_q6k.cu:11-25, gemm_q8.cu:1-7,
dequant_q4k.cu:38-93, gemm_q4.cu:89-96).  The upstream tests hold golden bytes only for Q4_K / Q4_0; this
closes Q5_K, Q6_K and Q8_0.

One layout quirk is part of the reference and therefore of the parity target: zerfoo's Q5_K super-block keeps the low
nibbles BEFORE the high bits (ql at [16:144], qh at [144:176]: gemv_q5k.cu:8-12,108-110; the loader hands the GGUF bytes to
the kernel unchanged, model/gguf/loader.go:200-205), whereas ggml's block_q5_K is d, dmin, scales, qh[32], qs[128].  The
committed `raw` is in the reference's order (what the oracle and the CUDA kernels decode); gguf-py is given the same
fields in ggml's order.  Everything else -- scale / min packing, nibble and bit assignment, arithmetic -- is identical.

For every format: 8 blocks produced by this repo's quantizer from N(0,1) data and 24 blocks of raw random bytes (every
nibble / high-bit / 6-bit-scale packing, negative int8 scales) with finite fp16 scale fields -> raw bytes + gguf-py's f32
values, committed as tests/golden/gguf_py/<format>.npz.  Run from the repo root:  python tests/golden/make_gguf_py_goldens.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gguf  # noqa: E402  (gguf-py)
from gguf import quants  # noqa: E402

from zerfoo_b200 import gguf as G  # noqa: E402

GG = {G.Q4_0: gguf.GGMLQuantizationType.Q4_0, G.Q8_0: gguf.GGMLQuantizationType.Q8_0, G.Q4_K: gguf.GGMLQuantizationType.Q4_K,
      G.Q5_K: gguf.GGMLQuantizationType.Q5_K, G.Q6_K: gguf.GGMLQuantizationType.Q6_K}


def blocks(qt: int) -> np.ndarray:
    rng = np.random.default_rng(2024 + qt)
    be, bb = G.BLOCK_ELEMS[qt], G.BLOCK_BYTES[qt]
    q = G.quantize(rng.standard_normal((1, 8 * be), dtype=np.float32), qt).reshape(8, bb)
    r = rng.integers(0, 256, size=(24, bb), dtype=np.uint8)
    sc = (rng.standard_normal((24, 2)) * 0.02).astype(np.float16).view(np.uint8).reshape(24, 4)
    if qt in (G.Q4_0, G.Q8_0):
        r[:, 0:2] = sc[:, 0:2]
    elif qt in (G.Q4_K, G.Q5_K):
        r[:, 0:4] = sc
    else:
        r[:, 208:210] = sc[:, 0:2]
    return np.concatenate([q, r])


def ggml_order(qt: int, raw: np.ndarray) -> np.ndarray:
    if qt != G.Q5_K:
        return raw
    return np.concatenate([raw[:, :16], raw[:, 144:176], raw[:, 16:144]], axis=1)


def main():
    out = os.path.join(ROOT, "tests", "golden", "gguf_py")
    os.makedirs(out, exist_ok=True)
    for qt, gq in GG.items():
        raw = blocks(qt)
        vals = quants.dequantize(np.ascontiguousarray(ggml_order(qt, raw)).reshape(1, -1), gq).astype(np.float32).reshape(-1)
        assert vals.size == raw.shape[0] * G.BLOCK_ELEMS[qt] and np.isfinite(vals).all()
        np.savez_compressed(os.path.join(out, f"{G.TYPE_NAMES[qt]}.npz"), raw=raw, values=vals, gguf_py_version=np.array(getattr(gguf, "__version__", "0.19.0")))
        print(G.TYPE_NAMES[qt], raw.shape, vals[:4])


if __name__ == "__main__":
    main()
