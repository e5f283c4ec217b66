"""CPU restatement of the integer tensor-core GEMV arithmetic (csrc/gemv_mma.cu, Q4_K): balanced base-256 digits of a 32-bit
fixed-point image of x, exact integer dot products per 32-weight sub-block, 6-bit scales applied as integers, the min term as a
second integer contraction, one f32 combination per super-block.  It must land within a few f32 ulps of sum |w x| of the f64
oracle -- the bar tests/test_gpu_mma.py::test_mma_integer_path_is_f32_accurate holds the CUDA kernel to."""
import numpy as np
import pytest

from oracle import oracle as O
from zerfoo_b200 import gguf as G


def digits_msb_first(v: np.ndarray) -> np.ndarray:
    """v int64 in [-2^30, 2^30] -> [..., 4] balanced digits d0..d3 with v = d0*2^24 + d1*2^16 + d2*2^8 + d3, each in [-128, 127]:
    the bytes of (v + 0x80808080) ^ 0x80808080 read as s8 (gemv_mma.cu: digit_bytes)."""
    u = ((v + 0x80808080) & 0xFFFFFFFF) ^ 0x80808080
    b = np.stack([(u >> s) & 0xFF for s in (24, 16, 8, 0)], axis=-1).astype(np.int64)
    return np.where(b >= 128, b - 256, b)


def fixed_scale(mx: float):
    eb = (np.float32(mx).view(np.uint32) >> 23) & 0xFF
    sh = int(np.clip(156 - int(eb), -60, 120))
    return np.float32(2.0 ** sh), np.float32(2.0 ** -sh)


def test_digit_bytes_are_balanced_digits():
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.integers(-2**30, 2**30, 10000), [0, 1, -1, 127, 128, -128, -129, 2**30, -2**30, 32767, 32768]]).astype(np.int64)
    d = digits_msb_first(v)
    assert d.min() >= -128 and d.max() <= 127
    assert np.array_equal(d[:, 0] * 2**24 + d[:, 1] * 2**16 + d[:, 2] * 2**8 + d[:, 3], v)


def gemv_q4k_integer_model(raw: np.ndarray, rows: int, K: int, x: np.ndarray) -> np.ndarray:
    blocks = np.ascontiguousarray(raw).view(np.uint8).reshape(rows, K // 256, 144)
    d = blocks[:, :, 0:2].copy().view(np.float16).astype(np.float32)[..., 0]
    dmin = blocks[:, :, 2:4].copy().view(np.float16).astype(np.float32)[..., 0]
    sc = blocks[:, :, 4:16].astype(np.int64)
    scales = np.empty((rows, K // 256, 8), np.int64)
    mins = np.empty_like(scales)
    for j in range(8):   # gemv_q4k.cu:38-56
        if j < 4:
            scales[..., j] = sc[..., j] & 63
            mins[..., j] = sc[..., 4 + j] & 63
        else:
            scales[..., j] = (sc[..., 4 + j] & 0xF) | ((sc[..., j - 4] >> 6) << 4)
            mins[..., j] = (sc[..., 4 + j] >> 4) | ((sc[..., j] >> 6) << 4)
    qs = blocks[:, :, 16:144].astype(np.int64).reshape(rows, K // 256, 4, 32)
    q = np.stack([qs & 15, qs >> 4], axis=3).reshape(rows, K // 256, 8, 32)       # [row][block][sub-block][l]
    w = np.array([2.0 ** 24, 2.0 ** 16, 2.0 ** 8, 1.0], np.float32)
    tot = np.zeros((rows, 4), np.float32)                                           # one f32 accumulator per digit column
    for b in range(K // 256):
        xb = x[256 * b:256 * (b + 1)].astype(np.float32)
        s, inv = fixed_scale(np.abs(xb).max())
        dig = digits_msb_first(np.rint((xb * s).astype(np.float32)).astype(np.int64)).reshape(8, 32, 4)
        xs = xb.reshape(8, 32)
        sums = np.zeros(8, np.float32)
        for l in range(32):   # any fixed f32 order: the kernel adds 8 per lane, then the four lanes
            sums = sums + xs[:, l]
        digs = digits_msb_first(np.rint((sums * s * np.float32(0.015625)).astype(np.float32)).astype(np.int64))   # [8][4]
        c = np.einsum("rsl,slj->rsj", q[:, b], dig)                                  # exact integer dot products [row][sub-block][digit]
        acc = np.einsum("rs,rsj->rj", scales[:, b], c)                               # IMAD
        cm = np.einsum("rs,sj->rj", mins[:, b], digs)
        assert np.abs(acc).max() < 2 ** 31 and np.abs(c).max() < 2 ** 24
        da = (d[:, b] * inv)[:, None]
        dma = (dmin[:, b] * inv * np.float32(64.0))[:, None]
        term = (da * acc.astype(np.float32) - dma * cm.astype(np.float32)).astype(np.float32)
        tot = (tot + w[None, :] * term).astype(np.float32)
    return ((tot[:, 0] + tot[:, 1]) + (tot[:, 2] + tot[:, 3])).astype(np.float32)


@pytest.mark.parametrize("rows,K,xscale", [(64, 1024, 1.0), (32, 2048, 1e-4), (48, 512, 3e3)])
def test_integer_gemv_model_is_f32_accurate(rows, K, xscale):
    rng = np.random.default_rng(rows + K)
    raw = G.quantize(rng.standard_normal((rows, K), dtype=np.float32) * np.float32(0.02), G.Q4_K)
    x = (rng.standard_normal(K, dtype=np.float32) * np.float32(xscale)).astype(np.float32)
    x[5] *= np.float32(1e4)   # one outlier: the fixed-point image is per 256-weight super-block
    got = gemv_q4k_integer_model(raw, rows, K, x)
    ref = O.gemv_f64(G.Q4_K, raw, rows, K, x)
    wabs = np.abs(O.dequant(G.Q4_K, raw, rows * K).reshape(rows, K).astype(np.float64)) @ np.abs(x.astype(np.float64))
    span = np.abs(O.dequant(G.Q4_K, raw, rows * K).reshape(rows, K)[:, :32]).max(axis=1).astype(np.float64) * abs(float(x[5]))
    assert np.all(np.abs(got - ref) <= 3e-7 * wabs + 2e-6 * span + 1e-30)
