"""Block formats: the C oracle's dequantisation must equal an independent numpy
restatement bit for bit, and its GEMV must reproduce the reference's in-test
CPU reference (gemv_q4k_test.go:51-92, gemm_q4_test.go:14-85) on the reference's
own deterministic inputs."""
import numpy as np
import pytest

from oracle import oracle as O
from zerfoo_b200 import gguf as G
import refdata as R

QTYPES = [G.Q4_0, G.Q8_0, G.Q4_K, G.Q5_K, G.Q6_K]


@pytest.mark.parametrize("qt", QTYPES, ids=lambda q: G.TYPE_NAMES[q])
def test_dequant_bit_exact_quantized(qt):
    rng = np.random.default_rng(11 + qt)
    w = rng.standard_normal((48, 1024), dtype=np.float32) * np.float32(0.05)
    raw = G.quantize(w, qt)
    a = O.dequant(qt, raw, w.size)
    b = R.np_dequant(qt, raw)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # and the quantizer is a sane inverse
    assert np.abs(a.reshape(w.shape) - w).max() < 0.08 * np.abs(w).max()


@pytest.mark.parametrize("qt", QTYPES, ids=lambda q: G.TYPE_NAMES[q])
def test_dequant_bit_exact_random_bytes(qt):
    """Every bit pattern of the payload (all nibbles, high bits, 6-bit scale
    packings, negative int8 scales) with finite fp16 super-scales."""
    rng = np.random.default_rng(5 + qt)
    nblk = 512
    bb = G.BLOCK_BYTES[qt]
    raw = rng.integers(0, 256, size=(nblk, bb), dtype=np.uint8)
    scales = (rng.standard_normal((nblk, 2)) * 0.01).astype(np.float16).view(np.uint8).reshape(nblk, 4)
    if qt in (G.Q4_0, G.Q8_0):
        raw[:, 0:2] = scales[:, 0:2]
    elif qt in (G.Q4_K, G.Q5_K):
        raw[:, 0:4] = scales
    else:
        raw[:, 208:210] = scales[:, 0:2]
    n = nblk * G.BLOCK_ELEMS[qt]
    a = O.dequant(qt, raw, n)
    b = R.np_dequant(qt, raw)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("m,k", [(32, 256), (64, 512), (256, 1024), (512, 2048)])
def test_q4k_reference_test_vectors(m, k):
    """TestGemvQ4KF32_MultipleSizes shapes (gemv_q4k_test.go:319-329)."""
    w, x = R.q4k_test_vectors(m, k)
    raw = G.quantize_q4_k(w)
    deq = O.dequant(G.Q4_K, raw, m * k).reshape(m, k)
    # in-test CPU reference: sequential f32 sum of dequant*x
    ref = np.zeros(m, np.float32)
    for i in range(k):
        ref = (ref + deq[:, i] * x[i]).astype(np.float32)
    got = O.gemv(G.Q4_K, raw, m, k, x)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    exact = O.gemv_f64(G.Q4_K, raw, m, k, x)
    ok, worst = R.gemv_close(got, exact)
    assert ok, worst
    # quantisation itself is faithful to the sin() weights
    assert np.abs(deq - w).max() < 0.11


@pytest.mark.parametrize("m,k", [(16, 64), (100, 1152), (64, 4096)])
def test_q4_0_cpu_order(m, k):
    """q4DotRow order (internal/xblas/q4dot.go:10-50): per block
    sum(lo*x[p] ; hi*x[p+16]) interleaved, times scale, blocks left to right."""
    rng = np.random.default_rng(3)
    w = rng.standard_normal((m, k), dtype=np.float32) * np.float32(0.02)
    x = R.deterministic_data(k)
    raw = G.quantize_q4_0(w)
    got = O.gemv(G.Q4_0, raw, m, k, x)
    b = raw.reshape(m, k // 32, 18)
    d = np.ascontiguousarray(b[..., :2]).view(np.float16).astype(np.float32).reshape(m, k // 32)
    lo = (b[..., 2:] & 0xF).astype(np.float32) - 8
    hi = (b[..., 2:] >> 4).astype(np.float32) - 8
    xs = x.reshape(k // 32, 32)
    total = np.zeros(m, np.float32)
    for bi in range(k // 32):
        s = np.zeros(m, np.float32)
        for p in range(16):
            s = (s + lo[:, bi, p] * xs[bi, p]).astype(np.float32)
            s = (s + hi[:, bi, p] * xs[bi, p + 16]).astype(np.float32)
        total = (total + s * d[:, bi]).astype(np.float32)
    assert np.array_equal(got.view(np.uint32), total.view(np.uint32))
    ok, worst = R.gemv_close(got, O.gemv_f64(G.Q4_0, raw, m, k, x))
    assert ok, worst


@pytest.mark.parametrize("qt", QTYPES, ids=lambda q: G.TYPE_NAMES[q])
def test_gemv_linearity_and_f64(qt):
    rng = np.random.default_rng(17)
    m, k = 96, 1024
    raw = G.quantize(rng.standard_normal((m, k), dtype=np.float32) * np.float32(0.02), qt)
    x1 = rng.standard_normal(k, dtype=np.float32)
    x2 = rng.standard_normal(k, dtype=np.float32)
    y1, y2 = O.gemv_f64(qt, raw, m, k, x1), O.gemv_f64(qt, raw, m, k, x2)
    y12 = O.gemv_f64(qt, raw, m, k, (x1 + x2).astype(np.float32))
    assert np.abs(y12 - (y1 + y2)).max() < 1e-5
    ok, worst = R.gemv_close(O.gemv(qt, raw, m, k, x1), y1)
    assert ok, worst


def test_ragged_k_rejected():
    with pytest.raises(ValueError):
        O.gemv(G.Q4_K, np.zeros(144, np.uint8), 1, 1152, np.zeros(1152, np.float32))  # 1152 % 256 != 0
    with pytest.raises(ValueError):
        O.dequant(G.Q4_0, np.zeros(18, np.uint8), 31)


def test_separated_q4_layout_roundtrip():
    rng = np.random.default_rng(2)
    w = rng.standard_normal((8, 96), dtype=np.float32)
    raw = G.quantize_q4_0(w)
    sep, off = G.q4_0_to_separated(raw)
    n = 8 * 3
    assert off == (n * 2 + 15) // 16 * 16 and sep.size == off + n * 16
    assert np.array_equal(sep[: n * 2].reshape(n, 2), raw.reshape(n, 18)[:, :2])
    assert np.array_equal(sep[off:].reshape(n, 16), raw.reshape(n, 18)[:, 2:])


def test_q8_zerfoo36_layout():
    rng = np.random.default_rng(4)
    w = rng.standard_normal((4, 64), dtype=np.float32)
    raw = G.quantize_q8_0(w)
    z = G.q8_0_to_zerfoo36(raw).reshape(-1, 36)
    scale = np.ascontiguousarray(z[:, :4]).view(np.float32).reshape(-1)
    deq = (z[:, 4:].view(np.int8).astype(np.float32) * scale[:, None]).reshape(-1)
    assert np.array_equal(deq.view(np.uint32), O.dequant(G.Q8_0, raw, w.size).view(np.uint32))
