"""Deterministic input generators restated from the reference's own tests, so
parity tests here read like the reference's (SURVEY 8c "Fixtures")."""
import numpy as np


def q4k_test_vectors(m: int, k: int):
    """buildQ4KTestData inputs (internal/cuda/kernels/gemv_q4k_test.go:51-70):
    x[i] = (i%17-8)*0.05 ; w[row,i] = float32(sin((row*K+i)*0.03))*1.5."""
    x = ((np.arange(k) % 17 - 8).astype(np.float32) * np.float32(0.05)).astype(np.float32)
    idx = np.arange(m * k, dtype=np.float64).reshape(m, k)
    w = (np.sin(idx * 0.03).astype(np.float32) * np.float32(1.5)).astype(np.float32)
    return w, x


def deterministic_data(n: int) -> np.ndarray:
    """tests/parity/gpu_parity_ops_test.go:65-71: float32(i%97)*0.03-1.5."""
    return ((np.arange(n) % 97).astype(np.float32) * np.float32(0.03) - np.float32(1.5)).astype(np.float32)


def generate_f32_data(n: int) -> np.ndarray:
    """inference/load_gguf_test.go:25-31: sin(i*0.01)*0.02."""
    return (np.sin(np.arange(n, dtype=np.float64) * 0.01) * 0.02).astype(np.float32)


def gemv_close(got, want, abs_tol=1e-5, rel_tol=1e-4):
    """checkGemvRelError bound (internal/cuda/kernels/tolerance_test.go:48-51):
    |got-want| <= gemvReductionAbsTol + gemvReductionRelTol*|want|."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    err = np.abs(got - want)
    bound = abs_tol + rel_tol * np.abs(want)
    return bool(np.all(err <= bound)), float((err / bound).max())


# ---- independent numpy dequantizers (second restatement, used to cross-check
# the C oracle bit for bit) --------------------------------------------------

def _f16(b):
    return np.ascontiguousarray(b).view(np.float16).astype(np.float32)


def np_dequant(qtype: int, raw: np.ndarray) -> np.ndarray:
    from zerfoo_b200 import gguf as G
    raw = np.ascontiguousarray(raw).reshape(-1)
    if qtype == G.Q4_0:
        b = raw.reshape(-1, 18)
        d = _f16(b[:, :2]).reshape(-1, 1)
        q = b[:, 2:]
        lo = (q & 0xF).astype(np.int32) - 8
        hi = (q >> 4).astype(np.int32) - 8
        return (np.concatenate([lo, hi], axis=1).astype(np.float32) * d).reshape(-1)
    if qtype == G.Q8_0:
        b = raw.reshape(-1, 34)
        d = _f16(b[:, :2]).reshape(-1, 1)
        return (b[:, 2:].view(np.int8).astype(np.float32) * d).reshape(-1)
    if qtype in (G.Q4_K, G.Q5_K):
        bb = G.BLOCK_BYTES[qtype]
        b = raw.reshape(-1, bb)
        n = b.shape[0]
        d = _f16(b[:, 0:2]).reshape(n, 1)
        dmin = _f16(b[:, 2:4]).reshape(n, 1)
        sc = b[:, 4:16]
        scales = np.empty((n, 8), np.uint8)
        mins = np.empty((n, 8), np.uint8)
        scales[:, :4] = sc[:, 0:4] & 63
        mins[:, :4] = sc[:, 4:8] & 63
        scales[:, 4:] = (sc[:, 8:12] & 0xF) | ((sc[:, 0:4] >> 6) << 4)
        mins[:, 4:] = (sc[:, 8:12] >> 4) | ((sc[:, 4:8] >> 6) << 4)
        s = d * scales.astype(np.float32)
        m = dmin * mins.astype(np.float32)
        ql = b[:, 16:144].reshape(n, 4, 32)
        q = np.stack([ql & 0xF, ql >> 4], axis=2).astype(np.int32)  # [n, 4, 2, 32]
        if qtype == G.Q5_K:
            qh = b[:, 144:176].reshape(n, 1, 1, 32).astype(np.int32)
            shift = (2 * np.arange(4).reshape(1, 4, 1, 1) + np.arange(2).reshape(1, 1, 2, 1))
            q = q | (((qh >> shift) & 1) << 4)
        q = q.reshape(n, 8, 32).astype(np.float32)
        return (s[:, :, None] * q - m[:, :, None]).astype(np.float32).reshape(-1)
    if qtype == G.Q6_K:
        b = raw.reshape(-1, 210)
        n = b.shape[0]
        ql = b[:, :128].reshape(n, 2, 2, 32).astype(np.int32)     # [half, which32, l]
        qh = b[:, 128:192].reshape(n, 2, 32).astype(np.int32)
        sc = b[:, 192:208].view(np.int8).astype(np.float32)       # [n, 16]
        d = _f16(b[:, 208:210]).reshape(n, 1)
        q1 = (ql[:, :, 0] & 0xF) | ((qh & 3) << 4)
        q2 = (ql[:, :, 1] & 0xF) | (((qh >> 2) & 3) << 4)
        q3 = (ql[:, :, 0] >> 4) | (((qh >> 4) & 3) << 4)
        q4 = (ql[:, :, 1] >> 4) | (((qh >> 6) & 3) << 4)
        q = np.stack([q1, q2, q3, q4], axis=2) - 32               # [n, half, quarter, 32]
        q = q.reshape(n, 16, 16).astype(np.float32)
        return ((d * sc)[:, :, None] * q).astype(np.float32).reshape(-1)
    raise ValueError(qtype)
