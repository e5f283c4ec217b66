"""Parity of the tcgen05/TMEM batched dequant-GEMM (zb_gemm_tc_f32) against the CPU oracle.

The kernel's contract: weights = bf16(round-to-nearest-even of the bit-exact f32 dequant), activations = bf16 hi
(+ bf16 lo residual when split), f32 accumulation in tensor memory.  Two checks per case:
  1. against an f64 contraction of exactly those rounded operands: only accumulation order may differ
     -> |err| <= 1e-6 + 1e-5 * sum_k |w_k x_k|;
  2. against the unrounded f32-weight reference (what the batch-1 GEMV computes): the bf16 rounding noise
     -> |err| <= 2^-7 * sum_k |w_k x_k|  (each product carries <= 2^-9 weight + 2^-9 / 2^-17 activation error)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O
from zerfoo_b200 import gguf as G

torch = pytest.importorskip("torch")

KQ = [G.Q4_K, G.Q5_K, G.Q6_K]


def bf16_round(a):
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).astype(np.uint32).view(np.float32)


CASES = [  # (tokens, rows, K)
    (16, 128, 256), (32, 256, 1024), (32, 4096, 4096), (17, 300, 512), (64, 1024, 3072), (128, 512, 2048), (256, 256, 1024),
    (300, 384, 768), (32, 1000, 14336),
]


@pytest.mark.parametrize("qt", KQ, ids=[G.TYPE_NAMES[t] for t in KQ])
@pytest.mark.parametrize("case", CASES, ids=[f"T{t}_N{n}_K{k}" for t, n, k in CASES])
def test_gemm_tc_matches_oracle(qt, case):
    from zerfoo_b200 import kernels as K
    T, N, Kd = case
    rng = np.random.default_rng(T * 131 + N + Kd)
    raw = G.quantize(rng.standard_normal((N, Kd), dtype=np.float32) * np.float32(0.02), qt)
    x = rng.standard_normal((T, Kd), dtype=np.float32)
    W = O.dequant(qt, raw, N * Kd).reshape(N, Kd)
    split = T <= 64
    got = K.gemm_tc(K.StreamWeight(qt, raw, N, Kd), torch.from_numpy(x).cuda()).cpu().numpy()
    Wb = bf16_round(W).astype(np.float64)
    xh = bf16_round(x)
    xe = xh.astype(np.float64) + (bf16_round(x - xh).astype(np.float64) if split else 0.0)
    ref_rounded = xe @ Wb.T
    mag = np.abs(x).astype(np.float64) @ np.abs(W).astype(np.float64).T
    err = np.abs(got - ref_rounded)
    assert (err <= 1e-6 + 1e-5 * mag).all(), f"max err vs rounded-operand oracle {err.max():.3e} (bound {(1e-6 + 1e-5 * mag).min():.3e})"
    ref = x.astype(np.float64) @ W.astype(np.float64).T
    assert (np.abs(got - ref) <= 2.0 ** -7 * mag + 1e-6).all()


def test_gemm_tc_agrees_with_stream_gemv_rows():
    """Same weights, B=32: every row of the batched result agrees with the batch-1 streamed GEMV within bf16 noise."""
    from zerfoo_b200 import kernels as K
    rng = np.random.default_rng(4)
    N, Kd, T = 512, 2048, 32
    raw = G.quantize(rng.standard_normal((N, Kd), dtype=np.float32) * np.float32(0.02), G.Q4_K)
    x = rng.standard_normal((T, Kd), dtype=np.float32)
    W = K.StreamWeight(G.Q4_K, raw, N, Kd)
    yb = K.gemm_tc(W, torch.from_numpy(x).cuda()).cpu().numpy()
    for t in (0, 7, 31):
        yv = K.gemv_stream(W, torch.from_numpy(x[t]).cuda()).cpu().numpy()
        assert np.abs(yb[t] - yv).max() <= 5e-3 * np.abs(yv).max() + 1e-4
