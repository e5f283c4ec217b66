"""Parity of the tensor-core batch-1 GEMV (csrc/gemv_mma.cu, through zb_gemv_mma_f32) against the CPU oracle.

Same bar as the CUDA-core streamed GEMV: |got - ref| <= 1e-5 + 1e-4*|ref| (internal/cuda/kernels/tolerance_test.go:48-51)
against an f64-accumulated oracle; results must be bit-identical run to run (fixed summation order, also across the CTAs
that share a row tile)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O
from zerfoo_b200 import gguf as G

torch = pytest.importorskip("torch")

TYPES = [G.Q4_K, G.Q6_K, G.Q4_0, G.Q5_K]


@pytest.fixture(scope="module")
def K():
    from zerfoo_b200 import kernels
    return kernels


def close(got, ref, atol=1e-5, rtol=1e-4):
    bad = np.abs(got - ref) > atol + rtol * np.abs(ref)
    assert not bad.any(), f"{bad.sum()} of {bad.size} outside tolerance; max err {np.abs(got - ref).max():.3e}"


def mk(qt, m, k, seed=0, sigma=0.02, xscale=1.0):
    rng = np.random.default_rng(seed)
    raw = G.quantize(rng.standard_normal((m, k), dtype=np.float32) * np.float32(sigma), qt)
    x = (rng.standard_normal(k, dtype=np.float32) * np.float32(xscale)).astype(np.float32)
    return raw, x


SHAPES = [(16, 256), (64, 256), (8, 512), (40, 512), (1000, 1024), (5120, 3072), (3072, 3072), (3072, 8192), (1031, 3072), (257, 2048),
          (16384, 3072), (4096, 4096)]


@pytest.mark.parametrize("qt", TYPES, ids=[G.TYPE_NAMES[t] for t in TYPES])
@pytest.mark.parametrize("shape", SHAPES, ids=[f"{m}x{k}" for m, k in SHAPES])
def test_mma_gemv_matches_oracle(K, qt, shape):
    m, k = shape
    raw, x = mk(qt, m, k, seed=m * 7 + k)
    w = K.MmaWeight(qt, raw, m, k)
    xd = torch.from_numpy(x).cuda()
    y = K.gemv_mma(w, xd).cpu().numpy()
    close(y, O.gemv_f64(qt, raw, m, k, x))
    for _ in range(3):   # tickets self-reset; PDL launch; bit-identical
        y2 = K.gemv_mma(w, xd, pdl=True).cpu().numpy()
        assert np.array_equal(y, y2)


C1_SHAPES = [(1536, 1152), (1152, 1024), (1152, 6912), (2304, 1152), (65536, 1152)]


@pytest.mark.parametrize("shape", C1_SHAPES, ids=[f"{m}x{k}" for m, k in C1_SHAPES])
def test_mma_gemv_q4_0_gemma_shapes(K, shape):
    """C1 (Gemma-3-1B shape): K = 1152 is 4.5 x-blocks of 256 -- the last one is half empty -- and the tied 262144-row head."""
    m, k = shape
    raw, x = mk(G.Q4_0, m, k, seed=m + 3 * k)
    w = K.MmaWeight(G.Q4_0, raw, m, k)
    y = K.gemv_mma(w, torch.from_numpy(x).cuda()).cpu().numpy()
    close(y, O.gemv_f64(G.Q4_0, raw, m, k, x))
    if m % 2 == 0:
        yp = K.gemv_mma(w, torch.from_numpy(x).cuda(), swiglu_pairs=True).cpu().numpy()
        full = O.gemv_f64(G.Q4_0, raw, m, k, x).astype(np.float32)
        close(yp, O.swiglu(full[0::2], full[1::2]), atol=2e-5, rtol=2e-4)


@pytest.mark.parametrize("xscale", [1e-6, 1e-3, 37.0, 3e4])
def test_mma_gemv_activation_range(K, xscale):
    """The three-term fp16 split is scaled by a power of two from max|x|: tiny and large activations keep the bar;
    one outlier 1e4 x larger than the rest must not cost the small elements their precision."""
    m, k = 512, 1024
    raw, x = mk(G.Q4_K, m, k, seed=3, xscale=xscale)
    w = K.MmaWeight(G.Q4_K, raw, m, k)
    y = K.gemv_mma(w, torch.from_numpy(x).cuda()).cpu().numpy()
    ref = O.gemv_f64(G.Q4_K, raw, m, k, x)
    close(y, ref, atol=1e-5 * max(1.0, xscale))
    # One activation 1e4 x larger than the rest.  The products are (unsigned nibble) x (activation) -- the block minimum is a
    # separate term -- so rounding scales with big = (quantisation range of the sub-block) x (largest activation), not with |ref|.
    # Integer path (default): products and sums are exact, only the final f32 combination rounds: <= 2e-6 of big.
    # f16 path (ZB_MMA_I8=0), measured on B200: inside one HMMA the addends are aligned to the largest product and truncated
    # ~17 bits below it: ~5e-5 of big; the bar is the reference's 1e-4 taken relative to big.
    x2 = x.copy()
    x2[17] = np.float32(1e4 * xscale)
    y = K.gemv_mma(w, torch.from_numpy(x2).cuda()).cpu().numpy()
    ref = O.gemv_f64(G.Q4_K, raw, m, k, x2)
    wd = O.dequant(G.Q4_K, raw, m * k).reshape(m, k)[:, :32].astype(np.float64)
    big = (wd.max(axis=1) - wd.min(axis=1)) * float(x2[17])
    rel_big = 2e-6 if os.environ.get("ZB_MMA_I8", "1") != "0" else 1e-4
    assert np.all(np.abs(y - ref) <= 1e-5 * max(1.0, xscale) + 1e-4 * np.abs(ref) + rel_big * big)


def test_mma_gemv_zero_input(K):
    m, k = 64, 512
    raw, x = mk(G.Q4_K, m, k, seed=1)
    w = K.MmaWeight(G.Q4_K, raw, m, k)
    y = K.gemv_mma(w, torch.zeros(k, device="cuda")).cpu().numpy()
    assert np.array_equal(y, np.zeros(m, np.float32))


@pytest.mark.parametrize("qt", TYPES, ids=[G.TYPE_NAMES[t] for t in TYPES])
def test_mma_gemv_prologues(K, qt):
    m, k, eps = 768, 1024, 1e-6
    raw, _ = mk(qt, m, k, seed=3)
    rng = np.random.default_rng(11)
    a = rng.standard_normal(k, dtype=np.float32)
    r = rng.standard_normal(k, dtype=np.float32)
    w1 = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32)
    w2 = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32)
    W = K.MmaWeight(qt, raw, m, k)
    d = lambda v: torch.from_numpy(np.ascontiguousarray(v)).cuda()
    ref_gemv = lambda x: O.gemv_f64(qt, raw, m, k, x.astype(np.float32))
    y = K.gemv_mma(W, d(a), w2=d(w2), eps=eps).cpu().numpy()
    close(y, ref_gemv(O.rmsnorm(a, w2, eps)), atol=2e-5)
    so = torch.zeros(k, device="cuda")
    y = K.gemv_mma(W, d(a), r=d(r), w2=d(w2), sum_out=so, eps=eps).cpu().numpy()
    normed, s = O.add_rmsnorm(a, r, w2, eps)
    close(y, ref_gemv(normed), atol=2e-5)
    assert np.array_equal(so.cpu().numpy(), s)
    so.zero_()
    y = K.gemv_mma(W, d(a), w1=d(w1), r=d(r), w2=d(w2), sum_out=so, eps=eps).cpu().numpy()
    mid = O.norm_add(a, w1, r, eps)
    close(y, ref_gemv(O.rmsnorm(mid, w2, eps)), atol=2e-5)
    gu = rng.standard_normal(2 * k, dtype=np.float32)
    y = K.gemv_mma(W, d(gu), swiglu=True).cpu().numpy()
    close(y, ref_gemv(O.swiglu(gu[:k], gu[k:])), atol=2e-5)


@pytest.mark.parametrize("qt", TYPES, ids=[G.TYPE_NAMES[t] for t in TYPES])
@pytest.mark.parametrize("shape", [(64, 256), (1024, 1024), (16384, 3072), (1030, 512)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_mma_gemv_swiglu_pairs(K, shape, qt):
    m, k = shape
    raw, x = mk(qt, m, k, seed=m + k)
    W = K.MmaWeight(qt, raw, m, k)
    y = K.gemv_mma(W, torch.from_numpy(x).cuda(), swiglu_pairs=True).cpu().numpy()
    full = O.gemv_f64(qt, raw, m, k, x).astype(np.float32)
    ref = O.swiglu(full[0::2], full[1::2])
    assert y.shape == (m // 2,)
    close(y, ref, atol=2e-5, rtol=2e-4)


@pytest.mark.parametrize("qt", TYPES, ids=[G.TYPE_NAMES[t] for t in TYPES])
def test_mma_matches_cuda_core_kernel(K, qt):
    """Both batch-1 kernels implement the same operator: they must agree far inside the oracle bar."""
    m, k = 3072, 3072
    raw, x = mk(qt, m, k, seed=42)
    xd = torch.from_numpy(x).cuda()
    y1 = K.gemv_stream(K.StreamWeight(qt, raw, m, k), xd).cpu().numpy()
    y2 = K.gemv_mma(K.MmaWeight(qt, raw, m, k), xd).cpu().numpy()
    assert np.abs(y1 - y2).max() <= 2e-6 + 2e-5 * np.abs(y1).max()


@pytest.mark.parametrize("qt", [G.Q4_K, G.Q5_K, G.Q6_K, G.Q4_0], ids=["Q4_K", "Q5_K", "Q6_K", "Q4_0"])
def test_mma_integer_path_is_f32_accurate(K, qt):
    """The integer tensor path computes every dot product exactly on a 32-bit fixed-point image of x; what is left is f32
    rounding of the per-super-block combination: the error stays within a few f32 ulps of sum |w x| (here: 3e-7 of it)."""
    if os.environ.get("ZB_MMA_I8", "1") == "0":
        pytest.skip("f16 tensor path selected")
    m, k = 1024, 4096
    raw, x = mk(qt, m, k, seed=77)
    y = K.gemv_mma(K.MmaWeight(qt, raw, m, k), torch.from_numpy(x).cuda()).cpu().numpy()
    ref = O.gemv_f64(qt, raw, m, k, x)
    wabs = np.abs(O.dequant(qt, raw, m * k).reshape(m, k).astype(np.float64)) @ np.abs(x.astype(np.float64))
    assert np.all(np.abs(y - ref) <= 3e-7 * wabs + 1e-7), float(np.max(np.abs(y - ref) / wabs))


def test_mma_gemv_reads_replicated_input(K):
    """zb_prologue.a_replicas: CTA c reads copy c % n of the input vector; identical copies give the identical result."""
    m, k = 2048, 1024
    raw, x = mk(G.Q4_K, m, k, seed=9)
    w = K.MmaWeight(G.Q4_K, raw, m, k)
    xd = torch.from_numpy(x).cuda()
    y0 = K.gemv_mma(w, xd).cpu().numpy()
    y4 = K.gemv_mma(w, xd.repeat(4).contiguous(), a_replicas=4, a_replica_stride=k).cpu().numpy()
    assert np.array_equal(y0, y4)


@pytest.mark.parametrize("pairs", [False, True], ids=["plain", "swiglu_pairs"])
def test_mma_gemv_expert_indirection(K, pairs):
    """MoE: a stack of experts in one block-tile buffer, slot k multiplies expert expert_sel[k] with its own input and writes its
    own output slice (layers/core/moe.go:110-146; the engine's expert GEMVs); a negative id zeroes the slot (expert on another rank)."""
    E, m, k = 4, 512, 1024
    rng = np.random.default_rng(2)
    raws = [G.quantize(rng.standard_normal((m, k), dtype=np.float32) * np.float32(0.02), G.Q4_K) for _ in range(E)]
    W = K.MmaWeight(G.Q4_K, np.concatenate([np.asarray(r).view(np.uint8).reshape(-1) for r in raws]), E * m, k, experts=E)
    x = rng.standard_normal((2, k), dtype=np.float32)
    sel = torch.tensor([3, 1], dtype=torch.int32, device="cuda")
    nout = m // 2 if pairs else m
    y = K.gemv_mma(W, torch.from_numpy(x).cuda(), sel=sel, a_slot_stride=k, swiglu_pairs=pairs).cpu().numpy().reshape(2, nout)
    for slot, ex in enumerate((3, 1)):
        full = O.gemv_f64(G.Q4_K, raws[ex], m, k, x[slot])
        if pairs:
            f32 = full.astype(np.float32)
            close(y[slot], O.swiglu(f32[0::2], f32[1::2]), atol=2e-5, rtol=2e-4)
        else:
            close(y[slot], full)
    sel2 = torch.tensor([-1, 2], dtype=torch.int32, device="cuda")
    y2 = K.gemv_mma(W, torch.from_numpy(x).cuda(), sel=sel2, a_slot_stride=k, swiglu_pairs=pairs).cpu().numpy().reshape(2, nout)
    assert np.array_equal(y2[0], np.zeros(nout, np.float32))
    y3 = K.gemv_mma(W, torch.from_numpy(x).cuda(), sel=sel2, a_slot_stride=k, swiglu_pairs=pairs, pdl=True).cpu().numpy().reshape(2, nout)
    assert np.array_equal(y2, y3)
