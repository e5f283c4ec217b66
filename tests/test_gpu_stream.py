"""Parity of the B200 streamed decode kernels (through the zb_* C ABI) against the CPU oracle:
TMA-streamed fused dequant-GEMV (all five block formats, the BASELINE shapes, every fused prologue,
MoE expert indirection) and the fused decode-attention stage.

GEMV tolerance is the reference's own: |got - ref| <= gemvReductionAbsTol + gemvReductionRelTol*|ref|
= 1e-5 + 1e-4*|ref| (internal/cuda/kernels/tolerance_test.go:48-51), against an f64-accumulated oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O
from zerfoo_b200 import gguf as G

torch = pytest.importorskip("torch")

ALL = [G.Q4_0, G.Q8_0, G.Q4_K, G.Q5_K, G.Q6_K]
NAMES = {t: G.TYPE_NAMES[t] for t in ALL}


@pytest.fixture(scope="module")
def K():
    from zerfoo_b200 import kernels
    return kernels


def close(got, ref, atol=1e-5, rtol=1e-4):
    bad = np.abs(got - ref) > atol + rtol * np.abs(ref)
    assert not bad.any(), f"{bad.sum()} of {bad.size} outside tolerance; max err {np.abs(got - ref).max():.3e}"


def mk(qt, m, k, seed=0, sigma=0.02):
    rng = np.random.default_rng(seed)
    w = rng.standard_normal((m, k), dtype=np.float32) * np.float32(sigma)
    raw = G.quantize(w, qt)
    x = rng.standard_normal(k, dtype=np.float32)
    return raw, x


SHAPES = [  # (rows, K): BASELINE shapes + ragged / tiny cases
    (64, 256), (8, 512), (1000, 1024), (1536, 1152), (1152, 1024), (2304, 1152), (1152, 6912),
    (5120, 3072), (3072, 8192), (4096, 4096), (1031, 3072), (4096, 14336), (257, 2048),
]


@pytest.mark.parametrize("qt", ALL, ids=[NAMES[t] for t in ALL])
@pytest.mark.parametrize("shape", SHAPES, ids=[f"{m}x{k}" for m, k in SHAPES])
def test_stream_gemv_matches_oracle(K, qt, shape):
    m, k = shape
    if qt in (G.Q4_K, G.Q5_K, G.Q6_K) and k % 256:
        pytest.skip("K-quants need K % 256 == 0 (model/gguf/loader.go:296-299)")
    raw, x = mk(qt, m, k, seed=m * 7 + k)
    w = K.StreamWeight(qt, raw, m, k)
    y = K.gemv_stream(w, torch.from_numpy(x).cuda()).cpu().numpy()
    close(y, O.gemv_f64(qt, raw, m, k, x))
    # same result with programmatic dependent launch enabled, and bit-identical run to run
    y2 = K.gemv_stream(w, torch.from_numpy(x).cuda(), pdl=True).cpu().numpy()
    assert np.array_equal(y, y2)


@pytest.mark.parametrize("qt", [G.Q4_0, G.Q6_K], ids=["Q4_0", "Q6_K"])
def test_stream_gemv_large_head(K, qt):
    """lm_head-sized matrix (the two-CTA-per-SM path): C1 262144 x 1152 for Q4_0, C2 128256 x 3072 slice for Q6_K."""
    m, k = (65536, 1152) if qt == G.Q4_0 else (40000, 3072)
    raw, x = mk(qt, m, k, seed=5)
    w = K.StreamWeight(qt, raw, m, k)
    y = K.gemv_stream(w, torch.from_numpy(x).cuda()).cpu().numpy()
    close(y, O.gemv_f64(qt, raw, m, k, x))


@pytest.mark.parametrize("qt", ALL, ids=[NAMES[t] for t in ALL])
def test_stream_gemv_prologues(K, qt):
    m, k, eps = 768, 1024, 1e-6
    raw, _ = mk(qt, m, k, seed=3)
    rng = np.random.default_rng(11)
    a = rng.standard_normal(k, dtype=np.float32)
    r = rng.standard_normal(k, dtype=np.float32)
    w1 = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32)
    w2 = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32)
    W = K.StreamWeight(qt, raw, m, k)
    d = lambda v: torch.from_numpy(np.ascontiguousarray(v)).cuda()
    ref_gemv = lambda x: O.gemv_f64(qt, raw, m, k, x.astype(np.float32))
    # RMSNorm prologue (FusedRMSNormGPU + MatMul)
    y = K.gemv_stream(W, d(a), w2=d(w2), eps=eps).cpu().numpy()
    close(y, ref_gemv(O.rmsnorm(a, w2, eps)), atol=2e-5)
    # Add + RMSNorm (fusedAddRMSNormNode): also returns the residual sum
    so = torch.zeros(k, device="cuda")
    y = K.gemv_stream(W, d(a), r=d(r), w2=d(w2), sum_out=so, eps=eps).cpu().numpy()
    normed, s = O.add_rmsnorm(a, r, w2, eps)
    close(y, ref_gemv(normed), atol=2e-5)
    assert np.array_equal(so.cpu().numpy(), s)
    # Norm + Add + Norm (Gemma 3: post-attention norm, residual, pre-FFN norm)
    so.zero_()
    y = K.gemv_stream(W, d(a), w1=d(w1), r=d(r), w2=d(w2), sum_out=so, eps=eps).cpu().numpy()
    mid = O.norm_add(a, w1, r, eps)
    close(y, ref_gemv(O.rmsnorm(mid, w2, eps)), atol=2e-5)
    np.testing.assert_allclose(so.cpu().numpy(), mid, rtol=1e-6, atol=1e-6)
    # SwiGLU prologue (GPUFusedSwiGLU + down projection)
    gu = rng.standard_normal(2 * k, dtype=np.float32)
    y = K.gemv_stream(W, d(gu), swiglu=True).cpu().numpy()
    close(y, ref_gemv(O.swiglu(gu[:k], gu[k:])), atol=2e-5)
    # MoE combine prologue: a = sum_k w_k * y_k, then residual add (moe.go:470-479)
    ys = rng.standard_normal((2, k), dtype=np.float32)
    mw = np.array([0.7, 0.3], np.float32)
    comb = (np.float32(0) + ys[0] * mw[0]) + ys[1] * mw[1]
    y = K.gemv_stream(W, d(ys), mix_w=d(mw), mix_n=2, mix_stride=k, r=d(r), w2=d(w2), eps=eps).cpu().numpy()
    close(y, ref_gemv(O.rmsnorm((comb + r).astype(np.float32), w2, eps)), atol=2e-5)


@pytest.mark.parametrize("qt", ALL, ids=[NAMES[t] for t in ALL])
@pytest.mark.parametrize("shape", [(1024, 512), (16384, 3072), (13824, 1152), (96, 256)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_stream_gemv_swiglu_pair_epilogue(K, qt, shape):
    """Rows interleaved (gate_i, up_i): the epilogue writes silu(gate_i)*up_i (GPUFusedSwiGLU fused into the gate|up MatMul)."""
    m, k = shape
    if qt in (G.Q4_K, G.Q5_K, G.Q6_K) and k % 256:
        pytest.skip("K-quants need K % 256 == 0")
    raw, x = mk(qt, m, k, seed=m + k)
    W = K.StreamWeight(qt, raw, m, k)
    y = K.gemv_stream(W, torch.from_numpy(x).cuda(), swiglu_pairs=True).cpu().numpy()
    full = O.gemv_f64(qt, raw, m, k, x).astype(np.float32)
    ref = O.swiglu(full[0::2], full[1::2])
    assert y.shape == (m // 2,)
    close(y, ref, atol=2e-5, rtol=2e-4)


def test_stream_gemv_expert_indirection(K):
    E, m, k = 4, 512, 1024
    rng = np.random.default_rng(2)
    raws = [G.quantize(rng.standard_normal((m, k), dtype=np.float32) * np.float32(0.02), G.Q4_K) for _ in range(E)]
    W = K.StreamWeight(G.Q4_K, np.concatenate([np.asarray(r).view(np.uint8).reshape(-1) for r in raws]), E * m, k, experts=E)
    x = rng.standard_normal((2, k), dtype=np.float32)
    sel = torch.tensor([3, 1], dtype=torch.int32, device="cuda")
    y = K.gemv_stream(W, torch.from_numpy(x).cuda(), sel=sel, a_slot_stride=k).cpu().numpy().reshape(2, m)
    close(y[0], O.gemv_f64(G.Q4_K, raws[3], m, k, x[0]))
    close(y[1], O.gemv_f64(G.Q4_K, raws[1], m, k, x[1]))


def test_stream_layout_is_a_pure_byte_move(K):
    """Bit-exact dequantisation must survive the repack: every stream layout is a permutation of the GGUF bytes."""
    for qt in ALL:
        raw, _ = mk(qt, 16, 512, seed=9)
        W = K.StreamWeight(qt, raw, 16, 512)
        rawb = np.asarray(raw).view(np.uint8).reshape(-1)
        got = np.sort(np.concatenate([W.main.cpu().numpy()[: 16 * 512 // G.BLOCK_ELEMS[qt] * (G.BLOCK_BYTES[qt] - (2 if W.aux is not None else 0))],
                                      W.aux.cpu().numpy()[: 16 * 512 // G.BLOCK_ELEMS[qt] * 2] if W.aux is not None else np.zeros(0, np.uint8)]))
        assert np.array_equal(got, np.sort(rawb))


ATTN = [  # (hd, n_q, n_kv, max_seq, qk_norm)
    (128, 24, 8, 256, False), (256, 4, 1, 128, True), (64, 8, 4, 96, False), (32, 8, 8, 64, False), (128, 32, 8, 160, False),
    (128, 64, 8, 64, False), (128, 4, 2, 512, True),
]


@pytest.mark.parametrize("cfg", ATTN, ids=[f"hd{c[0]}_q{c[1]}_kv{c[2]}_s{c[3]}" for c in ATTN])
def test_decode_attention_stage(K, cfg):
    hd, nq, nkv, max_seq, qkn = cfg
    rng = np.random.default_rng(hd + nq)
    eps, chunk = 1e-6, 32
    splits = (max_seq + chunk - 1) // chunk
    cs, sn = O.rope_tables(max_seq, hd, 1e4)
    wq = (1 + 0.1 * rng.standard_normal(hd)).astype(np.float32) if qkn else None
    wk = (1 + 0.1 * rng.standard_normal(hd)).astype(np.float32) if qkn else None
    d = lambda v: None if v is None else torch.from_numpy(np.ascontiguousarray(v)).cuda()
    kc = torch.zeros(nkv * max_seq * hd, device="cuda"); vc = torch.zeros_like(kc)
    out = torch.zeros(nq * hd, device="cuda")
    part_o = torch.zeros(nq * splits * hd, device="cuda"); part_ml = torch.zeros(2 * nq * splits, device="cuda")
    ticket = torch.zeros(nkv, dtype=torch.int32, device="cuda")
    pos = torch.zeros(1, dtype=torch.int32, device="cuda")
    dcs, dsn, dwq, dwk = d(cs), d(sn), d(wq), d(wk)
    Kref = np.zeros((max_seq, nkv * hd), np.float32); Vref = np.zeros_like(Kref)
    steps = list(range(0, min(max_seq, 70))) + ([max_seq - 1] if max_seq > 70 else [])
    for t in steps:
        if t == max_seq - 1 and t > 70:      # jump: fill the cache rows in between with random K/V on both sides
            fill = rng.standard_normal((max_seq, nkv * hd), dtype=np.float32)
            Kref[70:t] = fill[70:t]; Vref[70:t] = fill[70:t] * 0.5
            kc.copy_(d(np.ascontiguousarray(Kref.reshape(max_seq, nkv, hd).transpose(1, 0, 2)).reshape(-1)))
            vc.copy_(d(np.ascontiguousarray(Vref.reshape(max_seq, nkv, hd).transpose(1, 0, 2)).reshape(-1)))
        qkv = rng.standard_normal((nq + 2 * nkv) * hd, dtype=np.float32)
        pos.fill_(t)
        K.decode_attn(d(qkv), dwq, dwk, dcs, dsn, pos, kc, vc, out, part_o, part_ml, ticket, eps, hd, nq, nkv, max_seq, chunk, splits)
        q = qkv[: nq * hd].reshape(nq, hd).copy(); k = qkv[nq * hd:(nq + nkv) * hd].reshape(nkv, hd).copy()
        v = qkv[(nq + nkv) * hd:].reshape(nkv, hd)
        if qkn:
            q = np.stack([O.rmsnorm(r, wq, eps) for r in q]); k = np.stack([O.rmsnorm(r, wk, eps) for r in k])
        q = np.stack([O.rope(r, cs[t], sn[t]) for r in q]); k = np.stack([O.rope(r, cs[t], sn[t]) for r in k])
        Kref[t] = k.reshape(-1); Vref[t] = v.reshape(-1)
        ref = O.attn_decode(q.astype(np.float32), Kref, Vref, nkv, t + 1)
        got = out.cpu().numpy()
        assert np.abs(got - ref.reshape(-1)).max() <= 2e-5 * max(1.0, np.abs(ref).max()), f"pos {t}"
        assert int(ticket.abs().sum().item()) == 0
    gk = kc.cpu().numpy().reshape(nkv, max_seq, hd).transpose(1, 0, 2).reshape(max_seq, -1)
    np.testing.assert_allclose(gk[:70], Kref[:70], rtol=1e-6, atol=1e-6)
