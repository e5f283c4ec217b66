"""Chunked prompt prefill (tcgen05 dequant-GEMMs + causal chunk attention over the decode cache) against the CPU oracle.

Same tolerance as the batched decode path (bf16 weight / activation tiles): relative L2 error of the
last-token logits <= 2e-2, max error <= 5e-2 * max|logit|.  The K/V rows the chunk attention appends and the decode
steps that follow run on the same cache, so decode after a chunked prefill must track the oracle too.  The chunk
attention kernel itself is f32 and is checked at 2e-5 against a numpy causal softmax."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import modelzoo as Z
from oracle import oracle as O

torch = pytest.importorskip("torch")


def _close(got, ref):
    err = got - ref
    assert np.linalg.norm(err) <= 2e-2 * np.linalg.norm(ref), np.linalg.norm(err) / np.linalg.norm(ref)
    assert np.abs(err).max() <= 5e-2 * np.abs(ref).max()
    return np.abs(err).max()


@pytest.mark.parametrize("kind,n", [("llama_q4_k_m", 40), ("mistral_q5_k_m", 23), ("llama_q4_k_m", 200)])
def test_chunked_prefill_matches_oracle(kind, n):
    from zerfoo_b200 import engine as E
    path = Z.path(kind)
    g = E.load_file(path)
    om = O.Model(path)
    rng = np.random.default_rng(n)
    prompt = [int(t) for t in rng.integers(1, g.info.vocab, size=n)]
    ref = None
    for t in prompt:
        ref = om.forward(t)
    first, ms = g.prefill_chunked(prompt)
    assert ms > 0 and g.position == n
    e = _close(g.logits(), ref)
    srt = np.sort(ref)
    if srt[-1] - srt[-2] > 4 * e:
        assert first == O.argmax(ref)
    # K/V rows written by the chunk kernels == the oracle's cache rows up to bf16 noise of the projections
    k, v = g.kv(0, n)
    ok, ov = om.kv(0, n)
    assert np.abs(k - ok).max() <= 5e-2 * np.abs(ok).max() and np.abs(v - ov).max() <= 5e-2 * np.abs(ov).max()
    # decode continues on the same cache (exact GEMV path on top of the bf16-prefilled prefix)
    tok = O.argmax(ref)
    for _ in range(3):
        ref = om.forward(tok)
        nxt = g.decode_step(tok)
        e = _close(g.logits(), ref)
        tok = O.argmax(ref)
    # a second chunked call appends at the current position
    more = [int(t) for t in rng.integers(1, g.info.vocab, size=5)]
    for t in more:
        ref = om.forward(t)
    g.prefill_chunked(more)
    _close(g.logits(), ref)
    assert g.position == n + 3 + 5
    g.close()
    om.close()


def test_chunked_prefill_spans_chunks():
    """More than one 256-token chunk: the second chunk attends to the first through the cache."""
    from zerfoo_b200 import engine as E
    path = Z.path("preset:c2:2:512")
    g = E.load_file(path)
    om = O.Model(path)
    rng = np.random.default_rng(7)
    n = 300
    prompt = [int(t) for t in rng.integers(1, g.info.vocab, size=n)]
    for t in prompt:
        ref = om.forward(t)
    g.prefill_chunked(prompt)
    _close(g.logits(), ref)
    g.close()
    om.close()


def test_prompt_longer_than_the_sliding_window():
    """Mistral family: the prompt pass masks i - j >= window (one Forward of seqLen > 1 in the reference,
    grouped_query_attention.go:1074-1077), decode steps attend the whole cache.  Prompt of 300 tokens against a window of 32:
    chunked prefill (tcgen05 GEMMs, batched tolerance) and the token-by-token prefill (decode kernels, engine tolerance)
    against the CPU engine's prompt pass; then greedy decode continues identically."""
    from zerfoo_b200 import engine as E
    path = Z.path("mistral_q5_k_m")             # sliding_window = 32 in the miniature
    om = O.Model(path)
    rng = np.random.default_rng(11)
    n = 200
    prompt = [int(t) for t in rng.integers(1, om.vocab, size=n)]
    ref = om.prefill(prompt)
    # the mask matters: the unmasked pass gives different logits
    om2 = O.Model(path)
    for t in prompt:
        unmasked = om2.forward(t)
    assert np.abs(unmasked - ref).max() > 1e-2 * np.abs(ref).max()
    om2.close()
    g = E.load_file(path)
    first = g.prefill(prompt)                   # token by token through the decode kernels
    got = g.logits()
    assert np.abs(got - ref).max() <= 1e-3 * np.abs(ref).max()
    assert first == O.argmax(ref)
    tok, rtok = first, O.argmax(ref)
    for _ in range(24):                         # decode steps: whole cache on both sides
        tok = g.decode_step(tok)
        rtok = O.argmax(om.forward(rtok))
        assert tok == rtok
    g.reset()
    g.prefill_chunked(prompt)
    _close(g.logits(), ref)
    g.close()
    om.close()


def test_chunked_prefill_rejects_unsupported():
    from zerfoo_b200 import engine as E
    g = E.load_file(Z.path("gemma3_q4_0"))
    with pytest.raises(RuntimeError):
        g.prefill_chunked([1, 2, 3])
    g.close()


@pytest.mark.parametrize("f32", [True, False])
@pytest.mark.parametrize("hd,nq,nkv,p0,T,window", [(128, 8, 2, 0, 37, 0), (64, 4, 4, 19, 16, 0), (128, 6, 2, 5, 9, 0), (256, 4, 1, 3, 21, 0),
                                                    (32, 8, 1, 0, 33, 0), (128, 4, 2, 70, 130, 0), (64, 2, 1, 1, 64, 0),
                                                    # prompt longer than the sliding window (grouped_query_attention.go:1395-1415): row i sees i - j < window
                                                    (128, 8, 2, 0, 150, 32), (64, 4, 2, 40, 130, 48), (128, 4, 2, 0, 200, 70), (32, 8, 1, 10, 90, 1)])
def test_prefill_attention_kernel(hd, nq, nkv, p0, T, window, f32):
    from zerfoo_b200 import kernels as K
    rng = np.random.default_rng(hd + T)
    max_seq = 64 if p0 + T <= 64 else 256
    ld = (nq + 2 * nkv) * hd
    qkv = rng.standard_normal((T, ld)).astype(np.float32)
    wq = (1 + 0.1 * rng.standard_normal(hd)).astype(np.float32)
    wk = (1 + 0.1 * rng.standard_normal(hd)).astype(np.float32)
    half = hd // 2
    ang = np.arange(max_seq)[:, None] * (10000.0 ** (-np.arange(half) / half))[None, :]
    cos, sin = np.cos(ang).astype(np.float32), np.sin(ang).astype(np.float32)
    kc0 = rng.standard_normal((nkv, max_seq, hd)).astype(np.float32)
    vc0 = rng.standard_normal((nkv, max_seq, hd)).astype(np.float32)
    out, kc, vc = K.prefill_attn(qkv, wq, wk, cos, sin, p0, kc0, vc0, 1e-6, hd, nq, nkv, f32=f32, window=window)

    def norm_rope(x, w, pos):
        x = x.astype(np.float64)
        x = x / np.sqrt((x * x).mean() + 1e-6) * w
        a, b = x[:half], x[half:]
        return np.concatenate([a * cos[pos] - b * sin[pos], b * cos[pos] + a * sin[pos]])

    rk, rv = kc0.astype(np.float64), vc0.astype(np.float64)
    for i in range(T):
        for h in range(nkv):
            rk[h, p0 + i] = norm_rope(qkv[i, (nq + h) * hd:(nq + h + 1) * hd], wk, p0 + i)
            rv[h, p0 + i] = qkv[i, (nq + nkv + h) * hd:(nq + nkv + h + 1) * hd]
    np.testing.assert_allclose(kc, rk, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(vc, rv, rtol=0, atol=0)
    ref = np.zeros((T, nq * hd))
    for i in range(T):
        for h in range(nq):
            q = norm_rope(qkv[i, h * hd:(h + 1) * hd], wq, p0 + i)
            kvh = h // (nq // nkv)
            lo = max(0, p0 + i + 1 - window) if window > 0 else 0
            s = rk[kvh, lo:p0 + i + 1] @ q / np.sqrt(hd)
            p = np.exp(s - s.max())
            ref[i, h * hd:(h + 1) * hd] = (p / p.sum()) @ rv[kvh, lo:p0 + i + 1]
    # tensor-core path: fp16 operands (11-bit mantissa) on q, k, p, v
    tol = 2e-5 if f32 or hd == 256 else 3e-3
    np.testing.assert_allclose(out, ref, rtol=tol, atol=tol)
