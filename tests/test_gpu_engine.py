"""End-to-end parity of the CUDA decode engine (through the zb_engine_* C ABI)
against the restated reference CPU engine on identical synthetic-weight GGUFs:
identical greedy tokens, logits within tolerance, identical KV layout.

Tolerance on logits: the reference's GPU-vs-CPU layer bound is 1e-3
(tests/parity/gpu_parity_ops_test.go:457); we assert the tighter
1e-3 * max|logit| over the whole vocabulary after every step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import modelzoo as Z
from oracle import oracle as O
from zerfoo_b200 import gguf as G

torch = pytest.importorskip("torch")

DENSE = ["gemma3_q4_0", "llama_q4_k_m", "mistral_q5_k_m", "llama_q8_0", "mixtral_q4_k_m"]


@pytest.fixture(scope="module")
def E():
    from zerfoo_b200 import engine
    return engine


def margin_ok(logits_ref, tol):
    s = np.sort(logits_ref)
    return (s[-1] - s[-2]) > 2 * tol


@pytest.mark.parametrize("kind", DENSE)
@pytest.mark.parametrize("graph", [True, False], ids=["graph", "eager"])
def test_logits_kv_and_tokens_match_oracle(E, kind, graph):
    path = Z.path(kind)
    om = O.Model(path)
    g = E.load_file(path, use_graph=graph)
    assert (g.info.vocab, g.info.hidden, g.info.layers) == (om.vocab, om.hidden, om.layers)
    first = g.prefill(Z.PROMPT)
    for t in Z.PROMPT[:-1]:
        om.forward(t, want_logits=False)
    ref = om.forward(Z.PROMPT[-1])
    got = g.logits()
    tol = 1e-3 * np.abs(ref).max()
    assert np.abs(got - ref).max() <= tol
    if margin_ok(ref, np.abs(got - ref).max()):
        assert first == O.argmax(ref)
    n = len(Z.PROMPT)
    for layer in (0, g.info.layers - 1):
        k, v = g.kv(layer, n)
        rk, rv = om.kv(layer, n)
        assert np.abs(k - rk).max() <= 1e-3 * max(1.0, np.abs(rk).max())
        assert np.abs(v - rv).max() <= 1e-3 * max(1.0, np.abs(rv).max())
    # a few decode steps through the public per-token API (H2D token, D2H argmax)
    tok = O.argmax(ref)
    for _ in range(8):
        nxt = g.decode_step(tok)
        ref = om.forward(tok)
        got = g.logits()
        assert np.abs(got - ref).max() <= 1e-3 * np.abs(ref).max()
        if margin_ok(ref, np.abs(got - ref).max()):
            assert nxt == O.argmax(ref)
        tok = O.argmax(ref)
    assert g.position == om.pos
    g.close(); om.close()


@pytest.mark.parametrize("kind", DENSE)
def test_greedy_128_tokens_identical(E, kind):
    """North star: greedy token sequences identical for 128 tokens."""
    path = Z.path(kind)
    om = O.Model(path)
    ref = om.generate(Z.PROMPT, 128)
    g = E.load_file(path)
    got = g.generate(Z.PROMPT, 128)
    assert got == ref
    # the device-chained path (no host round trip per token) produces the same stream
    g.reset()
    first = g.prefill(Z.PROMPT)
    rest, ms = g.decode_n(first, 127)
    assert [first] + rest == ref and ms > 0
    g.close(); om.close()


@pytest.mark.parametrize("kind", ["llama_q4_k_m", "gemma3_q4_0"])
def test_attention_tile_switch_keeps_the_stream(E, kind, monkeypatch):
    """Short contexts run one long attention tile per KV head (16 warps, no split merge); once kv_len exceeds the tile the
    engine launches the split-KV graph.  With the tile forced to 64 positions a 128-token stream crosses the switch: tokens
    and logits must match the oracle on both sides, through the per-token API and through the chained decode."""
    path = Z.path(kind)
    om = O.Model(path)
    ref = om.generate(Z.PROMPT, 128)
    monkeypatch.setenv("ZB_ATTN_SHORT", "64")
    g = E.load_file(path)
    assert g.generate(Z.PROMPT, 128) == ref
    lr = om.forward(ref[-1])
    g.decode_step(ref[-1])
    assert np.abs(g.logits() - lr).max() <= 1e-3 * np.abs(lr).max()
    g.reset()
    first = g.prefill(Z.PROMPT)
    rest, _ = g.decode_n(first, 127)
    assert [first] + rest == ref
    g.close()
    monkeypatch.setenv("ZB_ATTN_SHORT", "0")   # split-KV graph only
    g = E.load_file(path)
    assert g.generate(Z.PROMPT, 128) == ref
    g.close(); om.close()


def test_generate_twice_is_idempotent_and_reset_works(E):
    g = E.load_file(Z.path("llama_q4_k_m"))
    a = g.generate(Z.PROMPT, 32)
    b = g.generate(Z.PROMPT, 32)
    assert a == b
    g.reset()
    assert g.position == 0
    g.close()


def test_engine_errors_are_loud(E, tmp_path):
    with pytest.raises(E.EngineError, match="open"):
        E.load_file(str(tmp_path / "missing.gguf"))
    bad = tmp_path / "bad.gguf"
    bad.write_bytes(b"GGUF" + b"\0" * 64)
    with pytest.raises(E.EngineError):
        E.load_file(str(bad))
    g = E.load_file(Z.path("llama_q8_0"), max_seq=24)
    with pytest.raises(E.EngineError, match="out of range"):
        g.decode_step(g.info.vocab)                # arch_llama.go:292-295
    with pytest.raises(E.EngineError, match="out of range"):
        g.prefill([1, -5])
    g.reset()
    g.prefill(list(range(1, 21)))
    with pytest.raises(E.EngineError, match="KV"):
        g.decode_n(3, 10)                          # 20 + 10 > 24
    g.close()


def test_c1_shape_reduced_layers(E):
    """BASELINE config 1 (Gemma-3-1B shape, Q4_0, tied 262144-row head) with 2 layers:
    full-width kernels (K=1152/6912, V=262144) against the oracle."""
    kind = "preset:c1:2:256"
    path = Z.path(kind)
    om = O.Model(path)
    ref = om.generate(Z.PROMPT, 16)
    g = E.load_file(path)
    assert g.generate(Z.PROMPT, 16) == ref
    lr = om.forward(ref[-1])
    g.decode_step(ref[-1])
    assert np.abs(g.logits() - lr).max() <= 1e-3 * np.abs(lr).max()
    assert abs(g.info.weight_bytes_per_token - G.model_weight_bytes(Z.spec(kind))) / g.info.weight_bytes_per_token < 0.02
    g.close(); om.close()


def test_c2_shape_reduced_layers(E):
    """BASELINE config 2 (Llama-3.2-3B shape, Q4_K_M: Q4_K + Q6_K) with 2 layers."""
    kind = "preset:c2:2:256"
    path = Z.path(kind)
    om = O.Model(path)
    ref = om.generate(Z.PROMPT, 12)
    g = E.load_file(path)
    assert g.generate(Z.PROMPT, 12) == ref
    g.close(); om.close()


def test_wide_ffn_down_projection_runs_as_column_slabs(E):
    """ffn > 16384 (the 70B shape has 28672): the down projection is cut into column slabs that run side by side on the
    tensor-core GEMV, their partial outputs summed by the next prologue.  128 greedy tokens identical to the CPU engine,
    logits within the engine bar."""
    path = Z.path("llama_wide_ffn")
    om = O.Model(path)
    ref = om.generate(Z.PROMPT, 128)
    g = E.load_file(path)
    got = g.generate(Z.PROMPT, 128)
    assert got == ref
    lr = om.forward(ref[-1])
    g.decode_step(ref[-1])
    assert np.abs(g.logits() - lr).max() <= 1e-3 * np.abs(lr).max()
    g.close(); om.close()


@pytest.mark.parametrize("kind", ["llama_q4_k_m", "gemma3_q4_0"])
def test_fp16_kv_cache_matches_fp16_rounded_oracle(E, kind):
    """KV stored as fp16 (generate/tensor_cache.go:224-238 kvFP16): the append rounds K / V, attention reads half the bytes,
    arithmetic stays f32.  Against the CPU engine with the same rounding on its cache writes: identical greedy tokens over
    several attention tiles, logits within the engine bar, cache rows equal up to one fp16 ulp, and half the KV bytes."""
    path = Z.path(kind)
    om = O.Model(path)
    om.set_kv_f16(True)
    ref = om.generate(Z.PROMPT, 100)
    g = E.load_file(path, kv_f16=True)
    g32 = E.load_file(path)
    assert g.refresh_info().kv_bytes_per_pos * 2 == g32.refresh_info().kv_bytes_per_pos
    g32.close()
    got = g.generate(Z.PROMPT, 100)
    assert got == ref
    lr = om.forward(ref[-1])
    g.decode_step(ref[-1])
    assert np.abs(g.logits() - lr).max() <= 1e-3 * np.abs(lr).max()
    n = len(Z.PROMPT) + 100
    for layer in (0, g.info.layers - 1):
        k, v = g.kv(layer, n)
        rk, rv = om.kv(layer, n)
        assert np.array_equal(k, k.astype(np.float16).astype(np.float32))          # what the tap returns IS fp16 data
        assert np.abs(k - rk).max() <= 2e-3 * max(1.0, np.abs(rk).max())
        assert np.abs(v - rv).max() <= 2e-3 * max(1.0, np.abs(rv).max())
    with pytest.raises(E.EngineError, match="fp16"):
        g.prefill_chunked([1, 2, 3])
    g.close(); om.close()
