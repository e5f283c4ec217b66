"""The C oracle's model wiring (layer order, GQA, RoPE position/base, norms,
MoE routing, KV cache) against an independent numpy/float64 restatement on tiny
synthetic GGUFs, plus GGUF writer/reader round trips."""
import numpy as np
import pytest

import modelzoo as Z
import np_model
from oracle import oracle as O
from zerfoo_b200 import gguf as G


@pytest.mark.parametrize("kind", Z.KINDS)
def test_oracle_matches_numpy_forward(kind):
    path = Z.path(kind)
    om, nm = O.Model(path), np_model.NpModel(path)
    toks = Z.PROMPT[:6]
    for t in toks:
        lo = om.forward(t)
        ln = nm.forward(t)
        scale = np.abs(ln).max()
        assert np.abs(lo - ln).max() <= 2e-4 * scale + 1e-5, kind
        assert int(np.argmax(ln)) == O.argmax(lo)
    k, v = om.kv(0, len(toks))
    assert np.abs(k - np.stack(nm.k[0]).reshape(len(toks), -1)).max() < 1e-4
    assert np.abs(v - np.stack(nm.v[0]).reshape(len(toks), -1)).max() < 1e-4
    om.close()


@pytest.mark.parametrize("kind", ["gemma3_q4_0", "llama_q4_k_m"])
def test_generate_is_deterministic_and_matches_stepwise(kind):
    path = Z.path(kind)
    m = O.Model(path)
    a = m.generate(Z.PROMPT, 24)
    b = m.generate(Z.PROMPT, 24)
    assert a == b and len(a) == 24
    m.reset()
    for t in Z.PROMPT[:-1]:
        m.forward(t, want_logits=False)
    tok = O.argmax(m.forward(Z.PROMPT[-1]))
    step = [tok]
    for _ in range(23):
        tok = O.argmax(m.forward(tok))
        step.append(tok)
    assert step == a
    assert m.pos == len(Z.PROMPT) + 23
    m.close()


def test_token_out_of_range_is_an_error():
    m = O.Model(Z.path("llama_q8_0"))
    with pytest.raises(RuntimeError):
        m.forward(m.vocab)          # "token ID %d out of range" (arch_llama.go:292-295)
    m.close()


def test_kv_capacity_is_enforced():
    m = O.Model(Z.path("llama_q8_0"), max_seq=4)
    for t in range(4):
        m.forward(t)
    with pytest.raises(RuntimeError):
        m.forward(1)
    m.close()


def test_gguf_roundtrip_and_alignment(tmp_path):
    s = Z.spec("llama_q4_k_m")
    p = str(tmp_path / "m.gguf")
    G.write_synthetic_gguf(p, s, seed=7)
    f = G.read_gguf(p)
    assert f.metadata["general.architecture"] == "llama"
    assert f.metadata["llama.block_count"] == s.layers
    plan = G.tensor_plan(s)
    assert set(f.tensors) == {n for n, *_ in plan}
    for name, qt, ne, _ in plan:
        t = f.tensors[name]
        assert t.qtype == qt and t.ne == tuple(ne)
        assert (t.data.ctypes.data - f.tensors[plan[0][0]].data.ctypes.data) % 32 == 0   # 32-byte data alignment
    # Q4_K_M mix: Q6_K on the "more bits" layers' attn_v / ffn_down
    assert f.tensors["blk.0.attn_v.weight"].qtype == G.Q6_K
    assert f.tensors["blk.1.attn_q.weight"].qtype == G.Q4_K


def test_truncated_gguf_rejected(tmp_path):
    p = Z.path("llama_q8_0")
    data = open(p, "rb").read()
    bad = tmp_path / "bad.gguf"
    bad.write_bytes(data[: len(data) // 2])
    with pytest.raises(RuntimeError):
        O.Model(str(bad))
    bad.write_bytes(b"NOPE" + data[4:])
    with pytest.raises(RuntimeError):
        O.Model(str(bad))


def test_weight_byte_accounting():
    s = G.preset("c1")
    b = G.model_weight_bytes(s)
    assert abs(b - 0.562e9) / 0.562e9 < 0.03     # C1 0.562 GB/token


def test_prompt_pass_sliding_window_against_numpy():
    """Mistral family: the prompt pass sees (i - window, i] only (grouped_query_attention.go:1074-1077,1395-1415), decode
    steps see the whole cache.  The C oracle's zo_model_prefill against the independent numpy restatement at a prompt longer
    than the window (the miniature has sliding_window = 32), then two decode steps on both."""
    path = Z.path("mistral_q5_k_m")
    om, nm = O.Model(path), np_model.NpModel(path)
    assert nm.window == 32
    rng = np.random.default_rng(3)
    prompt = [int(t) for t in rng.integers(1, om.vocab, size=48)]
    lo, ln = om.prefill(prompt), nm.prefill(prompt)
    scale = np.abs(ln).max()
    assert np.abs(lo - ln).max() <= 2e-4 * scale + 1e-5
    # the mask is not a no-op at this length ...
    nu = np_model.NpModel(path)
    for t in prompt:
        lu = nu.forward(t)
    assert np.abs(lu - ln).max() > 1e-3 * scale
    # ... and decode steps after the prompt attend everything again
    tok = int(np.argmax(ln))
    for _ in range(2):
        lo, ln = om.forward(tok), nm.forward(tok)
        assert np.abs(lo - ln).max() <= 2e-4 * np.abs(ln).max() + 1e-5
        tok = int(np.argmax(ln))
    # a prompt inside the window is unaffected
    om.reset()
    short = prompt[:20]
    a = om.prefill(short)
    om.reset()
    for t in short:
        b = om.forward(t)
    assert np.array_equal(a, b)
    om.close()


def test_gemma_builders_do_not_apply_a_prompt_window():
    """arch_gemma.go does not hand slidingWindowSize to the attention layers: Gemma's metadata window is ignored by the prompt pass."""
    path = Z.path("gemma3_q4_0")               # sliding_window = 64 in the metadata
    om = O.Model(path, max_seq=128)
    rng = np.random.default_rng(5)
    prompt = [int(t) for t in rng.integers(1, om.vocab, size=80)]
    a = om.prefill(prompt)
    om.reset()
    for t in prompt:
        b = om.forward(t)
    assert np.array_equal(a, b)
    om.close()


@pytest.mark.parametrize("kind", ["llama_q4_k_m", "mistral_q5_k_m"])
def test_fp16_kv_cache_against_numpy(kind):
    """kvFP16 (generate/tensor_cache.go:224-238): K and V are rounded to fp16 when they enter the cache, attention reads the
    rounded values.  Oracle vs numpy restatement, and the cache rows are exactly fp16-representable."""
    path = Z.path(kind)
    om, nm = O.Model(path), np_model.NpModel(path)
    om.set_kv_f16(True)
    nm.kv_f16 = True
    for t in Z.PROMPT[:8]:
        lo, ln = om.forward(t), nm.forward(t)
        assert np.abs(lo - ln).max() <= 3e-4 * np.abs(ln).max() + 1e-5
    k, v = om.kv(0, 8)
    assert np.array_equal(k, k.astype(np.float16).astype(np.float32)) and np.array_equal(v, v.astype(np.float16).astype(np.float32))
    om.close()
