"""Pin the CPU oracle against the reference's own golden vectors
(tests/golden/ref_layers/*.json, imported from the reference's
tests/golden/layers by tests/golden/import_reference_goldens.py) with the
tolerance each fixture carries -- the same check tests/parity/layer_parity_test.go
runs against the Go CPU engine."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O


def load(golden_dir, name):
    with open(os.path.join(golden_dir, "ref_layers", name + ".json")) as f:
        return json.load(f)


def arr(d, key):
    return np.array(d[key], dtype=np.float32).reshape(d[key.replace("expected_", "") + "_shape"] if False else -1)


def test_rms_norm(golden_dir):
    d = load(golden_dir, "norm_rms_norm")
    x = np.array(d["input"], np.float32).reshape(d["input_shape"])
    g = np.array(d["gain"], np.float32)
    want = np.array(d["expected_output"], np.float32).reshape(d["output_shape"])
    got = O.rmsnorm(x, g, d["epsilon"])
    assert np.abs(got - want).max() <= d["tolerance"]


def test_rotary(golden_dir):
    d = load(golden_dir, "embedding_rotary")
    x = np.array(d["input"], np.float32).reshape(d["input_shape"])
    want = np.array(d["expected_output"], np.float32).reshape(d["output_shape"])
    seq, hd = x.shape[1], x.shape[2]
    cs, sn = O.rope_tables(seq, hd, d["base"])
    assert np.abs(cs - np.array(d["cos_values"], np.float32).reshape(d["cos_shape"])).max() <= d["tolerance"]
    assert np.abs(sn - np.array(d["sin_values"], np.float32).reshape(d["sin_shape"])).max() <= d["tolerance"]
    got = np.stack([O.rope(x[0, p], cs[p], sn[p]) for p in range(seq)])[None]
    assert np.abs(got - want).max() <= d["tolerance"]


def test_rope_partial_passthrough():
    x = np.arange(12, dtype=np.float32)
    cs, sn = O.rope_tables(4, 8, 10000.0)
    y = O.rope(x, cs[3], sn[3])
    assert np.array_equal(y[8:], x[8:])


def test_swiglu(golden_dir):
    d = load(golden_dir, "activation_swiglu")
    x = np.array(d["input"], np.float32).reshape(d["input_shape"])
    want = np.array(d["expected_output"], np.float32).reshape(d["output_shape"])
    half = x.shape[1] // 2
    got = O.swiglu(x[:, :half].copy(), x[:, half:].copy())
    assert np.abs(got - want).max() <= d["tolerance"]


def test_silu(golden_dir):
    d = load(golden_dir, "activation_silu")
    x = np.array(d["input"], np.float32)
    want = np.array(d["expected_output"], np.float32)
    assert np.abs(O.silu(x) - want).max() <= max(d["tolerance"], 1e-6)


def test_softmax(golden_dir):
    d = load(golden_dir, "activation_softmax")
    x = np.array(d["input"], np.float32).reshape(d["input_shape"])
    want = np.array(d["expected_output"], np.float32).reshape(d["output_shape"])
    assert np.abs(O.softmax(x) - want).max() <= max(d["tolerance"], 1e-6)


def test_sdpa_causal(golden_dir):
    d = load(golden_dir, "attention_sdpa_causal")
    q = np.array(d["query"], np.float32).reshape(d["query_shape"])
    k = np.array(d["key"], np.float32).reshape(d["key_shape"])
    v = np.array(d["value"], np.float32).reshape(d["value_shape"])
    want = np.array(d["expected_output"], np.float32).reshape(d["output_shape"])
    got = O.attn_causal(q, k, v)
    assert np.abs(got - want).max() <= d["tolerance"]
    # the decode path (one query over a cache) must agree with the causal row
    seq, hd = q.shape[1], q.shape[2]
    for i in range(seq):
        row = O.attn_decode(q[0, i:i + 1], k[0], v[0], n_kv=1, kv_len=i + 1)
        assert np.abs(row[0] - want[0, i]).max() <= d["tolerance"]


def test_gqa(golden_dir):
    d = load(golden_dir, "attention_gqa")
    nq, nkv, hd = d["n_q_heads"], d["n_kv_heads"], d["head_dim"]
    x = np.array(d["input"], np.float32).reshape(d["input_shape"])[0]
    W = {k: np.array(d[k], np.float32).reshape(d[k + "_shape"]) for k in
         ("wq_w", "wq_b", "wk_w", "wk_b", "wv_w", "wv_b", "wo_w", "wo_b")}
    want = np.array(d["expected_output"], np.float32).reshape(d["output_shape"])[0]
    q = x @ W["wq_w"] + W["wq_b"]
    k = x @ W["wk_w"] + W["wk_b"]
    v = x @ W["wv_w"] + W["wv_b"]
    seq = x.shape[0]
    out = np.empty((seq, nq * hd), np.float32)
    for i in range(seq):  # decode-style: query i over cache[0..i], KV heads shared by q_head // (nq/nkv)
        o = O.attn_decode(q[i].reshape(nq, hd), k, v, n_kv=nkv, kv_len=i + 1)
        out[i] = o.reshape(-1)
    got = out @ W["wo_w"] + W["wo_b"]
    assert np.abs(got - want).max() <= d["tolerance"]


def test_ffn(golden_dir):
    d = load(golden_dir, "core_ffn")
    x = np.array(d["input"], np.float32).reshape(d["input_shape"])
    w1 = np.array(d["w1"], np.float32).reshape(d["w1_shape"])
    w2 = np.array(d["w2"], np.float32).reshape(d["w2_shape"])
    w3 = np.array(d["w3"], np.float32).reshape(d["w3_shape"])
    want = np.array(d["expected_output"], np.float32).reshape(d["output_shape"])
    from zerfoo_b200 import gguf as G
    # weights stored [out, in] row-major F32 "blocks", the layout the GEMV takes
    gate = np.stack([O.gemv(G.F32, np.ascontiguousarray(w1.T), w1.shape[1], w1.shape[0], r) for r in x])
    up = np.stack([O.gemv(G.F32, np.ascontiguousarray(w3.T), w3.shape[1], w3.shape[0], r) for r in x])
    act = O.swiglu(gate, up)
    got = np.stack([O.gemv(G.F32, np.ascontiguousarray(w2.T), w2.shape[1], w2.shape[0], r) for r in act])
    assert np.abs(got - want).max() <= d["tolerance"]


def test_moe_routing(golden_dir):
    d = load(golden_dir, "core_moe")
    x = np.array(d["input"], np.float32).reshape(d["input_shape"])
    gw = np.array(d["gate_weight"], np.float32).reshape(d["gate_weight_shape"])
    experts = [np.array(w, np.float32).reshape(d["expert_weight_shape"]) for w in d["expert_weights"]]
    want = np.array(d["expected_output"], np.float32).reshape(d["output_shape"])
    got = np.zeros_like(want)
    for t in range(x.shape[0]):
        idx, w = O.moe_route(gw @ x[t], d["top_k"])
        assert abs(float(w.sum()) - 1.0) < 1e-6
        for e, wk in zip(idx, w):
            got[t] += wk * (x[t] @ experts[e])
    assert np.abs(got - want).max() <= d["tolerance"]


def test_lm_head(golden_dir):
    d = load(golden_dir, "core_lm_head")
    x = np.array(d["input"], np.float32).reshape(d["input_shape"])[0]
    w = np.array(d["weight"], np.float32).reshape(d["weight_shape"])  # [hidden, vocab]
    want = np.array(d["expected_output"], np.float32).reshape(d["output_shape"])[0]
    from zerfoo_b200 import gguf as G
    got = O.gemm_nt(G.F32, np.ascontiguousarray(w.T), w.shape[1], w.shape[0], x)
    assert np.abs(got - want).max() <= d["tolerance"]


def test_argmax_lowest_index_wins():
    x = np.array([1.0, 3.0, 3.0, 2.0, 3.0], np.float32)
    assert O.argmax(x) == 1


def test_softcap_rational_tanh():
    # inference/arch_llama.go:15-27: x(27+x^2)/(27+9x^2), +-1 beyond |x|>=4.5
    x = np.array([0.0, 15.0, -15.0, 300.0, -300.0], np.float32)
    y = O.softcap(x, 30.0)
    t = np.float32(0.5)
    want = np.float32(30.0) * (t * (27 + t * t) / (27 + 9 * t * t))
    assert y[0] == 0 and abs(y[1] - want) < 1e-5 and abs(y[2] + want) < 1e-5
    assert y[3] == 30.0 and y[4] == -30.0


def test_fp16_decode_quirks():
    # internal/xblas/q4dot.go:53-80: subnormals scale, Inf/NaN decode to 0
    assert O.fp16_to_f32(0x3C00) == 1.0
    assert O.fp16_to_f32(0xC000) == -2.0
    assert O.fp16_to_f32(0x0001) == np.float32(2.0 ** -24)
    assert O.fp16_to_f32(0x7C00) == 0.0 and O.fp16_to_f32(0x7E00) == 0.0
    for bits in (0x0400, 0x3555, 0x7BFF, 0x83FF, 0x0200):
        assert O.fp16_to_f32(bits) == np.float32(np.array([bits], np.uint16).view(np.float16)[0])


# ---- block formats pinned on gguf-py (tests/golden/make_gguf_py_goldens.py) -------------------------------------------
GGUF_PY_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gguf_py")
GGUF_PY_TYPES = ["Q4_0", "Q8_0", "Q4_K", "Q5_K", "Q6_K"]


@pytest.mark.parametrize("name", GGUF_PY_TYPES)
def test_dequant_matches_gguf_py_golden(name):
    """The oracle's (and the independent numpy restatement's) dequantisation of committed raw blocks equals gguf-py's
    values bit for bit: Q5_K / Q6_K / Q8_0 have no golden bytes upstream, gguf-py is the canonical third party."""
    from zerfoo_b200 import gguf as G
    import refdata as R
    z = np.load(os.path.join(GGUF_PY_DIR, name + ".npz"))
    qt = {v: k for k, v in G.TYPE_NAMES.items()}[name]
    raw, want = z["raw"], z["values"]
    got = O.dequant(qt, raw, want.size)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.abs(got - want).max())
    assert np.array_equal(R.np_dequant(qt, raw).view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("name", GGUF_PY_TYPES)
def test_dequant_matches_gguf_py_live(name):
    """Same check against the installed gguf-py on fresh random blocks (skipped where the package is absent)."""
    gguf = pytest.importorskip("gguf")
    from gguf import quants
    from zerfoo_b200 import gguf as G
    qt = {v: k for k, v in G.TYPE_NAMES.items()}[name]
    rng = np.random.default_rng(99 + qt)
    w = rng.standard_normal((4, 2048), dtype=np.float32)
    raw = G.quantize(w, qt)
    gg = raw
    if name == "Q5_K":   # the reference keeps ql before qh (gemv_q5k.cu:8-12); ggml's block_q5_K has qh first
        b = raw.reshape(-1, 176)
        gg = np.ascontiguousarray(np.concatenate([b[:, :16], b[:, 144:176], b[:, 16:144]], axis=1)).reshape(raw.shape)
    want = quants.dequantize(gg, getattr(gguf.GGMLQuantizationType, name)).astype(np.float32).reshape(-1)
    got = O.dequant(qt, raw, w.size)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
