"""Tiny synthetic GGUF models in the shapes of the BASELINE configs, written by
zerfoo_b200.gguf in the pattern of the reference's writeTestGGUF
(inference/load_gguf_test.go:64-140).  Cached under a temp dir per session."""
import os
import tempfile

from zerfoo_b200 import gguf as G

_DIR = os.environ.get("ZB_TEST_MODEL_DIR") or os.path.join(tempfile.gettempdir(), "zb200_models")


def _mini(kind: str) -> G.ModelSpec:
    if kind == "gemma3_q4_0":      # C1 in miniature: qk-norm, post-norms, softcap, local/global rope, tied Q4_0 head
        return G.ModelSpec("gemma3", 640, 256, 7, 4, 1, 64, 512, ctx=256, rope_base=1e6, rope_local_base=1e4, eps=1e-6, softcap=30.0,
                           sliding_window=64, tied=True, base_type=G.Q4_0, name="mini-gemma3-q4_0")
    if kind == "llama_q4_k_m":     # C2 in miniature: Q4_K + Q6_K mix, tied Q6_K embedding/head, GQA 8/2
        return G.ModelSpec("llama", 512, 256, 8, 8, 2, 32, 512, ctx=256, rope_base=5e5, eps=1e-5, tied=True, base_type=G.Q4_K,
                           more_bits_type=G.Q6_K, embed_type=G.Q6_K, name="mini-llama-q4_k_m")
    if kind == "mistral_q5_k_m":   # C3 in miniature: Q5_K + Q6_K, untied head, sliding window (prefill only)
        return G.ModelSpec("mistral", 384, 256, 4, 8, 4, 32, 768, ctx=256, rope_base=1e6, eps=1e-5, tied=False, sliding_window=32,
                           base_type=G.Q5_K, more_bits_type=G.Q6_K, embed_type=G.Q5_K, name="mini-mistral-q5_k_m")
    if kind == "llama_q8_0":       # Q8_0 everywhere (embeddings / lm_head path of the reference)
        return G.ModelSpec("llama", 320, 128, 2, 4, 4, 32, 256, ctx=256, rope_base=1e4, eps=1e-5, tied=True, base_type=G.Q8_0,
                           name="mini-llama-q8_0")
    if kind == "mixtral_q4_k_m":   # C5 in miniature: 4 experts top-2
        return G.ModelSpec("mixtral", 384, 256, 3, 8, 2, 32, 512, ctx=256, rope_base=1e6, eps=1e-5, tied=False, n_experts=4, top_k=2,
                           base_type=G.Q4_K, more_bits_type=G.Q6_K, embed_type=G.Q4_K, name="mini-mixtral-q4_k_m")
    if kind == "llama_wide_ffn":   # ffn 16640 = 65 super-blocks > 16384: the down projection runs as 5 column slabs (engine.cu upload_down)
        return G.ModelSpec("llama", 384, 256, 3, 8, 2, 32, 16640, ctx=256, rope_base=5e5, eps=1e-5, tied=True, base_type=G.Q4_K,
                           more_bits_type=G.Q6_K, embed_type=G.Q6_K, name="mini-llama-wide-ffn")
    if kind == "llama_tp_q4_k_m":  # C4 in miniature: every K-quant split lands on a 256 boundary at TP = 2 (qd/2 = 256, ffn/2 = 512)
        return G.ModelSpec("llama", 512, 512, 4, 8, 2, 64, 1024, ctx=256, rope_base=5e5, eps=1e-5, tied=False, base_type=G.Q4_K,
                           more_bits_type=G.Q6_K, embed_type=G.Q4_K, name="mini-llama-tp-q4_k_m")
    if kind == "llama_tp8_q4_k_m":  # shards 8 ways: 16/8 heads x 128, qd/8 = 256, ffn/8 = 256
        return G.ModelSpec("llama", 512, 1024, 2, 16, 8, 128, 2048, ctx=256, rope_base=5e5, eps=1e-5, tied=False, base_type=G.Q4_K,
                           more_bits_type=G.Q6_K, embed_type=G.Q4_K, name="mini-llama-tp8-q4_k_m")
    if kind == "mixtral_tp_q4_k_m":  # C5 in miniature: 4 experts top-2, experts sharded 2 per rank at TP = 2
        return G.ModelSpec("mixtral", 384, 512, 3, 8, 2, 64, 512, ctx=256, rope_base=1e6, eps=1e-5, tied=False, n_experts=4, top_k=2,
                           base_type=G.Q4_K, more_bits_type=G.Q6_K, embed_type=G.Q4_K, name="mini-mixtral-tp-q4_k_m")
    if kind == "llama_tp_q8_0":    # 32-weight blocks shard at any multiple of 32
        return G.ModelSpec("llama", 320, 256, 3, 8, 2, 32, 512, ctx=256, rope_base=1e4, eps=1e-5, tied=True, base_type=G.Q8_0,
                           name="mini-llama-tp-q8_0")
    raise KeyError(kind)


KINDS = ["gemma3_q4_0", "llama_q4_k_m", "mistral_q5_k_m", "llama_q8_0", "mixtral_q4_k_m"]


def spec(kind: str) -> G.ModelSpec:
    if kind.startswith("preset:"):     # e.g. preset:c1:4  -> C1 with 4 layers
        parts = kind.split(":")
        return G.preset(parts[1], layers=int(parts[2]) if len(parts) > 2 else None, ctx=int(parts[3]) if len(parts) > 3 else None)
    return _mini(kind)


def path(kind: str, seed: int = 1234) -> str:
    os.makedirs(_DIR, exist_ok=True)
    p = os.path.join(_DIR, f"{kind.replace(':', '_')}_s{seed}.gguf")
    if not os.path.exists(p):
        tmp = p + f".tmp{os.getpid()}"
        G.write_synthetic_gguf(tmp, spec(kind), seed=seed)
        os.replace(tmp, p)
    return p


PROMPT = [2] + list(range(100, 116))   # fixed token ids
