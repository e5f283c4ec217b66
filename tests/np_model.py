"""Second, independent restatement of the decoder forward in numpy/float64
(dequantise everything with tests/refdata.np_dequant, wire the layers by hand
from inference/arch_common.go:150-526).  Only used to cross-check the wiring of
the C oracle on tiny models -- it is slow and allocates the full f64 weights."""
import numpy as np

import refdata as R
from zerfoo_b200 import gguf as G


class NpModel:
    def __init__(self, path):
        f = G.read_gguf(path)
        md = f.metadata
        a = md["general.architecture"]
        self.arch = a
        g = lambda k, d=None: md.get(f"{a}.{k}", d)
        self.hidden, self.layers = g("embedding_length"), g("block_count")
        self.nq, self.nkv = g("attention.head_count"), g("attention.head_count_kv")
        self.hd = g("attention.key_length") or self.hidden // self.nq
        self.eps = g("attention.layer_norm_rms_epsilon", 1e-5)
        self.base = g("rope.freq_base", 10000.0)
        self.local = g("rope.local.freq_base", 0.0)
        self.softcap = g("final_logit_softcapping", 0.0) if a == "gemma3" else 0.0
        self.n_exp, self.top_k = g("expert_count", 0), g("expert_used_count", 0)
        self.gemma3 = a == "gemma3"
        self.scale = np.sqrt(float(self.hidden)) if a.startswith("gemma") else 0.0
        self.w = {}
        for name, t in f.tensors.items():
            self.w[name] = R.np_dequant(t.qtype, np.asarray(t.data)).astype(np.float64).reshape(t.rows, t.cols) if t.qtype != G.F32 else \
                np.asarray(t.data).view(np.float32).astype(np.float64).reshape(t.rows, t.cols)
        self.k = [[] for _ in range(self.layers)]
        self.v = [[] for _ in range(self.layers)]
        self.pos = 0
        # prompt-pass sliding window (grouped_query_attention.go:1074-1077,1395-1415): Mistral-family builders only; 0 = off.
        self.window = g("attention.sliding_window", 0) if a in ("mistral", "mixtral", "starcoder2") else 0
        self.in_prefill = False
        self.kv_f16 = False     # generate/tensor_cache.go:224-238: K / V rounded to fp16 on the cache write

    def rms(self, x, w):
        return x / np.sqrt(np.mean(x * x) + self.eps) * w.reshape(-1)

    def rope(self, x, base):
        hd = self.hd
        half = hd // 2
        inv = 1.0 / base ** (2.0 * np.arange(half) / hd)
        ang = self.pos * inv
        c, s = np.float32(np.cos(ang)).astype(np.float64), np.float32(np.sin(ang)).astype(np.float64)
        a, b = x[:half], x[half:]
        return np.concatenate([a * c - b * s, b * c + a * s])

    def ffn(self, x, gate, up, down):
        g, u = gate @ x, up @ x
        return down @ (g / (1 + np.exp(-g)) * u)

    def forward(self, tok):
        W = self.w
        h = W["token_embd.weight"][tok].copy()
        if self.scale:
            h = h * np.float64(np.float32(self.scale))
        for i in range(self.layers):
            p = f"blk.{i}."
            n = self.rms(h, W[p + "attn_norm.weight"])
            q, k, v = W[p + "attn_q.weight"] @ n, W[p + "attn_k.weight"] @ n, W[p + "attn_v.weight"] @ n
            base = self.base
            if self.local and (i + 1) % 6 != 0:
                base = self.local
            q = q.reshape(self.nq, self.hd)
            k = k.reshape(self.nkv, self.hd)
            if self.gemma3:
                q = np.stack([self.rms(r, W[p + "attn_q_norm.weight"]) for r in q])
                k = np.stack([self.rms(r, W[p + "attn_k_norm.weight"]) for r in k])
            q = np.stack([self.rope(r, base) for r in q])
            k = np.stack([self.rope(r, base) for r in k])
            v = v.reshape(self.nkv, self.hd)
            if self.kv_f16:
                k = k.astype(np.float32).astype(np.float16).astype(np.float64)
                v = v.astype(np.float32).astype(np.float16).astype(np.float64)
            self.k[i].append(k)
            self.v[i].append(v)
            lo = max(0, self.pos + 1 - self.window) if (self.in_prefill and self.window) else 0
            K, V = np.stack(self.k[i][lo:]), np.stack(self.v[i][lo:])      # [T, nkv, hd]
            rep = self.nq // self.nkv
            out = []
            for hh in range(self.nq):
                s = K[:, hh // rep] @ q[hh] / np.sqrt(self.hd)
                e = np.exp(s - s.max())
                out.append((e / e.sum()) @ V[:, hh // rep])
            a = W[p + "attn_output.weight"] @ np.concatenate(out)
            if self.gemma3:
                a = self.rms(a, W[p + "post_attention_norm.weight"])
            res = a + h
            n2 = self.rms(res, W[p + "ffn_norm.weight"])
            if self.n_exp:
                logits = W[p + "ffn_gate_inp.weight"] @ n2
                pr = np.exp(logits - logits.max())
                pr /= pr.sum()
                top = np.argsort(-pr, kind="stable")[: self.top_k]
                wk = pr[top] / pr[top].sum()
                E = self.n_exp
                ge, ue, de = W[p + "ffn_gate_exps.weight"], W[p + "ffn_up_exps.weight"], W[p + "ffn_down_exps.weight"]
                fr, dr = ge.shape[0] // E, de.shape[0] // E
                f = sum(wk[j] * self.ffn(n2, ge[e * fr:(e + 1) * fr], ue[e * fr:(e + 1) * fr], de[e * dr:(e + 1) * dr]) for j, e in enumerate(top))
            else:
                f = self.ffn(n2, W[p + "ffn_gate.weight"], W[p + "ffn_up.weight"], W[p + "ffn_down.weight"])
            h = self.rms(f, W[p + "post_ffw_norm.weight"]) + res if self.gemma3 else f + res
        self.pos += 1
        self.last_hidden = h
        n = self.rms(h, W["output_norm.weight"])
        head = W.get("output.weight", W["token_embd.weight"])
        lg = head @ n
        if self.softcap:
            x = lg / self.softcap
            t = np.where(np.abs(x) > 4.5, np.sign(x), x * (27 + x * x) / (27 + 9 * x * x))
            lg = self.softcap * t
        return lg

    def prefill(self, prompt):
        """One Forward of seqLen = len(prompt) in the reference: the only place the sliding-window mask is applied."""
        self.in_prefill = True
        try:
            for t in prompt:
                lg = self.forward(t)
        finally:
            self.in_prefill = False
        return lg
