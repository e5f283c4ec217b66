"""ctypes face of the CPU oracle (oracle/zoracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never import this from
zerfoo_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libzoracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "zoracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        f32p, i32p, u8p = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
        L.zo_fp16_to_f32.restype = C.c_float
        L.zo_fp16_to_f32.argtypes = [C.c_uint16]
        L.zo_f32_to_fp16.restype = C.c_uint16
        L.zo_f32_to_fp16.argtypes = [C.c_float]
        L.zo_row_bytes.restype = C.c_int64
        L.zo_row_bytes.argtypes = [C.c_int, C.c_int64]
        L.zo_dequant.argtypes = [C.c_int, u8p, f32p, C.c_int64]
        L.zo_gemv.argtypes = [C.c_int, u8p, C.c_int64, C.c_int64, f32p, f32p]
        L.zo_gemv_f64.argtypes = [C.c_int, u8p, C.c_int64, C.c_int64, f32p, C.POINTER(C.c_double)]
        L.zo_gemm_nt.argtypes = [C.c_int, u8p, C.c_int64, C.c_int64, f32p, C.c_int64, f32p]
        L.zo_rmsnorm.restype = C.c_float
        L.zo_rmsnorm.argtypes = [f32p, f32p, f32p, C.c_int, C.c_float]
        L.zo_add_rmsnorm.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_int, C.c_float]
        L.zo_norm_add.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_float]
        L.zo_silu.argtypes = [f32p, f32p, C.c_int]
        L.zo_swiglu.argtypes = [f32p, f32p, f32p, C.c_int]
        L.zo_softmax.argtypes = [f32p, C.c_int]
        L.zo_rope.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int]
        L.zo_rope_tables.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_double]
        L.zo_softcap.argtypes = [f32p, C.c_int, C.c_float]
        L.zo_argmax.restype = C.c_int
        L.zo_argmax.argtypes = [f32p, C.c_int]
        L.zo_attn_decode_head.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int64, C.c_float, f32p]
        L.zo_attn_causal_head.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_float, f32p]
        L.zo_moe_route.argtypes = [f32p, C.c_int, C.c_int, i32p, f32p]
        L.zo_model_load.restype = C.c_void_p
        L.zo_model_load.argtypes = [C.c_char_p, C.c_int]
        L.zo_model_free.argtypes = [C.c_void_p]
        L.zo_model_reset.argtypes = [C.c_void_p]
        L.zo_model_pos.restype = C.c_int
        L.zo_model_pos.argtypes = [C.c_void_p]
        L.zo_model_dims.argtypes = [C.c_void_p, i32p]
        L.zo_model_logits.restype = f32p
        L.zo_model_logits.argtypes = [C.c_void_p]
        L.zo_model_hidden.restype = f32p
        L.zo_model_hidden.argtypes = [C.c_void_p]
        L.zo_model_kcache.restype = f32p
        L.zo_model_kcache.argtypes = [C.c_void_p, C.c_int]
        L.zo_model_vcache.restype = f32p
        L.zo_model_vcache.argtypes = [C.c_void_p, C.c_int]
        L.zo_model_forward.restype = C.c_int
        L.zo_model_forward.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.zo_model_generate.restype = C.c_int
        L.zo_model_generate.argtypes = [C.c_void_p, i32p, C.c_int, C.c_int, i32p]
        L.zo_num_threads.restype = C.c_int
        L.zo_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _f32(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _raw(a: np.ndarray):
    a = np.ascontiguousarray(a)
    return a, C.c_void_p(a.ctypes.data)


def fp16_to_f32(bits: int) -> float:
    return float(lib().zo_fp16_to_f32(bits))


def dequant(qtype: int, raw: np.ndarray, n: int) -> np.ndarray:
    keep, p = _raw(raw)
    out = np.empty(n, dtype=np.float32)
    if lib().zo_dequant(qtype, p, _f32(out), n):
        raise ValueError("zo_dequant: bad type or length")
    return out


def gemv(qtype: int, raw: np.ndarray, rows: int, k: int, x: np.ndarray) -> np.ndarray:
    keep, p = _raw(raw)
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty(rows, dtype=np.float32)
    if lib().zo_gemv(qtype, p, rows, k, _f32(x), _f32(y)):
        raise ValueError("zo_gemv: bad type or K")
    return y


def gemv_f64(qtype: int, raw: np.ndarray, rows: int, k: int, x: np.ndarray) -> np.ndarray:
    keep, p = _raw(raw)
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty(rows, dtype=np.float64)
    if lib().zo_gemv_f64(qtype, p, rows, k, _f32(x), y.ctypes.data_as(C.POINTER(C.c_double))):
        raise ValueError("zo_gemv_f64: bad type or K")
    return y


def gemm_nt(qtype: int, raw: np.ndarray, n: int, k: int, x: np.ndarray) -> np.ndarray:
    keep, p = _raw(raw)
    x = np.ascontiguousarray(x, dtype=np.float32)
    m = x.shape[0]
    c = np.empty((m, n), dtype=np.float32)
    if lib().zo_gemm_nt(qtype, p, n, k, _f32(x), m, _f32(c)):
        raise ValueError("zo_gemm_nt: bad type or K")
    return c


def rmsnorm(x: np.ndarray, w: np.ndarray, eps: float) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    out = np.empty_like(x)
    x2, o2 = x.reshape(-1, x.shape[-1]), out.reshape(-1, x.shape[-1])
    for i in range(x2.shape[0]):
        lib().zo_rmsnorm(_f32(o2[i]), _f32(x2[i]), _f32(w), x.shape[-1], eps)
    return out


def add_rmsnorm(a: np.ndarray, r: np.ndarray, w: np.ndarray, eps: float):
    a = np.ascontiguousarray(a, dtype=np.float32)
    r = np.ascontiguousarray(r, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    normed, s = np.empty_like(a), np.empty_like(a)
    d = a.shape[-1]
    for i in range(a.reshape(-1, d).shape[0]):
        lib().zo_add_rmsnorm(_f32(normed.reshape(-1, d)[i]), _f32(s.reshape(-1, d)[i]), _f32(a.reshape(-1, d)[i]),
                             _f32(r.reshape(-1, d)[i]), _f32(w), d, eps)
    return normed, s


def norm_add(x: np.ndarray, w: np.ndarray, r: np.ndarray, eps: float) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    r = np.ascontiguousarray(r, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    out = np.empty_like(x)
    d = x.shape[-1]
    for i in range(x.reshape(-1, d).shape[0]):
        lib().zo_norm_add(_f32(out.reshape(-1, d)[i]), _f32(x.reshape(-1, d)[i]), _f32(w), _f32(r.reshape(-1, d)[i]), d, eps)
    return out


def silu(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().zo_silu(_f32(out.reshape(-1)), _f32(x.reshape(-1)), x.size)
    return out


def swiglu(gate: np.ndarray, up: np.ndarray) -> np.ndarray:
    gate = np.ascontiguousarray(gate, dtype=np.float32)
    up = np.ascontiguousarray(up, dtype=np.float32)
    out = np.empty_like(gate)
    lib().zo_swiglu(_f32(out.reshape(-1)), _f32(gate.reshape(-1)), _f32(up.reshape(-1)), gate.size)
    return out


def softmax(x: np.ndarray) -> np.ndarray:
    out = np.array(x, dtype=np.float32, copy=True)
    o2 = out.reshape(-1, out.shape[-1])
    for i in range(o2.shape[0]):
        lib().zo_softmax(_f32(o2[i]), o2.shape[1])
    return out


def rope_tables(positions: int, rotary_dim: int, base: float):
    half = rotary_dim // 2
    cs = np.empty((positions, half), dtype=np.float32)
    sn = np.empty((positions, half), dtype=np.float32)
    lib().zo_rope_tables(_f32(cs.reshape(-1)), _f32(sn.reshape(-1)), positions, rotary_dim, base)
    return cs, sn


def rope(x: np.ndarray, cs: np.ndarray, sn: np.ndarray) -> np.ndarray:
    """x: [..., head_dim]; cs/sn: [half] for one position."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    cs = np.ascontiguousarray(cs, dtype=np.float32)
    sn = np.ascontiguousarray(sn, dtype=np.float32)
    out = np.empty_like(x)
    hd = x.shape[-1]
    for i in range(x.reshape(-1, hd).shape[0]):
        lib().zo_rope(_f32(out.reshape(-1, hd)[i]), _f32(x.reshape(-1, hd)[i]), _f32(cs), _f32(sn), cs.size, hd)
    return out


def softcap(logits: np.ndarray, cap: float) -> np.ndarray:
    out = np.array(logits, dtype=np.float32, copy=True)
    lib().zo_softcap(_f32(out.reshape(-1)), out.size, cap)
    return out


def argmax(x: np.ndarray) -> int:
    x = np.ascontiguousarray(x, dtype=np.float32)
    return int(lib().zo_argmax(_f32(x), x.size))


def attn_decode(q: np.ndarray, k: np.ndarray, v: np.ndarray, n_kv: int, kv_len: int) -> np.ndarray:
    """q: [nQ, hd]; k, v: [max_kv, nKV*hd] (flash_decode layout).  Returns [nQ, hd]."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    k = np.ascontiguousarray(k, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    nq, hd = q.shape
    rep = nq // n_kv
    out = np.empty_like(q)
    scratch = np.empty(kv_len, dtype=np.float32)
    scale = np.float32(1.0 / np.sqrt(float(hd)))
    stride = k.shape[-1]
    for h in range(nq):
        kvh = h // rep
        lib().zo_attn_decode_head(_f32(out[h]), _f32(q[h]),
                                  C.cast(k.ctypes.data + 4 * kvh * hd, C.POINTER(C.c_float)),
                                  C.cast(v.ctypes.data + 4 * kvh * hd, C.POINTER(C.c_float)),
                                  kv_len, hd, stride, scale, _f32(scratch))
    return out


def attn_causal(q: np.ndarray, k: np.ndarray, v: np.ndarray) -> np.ndarray:
    """q,k,v: [heads, seq, hd] (heads already matched).  Returns [heads, seq, hd]."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    k = np.ascontiguousarray(k, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    h, s, hd = q.shape
    out = np.empty_like(q)
    scratch = np.empty(s, dtype=np.float32)
    scale = np.float32(1.0 / np.sqrt(float(hd)))
    for i in range(h):
        lib().zo_attn_causal_head(_f32(out[i]), _f32(q[i]), _f32(k[i]), _f32(v[i]), s, hd, scale, _f32(scratch))
    return out


def moe_route(logits: np.ndarray, top_k: int):
    """logits [E] -> (expert indices [K], renormalised weights [K])."""
    p = np.array(logits, dtype=np.float32, copy=True)
    idx = (C.c_int * top_k)()
    w = np.empty(top_k, dtype=np.float32)
    lib().zo_moe_route(_f32(p), p.size, top_k, idx, _f32(w))
    return list(idx), w


class Model:
    """The reference CPU engine's decode loop, restated (zo_model_* in zoracle.c)."""

    def __init__(self, path: str, max_seq: int = 0):
        self._h = lib().zo_model_load(path.encode(), max_seq)
        if not self._h:
            raise RuntimeError(f"oracle: cannot load {path}")
        d = (C.c_int * 10)()
        lib().zo_model_dims(self._h, d)
        (self.vocab, self.hidden, self.layers, self.n_q, self.n_kv, self.head_dim, self.ffn, self.max_seq,
         self.n_experts, self.top_k) = list(d)

    def close(self):
        if self._h:
            lib().zo_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        lib().zo_model_reset(self._h)

    @property
    def pos(self) -> int:
        return lib().zo_model_pos(self._h)

    def forward(self, token: int, want_logits: bool = True):
        rc = lib().zo_model_forward(self._h, int(token), int(want_logits))
        if rc:
            raise RuntimeError(f"oracle forward failed rc={rc}")
        if want_logits:
            return np.ctypeslib.as_array(lib().zo_model_logits(self._h), shape=(self.vocab,)).copy()
        return None

    def set_kv_f16(self, on: bool = True) -> None:
        """Store K / V rounded to fp16 (the reference's kvFP16 cache, generate/tensor_cache.go:224-238)."""
        L = lib()
        L.zo_model_set_kv_f16.argtypes = [C.c_void_p, C.c_int]
        L.zo_model_set_kv_f16.restype = None
        L.zo_model_set_kv_f16(self._h, int(on))

    def prefill(self, prompt: Sequence[int], want_logits: bool = True):
        """The prompt pass (one Forward of seqLen = n in the reference): a sliding-window model masks i - j >= window here,
        and only here (grouped_query_attention.go:1074-1077)."""
        L = lib()
        L.zo_model_prefill.restype = C.c_int
        L.zo_model_prefill.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int]
        p = (C.c_int * len(prompt))(*prompt)
        if L.zo_model_prefill(self._h, p, len(prompt), int(want_logits)):
            raise RuntimeError("oracle prefill failed")
        if want_logits:
            return np.ctypeslib.as_array(L.zo_model_logits(self._h), shape=(self.vocab,)).copy()
        return None

    def hidden_state(self) -> np.ndarray:
        return np.ctypeslib.as_array(lib().zo_model_hidden(self._h), shape=(self.hidden,)).copy()

    def kv(self, layer: int, n: int):
        d = self.n_kv * self.head_dim
        k = np.ctypeslib.as_array(lib().zo_model_kcache(self._h, layer), shape=(self.max_seq, d))[:n].copy()
        v = np.ctypeslib.as_array(lib().zo_model_vcache(self._h, layer), shape=(self.max_seq, d))[:n].copy()
        return k, v

    def generate(self, prompt: Sequence[int], n_new: int) -> List[int]:
        p = (C.c_int * len(prompt))(*prompt)
        out = (C.c_int * n_new)()
        n = lib().zo_model_generate(self._h, p, len(prompt), n_new, out)
        if n < 0:
            raise RuntimeError("oracle generate failed")
        return list(out[:n])


def num_threads() -> int:
    return int(lib().zo_num_threads())


def set_num_threads(n: int) -> None:
    """omp_set_num_threads for the oracle's row-parallel loops (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    L = lib()
    L.zo_set_num_threads.argtypes = [C.c_int]
    L.zo_set_num_threads.restype = None
    L.zo_set_num_threads(int(n))


def set_threads(n: int) -> None:
    lib().zo_set_threads(n)
