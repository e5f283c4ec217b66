"""GGUF v3 writer/reader and block quantizers for synthetic-weight models.

Host-side utility (numpy only).  It produces the *inputs* of the hot path:
GGUF files in the shape of the reference's ``writeTestGGUF``
(inference/load_gguf_test.go:64-140) holding already-quantized native blocks,
so neither the reference's load-time requantisation (model/gguf/loader.go:285-346)
nor the absent ``tensor.QuantizeQ4`` rounding matters.

Block layouts follow model/gguf/loader.go:140-190 and the kernel headers
internal/cuda/kernels/gemv_q4k.cu:8-21, gemv_q5k.cu:7-23, gemv_q6k.cu:7-25.
Quantizers restate the reference's in-test builders:
  * Q4_0: internal/cuda/kernels/gemm_q4_test.go:14-85 (scale=absmax/7,
    round-half-away, clamp [-8,7], low nibble = element j, high = j+16)
  * Q4_K: internal/cuda/kernels/gemv_q4k_test.go:94-196 (per-sub-block
    min/max, 6-bit scales/mins against d=max/63)
Q5_K / Q6_K / Q8_0 use the same min/max (resp. absmax) scheme on their
published layouts.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

GGUF_MAGIC = 0x46554747
ALIGN = 32

F32, F16, Q4_0, Q8_0, Q4_K, Q5_K, Q6_K = 0, 1, 2, 8, 12, 13, 14
TYPE_NAMES = {F32: "F32", F16: "F16", Q4_0: "Q4_0", Q8_0: "Q8_0", Q4_K: "Q4_K", Q5_K: "Q5_K", Q6_K: "Q6_K"}
BLOCK_ELEMS = {F32: 1, F16: 1, Q4_0: 32, Q8_0: 32, Q4_K: 256, Q5_K: 256, Q6_K: 256}
BLOCK_BYTES = {F32: 4, F16: 2, Q4_0: 18, Q8_0: 34, Q4_K: 144, Q5_K: 176, Q6_K: 210}

# GGUF metadata value types
_U32, _I32, _F32, _BOOL, _STR, _ARR, _U64 = 4, 5, 6, 7, 8, 9, 10


def row_bytes(qtype: int, k: int) -> int:
    be = BLOCK_ELEMS[qtype]
    if k % be:
        raise ValueError(f"K={k} is not a multiple of the {TYPE_NAMES[qtype]} block size {be}")
    return k // be * BLOCK_BYTES[qtype]


def _round_half_away(x: np.ndarray) -> np.ndarray:
    # Go's math.Round: half away from zero.
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def _f16_bytes(x: np.ndarray) -> np.ndarray:
    return x.astype(np.float16).view(np.uint8).reshape(*x.shape, 2)


# ----------------------------------------------------------------------------
# Quantizers: float32 [rows, K] -> uint8 [rows, row_bytes]
# ----------------------------------------------------------------------------

def quantize_q4_0(w: np.ndarray) -> np.ndarray:
    w = np.ascontiguousarray(w, dtype=np.float32)
    rows, k = w.shape
    b = w.reshape(rows, k // 32, 32)
    absmax = np.abs(b).max(axis=-1)
    scale = (absmax / np.float32(7.0)).astype(np.float32)
    inv = np.where(scale > 0, np.float32(1.0) / np.where(scale > 0, scale, 1), 0).astype(np.float32)
    q = _round_half_away(b * inv[..., None])
    q = (np.clip(q, -8, 7).astype(np.int8) + 8).astype(np.uint8)
    out = np.empty((rows, k // 32, 18), dtype=np.uint8)
    out[..., 0:2] = _f16_bytes(scale)
    out[..., 2:] = (q[..., :16] | (q[..., 16:] << 4)).astype(np.uint8)
    return out.reshape(rows, -1)


def quantize_q8_0(w: np.ndarray) -> np.ndarray:
    w = np.ascontiguousarray(w, dtype=np.float32)
    rows, k = w.shape
    b = w.reshape(rows, k // 32, 32)
    absmax = np.abs(b).max(axis=-1)
    scale = (absmax / np.float32(127.0)).astype(np.float32)
    inv = np.where(scale > 0, np.float32(1.0) / np.where(scale > 0, scale, 1), 0).astype(np.float32)
    q = np.clip(_round_half_away(b * inv[..., None]), -127, 127).astype(np.int8)
    out = np.empty((rows, k // 32, 34), dtype=np.uint8)
    out[..., 0:2] = _f16_bytes(scale)
    out[..., 2:] = q.view(np.uint8)
    return out.reshape(rows, -1)


def _kquant_scales(b: np.ndarray, levels: int):
    """Shared Q4_K/Q5_K sub-block statistics. b: [rows, nb, 8, 32]."""
    mn = np.minimum(b.min(axis=-1), 0)
    mx = b.max(axis=-1)
    sub_scale = ((mx - mn) / np.float32(levels)).astype(np.float32)
    sub_min = (-mn).astype(np.float32)
    d = (sub_scale.max(axis=-1) / np.float32(63.0)).astype(np.float32)
    dmin = (sub_min.max(axis=-1) / np.float32(63.0)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        sq = np.where(d[..., None] > 0, _round_half_away(sub_scale / np.where(d > 0, d, 1)[..., None]), 0)
        mq = np.where(dmin[..., None] > 0, _round_half_away(sub_min / np.where(dmin > 0, dmin, 1)[..., None]), 0)
    sq = np.clip(sq, 0, 63).astype(np.uint8)
    mq = np.clip(mq, 0, 63).astype(np.uint8)
    d16 = d.astype(np.float16)
    dm16 = dmin.astype(np.float16)
    sc = d16.astype(np.float32)[..., None] * sq.astype(np.float32)
    mnv = dm16.astype(np.float32)[..., None] * mq.astype(np.float32)
    inv = np.where(sc > 0, np.float32(1.0) / np.where(sc > 0, sc, 1), 0).astype(np.float32)
    q = _round_half_away((b + mnv[..., None]) * inv[..., None])
    q = np.clip(q, 0, levels).astype(np.uint8)
    return d16, dm16, sq, mq, q


def _pack_k_scales(sq: np.ndarray, mq: np.ndarray) -> np.ndarray:
    out = np.empty(sq.shape[:-1] + (12,), dtype=np.uint8)
    out[..., 0:4] = (sq[..., 0:4] & 63) | ((sq[..., 4:8] >> 4) << 6)
    out[..., 4:8] = (mq[..., 0:4] & 63) | ((mq[..., 4:8] >> 4) << 6)
    out[..., 8:12] = (sq[..., 4:8] & 0xF) | ((mq[..., 4:8] & 0xF) << 4)
    return out


def quantize_q4_k(w: np.ndarray) -> np.ndarray:
    w = np.ascontiguousarray(w, dtype=np.float32)
    rows, k = w.shape
    nb = k // 256
    b = w.reshape(rows, nb, 8, 32)
    d16, dm16, sq, mq, q = _kquant_scales(b, 15)
    out = np.empty((rows, nb, 144), dtype=np.uint8)
    out[..., 0:2] = d16.view(np.uint8).reshape(rows, nb, 2)
    out[..., 2:4] = dm16.view(np.uint8).reshape(rows, nb, 2)
    out[..., 4:16] = _pack_k_scales(sq, mq)
    qg = q.reshape(rows, nb, 4, 2, 32)
    out[..., 16:] = (qg[..., 0, :] | (qg[..., 1, :] << 4)).reshape(rows, nb, 128)
    return out.reshape(rows, -1)


def quantize_q5_k(w: np.ndarray) -> np.ndarray:
    w = np.ascontiguousarray(w, dtype=np.float32)
    rows, k = w.shape
    nb = k // 256
    b = w.reshape(rows, nb, 8, 32)
    d16, dm16, sq, mq, q = _kquant_scales(b, 31)
    out = np.empty((rows, nb, 176), dtype=np.uint8)
    out[..., 0:2] = d16.view(np.uint8).reshape(rows, nb, 2)
    out[..., 2:4] = dm16.view(np.uint8).reshape(rows, nb, 2)
    out[..., 4:16] = _pack_k_scales(sq, mq)
    qg = q.reshape(rows, nb, 4, 2, 32)
    lo4 = qg & 0xF
    out[..., 16:144] = (lo4[..., 0, :] | (lo4[..., 1, :] << 4)).reshape(rows, nb, 128)
    hb = (qg >> 4).astype(np.uint8)  # [rows, nb, 4, 2, 32]; bit position 2g + j
    qh = np.zeros((rows, nb, 32), dtype=np.uint8)
    for g in range(4):
        for j in range(2):
            qh |= hb[:, :, g, j, :] << (2 * g + j)
    out[..., 144:] = qh
    return out.reshape(rows, -1)


def quantize_q6_k(w: np.ndarray) -> np.ndarray:
    w = np.ascontiguousarray(w, dtype=np.float32)
    rows, k = w.shape
    nb = k // 256
    b = w.reshape(rows, nb, 16, 16)
    absmax = np.abs(b).max(axis=-1)
    sub = (absmax / np.float32(32.0)).astype(np.float32)
    d = (sub.max(axis=-1) / np.float32(127.0)).astype(np.float32)
    d16 = d.astype(np.float16)
    with np.errstate(divide="ignore", invalid="ignore"):
        sc = np.where(d[..., None] > 0, _round_half_away(sub / np.where(d > 0, d, 1)[..., None]), 0)
    sc = np.clip(sc, -128, 127).astype(np.int8)
    eff = d16.astype(np.float32)[..., None] * sc.astype(np.float32)
    inv = np.where(eff != 0, np.float32(1.0) / np.where(eff != 0, eff, 1), 0).astype(np.float32)
    q = np.clip(_round_half_away(b * inv[..., None]), -32, 31).astype(np.int16) + 32
    q = q.astype(np.uint8).reshape(rows, nb, 2, 4, 32)  # [half, quarter(0..3), l]
    out = np.empty((rows, nb, 210), dtype=np.uint8)
    ql = np.empty((rows, nb, 2, 64), dtype=np.uint8)
    ql[..., 0:32] = (q[..., 0, :] & 0xF) | ((q[..., 2, :] & 0xF) << 4)
    ql[..., 32:64] = (q[..., 1, :] & 0xF) | ((q[..., 3, :] & 0xF) << 4)
    qh = (q[..., 0, :] >> 4) | ((q[..., 1, :] >> 4) << 2) | ((q[..., 2, :] >> 4) << 4) | ((q[..., 3, :] >> 4) << 6)
    out[..., 0:128] = ql.reshape(rows, nb, 128)
    out[..., 128:192] = qh.reshape(rows, nb, 64)
    out[..., 192:208] = sc.view(np.uint8)
    out[..., 208:210] = d16.view(np.uint8).reshape(rows, nb, 2)
    return out.reshape(rows, -1)


QUANTIZERS = {Q4_0: quantize_q4_0, Q8_0: quantize_q8_0, Q4_K: quantize_q4_k, Q5_K: quantize_q5_k, Q6_K: quantize_q6_k}


def quantize(w: np.ndarray, qtype: int) -> np.ndarray:
    """float32 [rows, K] -> raw block bytes uint8 [rows, row_bytes]."""
    if qtype == F32:
        return np.ascontiguousarray(w, dtype=np.float32).view(np.uint8).reshape(w.shape[0], -1)
    if qtype == F16:
        return np.ascontiguousarray(w.astype(np.float16)).view(np.uint8).reshape(w.shape[0], -1)
    return QUANTIZERS[qtype](w)


def q4_0_to_separated(raw: np.ndarray) -> Tuple[np.ndarray, int]:
    """GGUF-interleaved Q4_0 blocks -> the reference's GPU "separated" layout
    ``[scales: nblk*2 B][pad to 16 B][data: nblk*16 B]`` (gemm_q4.h:3-4,
    gemm_q4_test.go:87-103).  Returns (bytes, data_offset)."""
    blocks = raw.reshape(-1, 18)
    n = blocks.shape[0]
    pad = (n * 2 + 15) & ~15
    out = np.zeros(pad + n * 16, dtype=np.uint8)
    out[: n * 2] = blocks[:, :2].reshape(-1)
    out[pad:] = blocks[:, 2:].reshape(-1)
    return out, pad


def q8_0_to_zerfoo36(raw: np.ndarray) -> np.ndarray:
    """GGUF 34 B Q8_0 blocks -> the reference's device layout: f32 scale +
    32 int8 = 36 B (gemm_q8.cu:1-7, model/gguf/loader.go:459-502)."""
    blocks = raw.reshape(-1, 34)
    n = blocks.shape[0]
    out = np.empty((n, 36), dtype=np.uint8)
    scale = np.ascontiguousarray(blocks[:, :2]).view(np.float16).astype(np.float32)
    out[:, :4] = scale.view(np.uint8).reshape(n, 4)
    out[:, 4:] = blocks[:, 2:]
    return out.reshape(-1)


# ----------------------------------------------------------------------------
# Writer
# ----------------------------------------------------------------------------

def _enc_str(s: str) -> bytes:
    b = s.encode("utf-8")
    return struct.pack("<Q", len(b)) + b


def _enc_kv(key: str, value) -> bytes:
    out = _enc_str(key)
    if isinstance(value, bool):
        out += struct.pack("<IB", _BOOL, int(value))
    elif isinstance(value, int):
        out += struct.pack("<II", _U32, value)
    elif isinstance(value, float):
        out += struct.pack("<If", _F32, value)
    elif isinstance(value, str):
        out += struct.pack("<I", _STR) + _enc_str(value)
    elif isinstance(value, (list, tuple)):
        if all(isinstance(v, str) for v in value):
            out += struct.pack("<IIQ", _ARR, _STR, len(value)) + b"".join(_enc_str(v) for v in value)
        elif all(isinstance(v, int) for v in value):
            out += struct.pack("<IIQ", _ARR, _I32, len(value)) + struct.pack(f"<{len(value)}i", *value)
        else:
            out += struct.pack("<IIQ", _ARR, _F32, len(value)) + struct.pack(f"<{len(value)}f", *value)
    else:
        raise TypeError(f"unsupported metadata value for {key}: {type(value)}")
    return out


@dataclass
class TensorSpec:
    name: str
    qtype: int
    ne: Tuple[int, ...]          # GGML order, innermost first
    producer: object = None      # callable() -> np.uint8 array of raw bytes

    @property
    def nbytes(self) -> int:
        n = 1
        for d in self.ne:
            n *= d
        return n // BLOCK_ELEMS[self.qtype] * BLOCK_BYTES[self.qtype]


class GGUFWriter:
    """Streams tensors to disk so multi-GB synthetic models never sit in RAM."""

    def __init__(self, path: str, metadata: Dict[str, object]):
        self.path = path
        self.metadata = dict(metadata)
        self.tensors: List[TensorSpec] = []

    def add(self, name: str, qtype: int, ne: Tuple[int, ...], producer) -> None:
        self.tensors.append(TensorSpec(name, qtype, tuple(int(d) for d in ne), producer))

    def add_array(self, name: str, raw: np.ndarray, qtype: int, ne: Tuple[int, ...]) -> None:
        raw = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
        self.add(name, qtype, ne, lambda raw=raw: raw)

    def write(self) -> None:
        head = struct.pack("<IIQQ", GGUF_MAGIC, 3, len(self.tensors), len(self.metadata))
        for k, v in self.metadata.items():
            head += _enc_kv(k, v)
        off = 0
        for t in self.tensors:
            head += _enc_str(t.name) + struct.pack("<I", len(t.ne)) + struct.pack(f"<{len(t.ne)}Q", *t.ne)
            head += struct.pack("<IQ", t.qtype, off)
            off += (t.nbytes + ALIGN - 1) // ALIGN * ALIGN
        pad = (-len(head)) % ALIGN
        with open(self.path, "wb") as f:
            f.write(head + b"\0" * pad)
            for t in self.tensors:
                raw = np.ascontiguousarray(t.producer()).view(np.uint8).reshape(-1)
                if raw.nbytes != t.nbytes:
                    raise ValueError(f"{t.name}: produced {raw.nbytes} bytes, expected {t.nbytes}")
                raw.tofile(f)
                f.write(b"\0" * ((-t.nbytes) % ALIGN))


# ----------------------------------------------------------------------------
# Reader (numpy memmap; used by tests and by host-side tooling)
# ----------------------------------------------------------------------------

@dataclass
class GGUFTensor:
    name: str
    qtype: int
    ne: Tuple[int, ...]
    data: np.ndarray  # uint8 view

    @property
    def rows(self) -> int:
        n = 1
        for d in self.ne[1:]:
            n *= d
        return n

    @property
    def cols(self) -> int:
        return self.ne[0]


@dataclass
class GGUFFile:
    metadata: Dict[str, object] = field(default_factory=dict)
    tensors: Dict[str, GGUFTensor] = field(default_factory=dict)


def read_gguf(path: str) -> GGUFFile:
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    buf = memoryview(mm)
    pos = 0

    def take(fmt):
        nonlocal pos
        v = struct.unpack_from(fmt, buf, pos)
        pos += struct.calcsize(fmt)
        return v if len(v) > 1 else v[0]

    def take_str():
        nonlocal pos
        n = take("<Q")
        s = bytes(buf[pos:pos + n]).decode("utf-8")
        pos += n
        return s

    scal = {0: "<B", 1: "<b", 2: "<H", 3: "<h", 4: "<I", 5: "<i", 6: "<f", 7: "<B", 10: "<Q", 11: "<q", 12: "<d"}

    def take_val(t):
        if t == _STR:
            return take_str()
        if t == _ARR:
            et, cnt = take("<I"), take("<Q")
            return [take_val(et) for _ in range(cnt)]
        return take(scal[t])

    magic, ver, nt, nkv = take("<I"), take("<I"), take("<Q"), take("<Q")
    if magic != GGUF_MAGIC or ver not in (2, 3):
        raise ValueError("not a GGUF v2/v3 file")
    out = GGUFFile()
    for _ in range(nkv):
        k = take_str()
        out.metadata[k] = take_val(take("<I"))
    infos = []
    for _ in range(nt):
        name = take_str()
        nd = take("<I")
        ne = tuple(take("<Q") for _ in range(nd))
        qt, off = take("<I"), take("<Q")
        infos.append((name, qt, ne, off))
    align = int(out.metadata.get("general.alignment", ALIGN))
    base = (pos + align - 1) // align * align
    for name, qt, ne, off in infos:
        n = 1
        for d in ne:
            n *= d
        nbytes = n // BLOCK_ELEMS[qt] * BLOCK_BYTES[qt]
        out.tensors[name] = GGUFTensor(name, qt, ne, mm[base + off: base + off + nbytes])
    return out


# ----------------------------------------------------------------------------
# Synthetic decoder models (SURVEY 8d "Synthetic inputs")
# ----------------------------------------------------------------------------

@dataclass
class ModelSpec:
    arch: str
    vocab: int
    hidden: int
    layers: int
    n_q: int
    n_kv: int
    head_dim: int
    ffn: int
    ctx: int = 2048
    rope_base: float = 10000.0
    rope_local_base: float = 0.0
    eps: float = 1e-6
    softcap: float = 0.0
    sliding_window: int = 0
    n_experts: int = 0
    top_k: int = 0
    tied: bool = True
    base_type: int = Q4_0
    more_bits_type: Optional[int] = None   # attn_v / ffn_down on "more bits" layers, and output
    embed_type: Optional[int] = None
    name: str = "synthetic"

    def use_more_bits(self, i: int) -> bool:
        n = self.layers
        return i < n // 8 or i >= 7 * n // 8 or (i - n // 8) % 3 == 2


def preset(name: str, layers: Optional[int] = None, vocab: Optional[int] = None, ctx: Optional[int] = None) -> ModelSpec:
    """BASELINE.json configs (SURVEY 8 shapes).  ``layers``/``vocab`` override for reduced parity cases."""
    p = {
        "c1": ModelSpec("gemma3", 262144, 1152, 26, 4, 1, 256, 6912, ctx=2048, rope_base=1e6, rope_local_base=1e4,
                        eps=1e-6, softcap=30.0, sliding_window=512, tied=True, base_type=Q4_0, name="gemma3-1b-shape-q4_0"),
        "c2": ModelSpec("llama", 128256, 3072, 28, 24, 8, 128, 8192, ctx=4096, rope_base=5e5, eps=1e-5, tied=True,
                        base_type=Q4_K, more_bits_type=Q6_K, embed_type=Q6_K, name="llama3.2-3b-shape-q4_k_m"),
        "c3": ModelSpec("mistral", 32000, 4096, 32, 32, 8, 128, 14336, ctx=8192, rope_base=1e6, eps=1e-5, tied=False,
                        base_type=Q5_K, more_bits_type=Q6_K, embed_type=Q5_K, name="mistral-7b-shape-q5_k_m"),
        "c4": ModelSpec("llama", 128256, 8192, 80, 64, 8, 128, 28672, ctx=4096, rope_base=5e5, eps=1e-5, tied=False,
                        base_type=Q4_K, more_bits_type=Q6_K, embed_type=Q4_K, name="llama3-70b-shape-q4_k_m"),
        "c5": ModelSpec("mixtral", 32000, 4096, 32, 32, 8, 128, 14336, ctx=4096, rope_base=1e6, eps=1e-5, tied=False,
                        n_experts=8, top_k=2, base_type=Q4_K, more_bits_type=Q6_K, embed_type=Q4_K,
                        name="mixtral-8x7b-shape-q4_k_m"),
    }[name]
    if layers is not None:
        p.layers = layers
    if vocab is not None:
        p.vocab = vocab
    if ctx is not None:
        p.ctx = ctx
    return p


def model_metadata(s: ModelSpec) -> Dict[str, object]:
    a = s.arch
    md: Dict[str, object] = {
        "general.architecture": a,
        "general.name": s.name,
        "general.alignment": ALIGN,
        f"{a}.vocab_size": s.vocab,
        f"{a}.embedding_length": s.hidden,
        f"{a}.block_count": s.layers,
        f"{a}.attention.head_count": s.n_q,
        f"{a}.attention.head_count_kv": s.n_kv,
        f"{a}.feed_forward_length": s.ffn,
        f"{a}.context_length": s.ctx,
        f"{a}.rope.freq_base": float(s.rope_base),
        f"{a}.attention.key_length": s.head_dim,
        f"{a}.attention.value_length": s.head_dim,
        f"{a}.attention.layer_norm_rms_epsilon": float(s.eps),
    }
    if s.rope_local_base:
        md[f"{a}.rope.local.freq_base"] = float(s.rope_local_base)
    if s.softcap:
        md[f"{a}.final_logit_softcapping"] = float(s.softcap)
    if s.sliding_window:
        md[f"{a}.attention.sliding_window"] = s.sliding_window
    if s.n_experts:
        md[f"{a}.expert_count"] = s.n_experts
        md[f"{a}.expert_used_count"] = s.top_k
    # gpt2-style tokenizer stub as in writeTestGGUF (load_gguf_test.go:100-118);
    # token ids are fed directly, the tokenizer is not on the hot path.
    md["tokenizer.ggml.model"] = "gpt2"
    md["tokenizer.ggml.bos_token_id"] = 1
    md["tokenizer.ggml.eos_token_id"] = 2
    return md


def tensor_plan(s: ModelSpec) -> List[Tuple[str, int, Tuple[int, ...], str]]:
    """[(name, qtype, ne(GGML order), kind)] with kind in {'matmul','norm','router'}."""
    qd, kvd = s.n_q * s.head_dim, s.n_kv * s.head_dim
    mb = s.more_bits_type
    plan: List[Tuple[str, int, Tuple[int, ...], str]] = []
    et = s.embed_type if s.embed_type is not None else s.base_type
    plan.append(("token_embd.weight", et, (s.hidden, s.vocab), "matmul"))
    for i in range(s.layers):
        p = f"blk.{i}."
        more = mb is not None and s.use_more_bits(i)
        plan.append((p + "attn_norm.weight", F32, (s.hidden,), "norm"))
        plan.append((p + "attn_q.weight", s.base_type, (s.hidden, qd), "matmul"))
        plan.append((p + "attn_k.weight", s.base_type, (s.hidden, kvd), "matmul"))
        plan.append((p + "attn_v.weight", mb if more else s.base_type, (s.hidden, kvd), "matmul"))
        plan.append((p + "attn_output.weight", s.base_type, (qd, s.hidden), "matmul"))
        if s.arch == "gemma3":
            plan.append((p + "attn_q_norm.weight", F32, (s.head_dim,), "norm"))
            plan.append((p + "attn_k_norm.weight", F32, (s.head_dim,), "norm"))
            plan.append((p + "post_attention_norm.weight", F32, (s.hidden,), "norm"))
        plan.append((p + "ffn_norm.weight", F32, (s.hidden,), "norm"))
        if s.n_experts:
            plan.append((p + "ffn_gate_inp.weight", F32, (s.hidden, s.n_experts), "router"))
            plan.append((p + "ffn_gate_exps.weight", s.base_type, (s.hidden, s.ffn, s.n_experts), "matmul"))
            plan.append((p + "ffn_up_exps.weight", s.base_type, (s.hidden, s.ffn, s.n_experts), "matmul"))
            plan.append((p + "ffn_down_exps.weight", mb if more else s.base_type, (s.ffn, s.hidden, s.n_experts), "matmul"))
        else:
            plan.append((p + "ffn_gate.weight", s.base_type, (s.hidden, s.ffn), "matmul"))
            plan.append((p + "ffn_up.weight", s.base_type, (s.hidden, s.ffn), "matmul"))
            plan.append((p + "ffn_down.weight", mb if more else s.base_type, (s.ffn, s.hidden), "matmul"))
        if s.arch == "gemma3":
            plan.append((p + "post_ffw_norm.weight", F32, (s.hidden,), "norm"))
    plan.append(("output_norm.weight", F32, (s.hidden,), "norm"))
    if not s.tied:
        plan.append(("output.weight", mb if mb is not None else s.base_type, (s.hidden, s.vocab), "matmul"))
    return plan


_POOL = None


def _pool():
    global _POOL
    if _POOL is None:
        import concurrent.futures
        import os
        _POOL = concurrent.futures.ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 4))
    return _POOL


def _produce(idx: int, qtype: int, ne: Tuple[int, ...], kind: str, seed: int, sigma: float, chunk_rows: int = 2048):
    """Row chunks are generated from their own RNG stream (seed, tensor, chunk), so the
    bytes do not depend on how many threads quantize them."""
    cols = ne[0]
    rows = 1
    for d in ne[1:]:
        rows *= d

    def chunk(r0: int) -> np.ndarray:
        n = min(chunk_rows, rows - r0)
        rng = np.random.default_rng([seed, idx, r0 // chunk_rows])
        scale = 1.0 if kind == "router" else sigma
        w = rng.standard_normal((n, cols), dtype=np.float32) * np.float32(scale)
        return quantize(w, qtype).reshape(-1)

    def run() -> np.ndarray:
        if kind == "norm":
            rng = np.random.default_rng([seed, idx])
            return (1.0 + sigma * rng.standard_normal(cols, dtype=np.float32)).astype(np.float32).view(np.uint8)
        starts = list(range(0, rows, chunk_rows))
        parts = list(_pool().map(chunk, starts)) if len(starts) > 1 else [chunk(0)]
        return np.concatenate(parts) if len(parts) > 1 else parts[0]

    return run


# ---- fast synthetic blocks: random block BYTES instead of quantized random floats -------------------------------------
# Any byte pattern is a valid quant block once its fp16 scale fields are sane.  Quants, sub-block scales and mins are
# uniform random bytes; d (and dmin = d * mean(q), which centres the weights) is set so that the dequantized weights have
# standard deviation ~sigma.  ~50x faster than quantizing floats: a 70B-shape file (39 GB) is written in about a minute,
# which is what lets bench.py run BASELINE config 4 at full depth.  Layout offsets: gguf block structs (zb_quant.cuh).
_FAST = {   # qtype: (block bytes, weights per block, std of the dequantized weight per unit d, [(offset, multiple of d)] fp16 fields)
    Q4_0: (18, 32, 4.61, [(0, 1.0)]),
    Q8_0: (34, 32, 73.9, [(0, 1.0)]),
    Q4_K: (144, 256, 258.0, [(0, 1.0), (2, 7.5)]),
    Q5_K: (176, 256, 527.0, [(0, 1.0), (2, 15.5)]),
    Q6_K: (210, 256, 1366.0, [(208, 1.0)]),
}


def random_blocks(qtype: int, n_blocks: int, rng: np.random.Generator, sigma: float) -> np.ndarray:
    bb, _, unit_std, fields = _FAST[qtype]
    n = n_blocks * bb
    raw = rng.integers(0, 2 ** 64 - 1, (n + 7) // 8, dtype=np.uint64, endpoint=True).view(np.uint8)[:n].reshape(n_blocks, bb)   # raw generator words
    d = (np.float32(sigma / unit_std) * (np.float32(0.75) + np.float32(0.5) * rng.random(n_blocks, dtype=np.float32))).astype(np.float32)
    for off, mult in fields:
        raw[:, off:off + 2] = (d * np.float32(mult)).astype(np.float16).view(np.uint8).reshape(n_blocks, 2)
    return raw.reshape(-1)


def _produce_fast(idx: int, qtype: int, ne: Tuple[int, ...], kind: str, seed: int, sigma: float, chunk_blocks: int = 1 << 18):
    if kind != "matmul" or qtype not in _FAST:
        return _produce(idx, qtype, ne, kind, seed, sigma)
    n_w = 1
    for d in ne:
        n_w *= d
    n_blocks = n_w // _FAST[qtype][1]

    def chunk(c0: int) -> np.ndarray:
        rng = np.random.Generator(np.random.SFC64([seed, idx, c0 // chunk_blocks]))
        return random_blocks(qtype, min(chunk_blocks, n_blocks - c0), rng, sigma)

    def run() -> np.ndarray:
        starts = list(range(0, n_blocks, chunk_blocks))
        parts = list(_pool().map(chunk, starts)) if len(starts) > 1 else [chunk(0)]
        return np.concatenate(parts) if len(parts) > 1 else parts[0]

    return run


def write_synthetic_gguf(path: str, spec: ModelSpec, seed: int = 1234, sigma: float = 0.02, fast: bool = False) -> None:
    """Seeded N(0, sigma) matmul weights, 1+N(0,sigma) norm gains, quantized once into native blocks.
    fast=True: matmul tensors are random block bytes of the same scale (see random_blocks) -- for the large bench shapes."""
    w = GGUFWriter(path, model_metadata(spec))
    for idx, (name, qt, ne, kind) in enumerate(tensor_plan(spec)):
        w.add(name, qt, ne, (_produce_fast if fast else _produce)(idx, qt, ne, kind, seed, sigma))
    w.write()


def model_weight_bytes(spec: ModelSpec, active_only: bool = True) -> int:
    """Bytes of weights read per decoded token (embedding row excluded, lm_head included)."""
    total = 0
    for name, qt, ne, kind in tensor_plan(spec):
        n = 1
        for d in ne:
            n *= d
        b = n // BLOCK_ELEMS[qt] * BLOCK_BYTES[qt]
        if name == "token_embd.weight":
            if spec.tied:
                total += b
            continue
        if "_exps" in name and active_only and spec.n_experts:
            b = b * spec.top_k // spec.n_experts
        total += b
    return total
