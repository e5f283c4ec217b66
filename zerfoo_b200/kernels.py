"""Host-side mirror of the reference's Go kernel wrappers
(internal/cuda/kernels/*_purego.go): same names, same argument meaning, same
error behaviour (non-zero return -> "<op> kernel failed (cuda error N)").

Arguments are torch CUDA tensors (device memory + streams are PyTorch's job
here, exactly as ``cuda.Malloc``/arena buffers are the Go host's); the data
path is the C-ABI call into libkernels.so, nothing is computed in Python.
"""
from __future__ import annotations

import struct
from typing import Optional, Tuple

import numpy as np
import torch

from . import lib as _lib
from . import gguf as G


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("zerfoo_b200.kernels: tensors must live on a CUDA device (no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError("zerfoo_b200.kernels: tensors must be contiguous")
    return t.data_ptr()


def _bits(f: float) -> int:
    return struct.unpack("<I", struct.pack("<f", f))[0]


def _dev_bytes(raw: np.ndarray, device="cuda") -> torch.Tensor:
    raw = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
    pad = (-raw.size) % 16
    if pad:
        raw = np.concatenate([raw, np.zeros(pad, np.uint8)])
    return torch.from_numpy(raw.copy()).to(device)


# ---- weight upload helpers (compute.WeightUploader.UploadWeights) -----------

def upload_q4_0(raw_gguf: np.ndarray) -> Tuple[torch.Tensor, int]:
    """GGUF Q4_0 blocks -> separated device layout; returns (buffer, data_offset)."""
    sep, off = G.q4_0_to_separated(raw_gguf)
    return _dev_bytes(sep), off


def upload_q8_0(raw_gguf: np.ndarray) -> torch.Tensor:
    return _dev_bytes(G.q8_0_to_zerfoo36(raw_gguf))


def upload_raw(raw: np.ndarray) -> torch.Tensor:
    return _dev_bytes(raw)


# ---- dequant-matmul --------------------------------------------------------

def GemmQ4F32(A_q4: torch.Tensor, B: torch.Tensor, C: torch.Tensor, M: int, K: int, N: int, dataOffset: int) -> None:
    """gemm_q4_purego.go:15-29 -> gemm_q4_f32."""
    _lib.check(_lib.load().gemm_q4_f32(_p(A_q4), _p(B), _p(C), M, K, N, dataOffset, _stream()), "gemm_q4_f32")


def GemmQ8F32(A_q8: torch.Tensor, B: torch.Tensor, C: torch.Tensor, M: int, K: int, N: int) -> None:
    _lib.check(_lib.load().gemm_q8_f32(_p(A_q8), _p(B), _p(C), M, K, N, _stream()), "gemm_q8_f32")


def GemvQ4KF32(W: torch.Tensor, x: torch.Tensor, y: torch.Tensor, M: int, K: int) -> None:
    """gemv_q4k_purego.go -> gemv_q4k_f32."""
    _lib.check(_lib.load().gemv_q4k_f32(_p(W), _p(x), _p(y), M, K, _stream()), "gemv_q4k_f32")


def GemvQ5KF32(W: torch.Tensor, x: torch.Tensor, y: torch.Tensor, M: int, K: int) -> None:
    _lib.check(_lib.load().gemv_q5k_f32(_p(W), _p(x), _p(y), M, K, _stream()), "gemv_q5k_f32")


def GemvQ6KF32(W: torch.Tensor, x: torch.Tensor, y: torch.Tensor, M: int, K: int) -> None:
    _lib.check(_lib.load().gemv_q6k_f32(_p(W), _p(x), _p(y), M, K, _stream()), "gemv_q6k_f32")


def GemvQ4_0RawF32(W: torch.Tensor, x: torch.Tensor, y: torch.Tensor, M: int, K: int) -> None:
    _lib.check(_lib.load().zb_gemv_q4_0_f32(_p(W), _p(x), _p(y), M, K, _stream()), "zb_gemv_q4_0_f32")


def GemvQ8_0RawF32(W: torch.Tensor, x: torch.Tensor, y: torch.Tensor, M: int, K: int) -> None:
    _lib.check(_lib.load().zb_gemv_q8_0_f32(_p(W), _p(x), _p(y), M, K, _stream()), "zb_gemv_q8_0_f32")


def SgemvM1(y: torch.Tensor, A: torch.Tensor, x: torch.Tensor, M: int, N: int) -> None:
    _lib.check(_lib.load().launch_sgemv_m1(_p(y), _p(A), _p(x), M, N, _stream()), "launch_sgemv_m1")


def DequantQ4KF32(src: torch.Tensor, dst: torch.Tensor, rows: int, K: int) -> None:
    _lib.check(_lib.load().dequant_q4k_f32(_p(src), _p(dst), rows, K, _stream()), "dequant_q4k_f32")


def DequantF32(qtype: int, src: torch.Tensor, dst: torch.Tensor, n: int) -> None:
    _lib.check(_lib.load().zb_dequant_f32(qtype, _p(src), _p(dst), n, _stream()), "zb_dequant_f32")


def gemv(qtype: int, raw_gguf: np.ndarray, rows: int, k: int, x: torch.Tensor) -> torch.Tensor:
    """Convenience for tests: upload raw GGUF blocks in the engine's device layout and run the GEMV."""
    y = torch.empty(rows, dtype=torch.float32, device=x.device)
    if qtype == G.Q4_0:
        buf, off = upload_q4_0(raw_gguf)
        GemmQ4F32(buf, x, y, rows, k, 1, off)
    elif qtype == G.Q8_0:
        GemmQ8F32(upload_q8_0(raw_gguf), x, y, rows, k, 1)
    elif qtype == G.Q4_K:
        GemvQ4KF32(upload_raw(raw_gguf), x, y, rows, k)
    elif qtype == G.Q5_K:
        GemvQ5KF32(upload_raw(raw_gguf), x, y, rows, k)
    elif qtype == G.Q6_K:
        GemvQ6KF32(upload_raw(raw_gguf), x, y, rows, k)
    elif qtype == G.F32:
        SgemvM1(y, upload_raw(raw_gguf).view(torch.float32) if False else torch.from_numpy(np.ascontiguousarray(raw_gguf).view(np.float32).reshape(rows, k).copy()).to(x.device), x, rows, k)
    else:
        raise ValueError(f"unsupported qtype {qtype}")
    return y


# ---- fused epilogues ---------------------------------------------------------

def FusedAddRMSNormF32(inp, residual, weight, normed_out, sum_out, eps: float, rows: int, D: int) -> None:
    _lib.check(_lib.load().fused_add_rmsnorm_f32(_p(inp), _p(residual), _p(weight), _p(normed_out), _p(sum_out), _bits(eps), rows, D, _stream()),
               "fused_add_rmsnorm_f32")


def FusedNormAddF32(inp, weight, residual, output, eps: float, rows: int, D: int) -> None:
    _lib.check(_lib.load().fused_norm_add_f32(_p(inp), _p(weight), _p(residual), _p(output), _bits(eps), rows, D, _stream()), "fused_norm_add_f32")


def RMSNorm(inp, weight, output, scales, eps: float, rows: int, D: int) -> None:
    _lib.check(_lib.load().launch_rmsnorm(_p(inp), _p(weight), _p(output), _p(scales), _bits(eps), rows, D, _stream()), "launch_rmsnorm")


def FusedSwiGLUF32(w1, w3, output, n: int) -> None:
    _lib.check(_lib.load().fused_swiglu_f32(_p(w1), _p(w3), _p(output), n, _stream()), "fused_swiglu_f32")


def FusedQKNormRoPEF32(inp, weightQ, weightK, cosA, sinA, output, eps: float, totalHeads: int, headDim: int, numQHeads: int, halfRotary: int) -> None:
    _lib.check(_lib.load().fused_qk_norm_rope_f32(_p(inp), _p(weightQ), _p(weightK), _p(cosA), _p(sinA), _p(output), _bits(eps), totalHeads, headDim,
                                                  numQHeads, halfRotary, _stream()), "fused_qk_norm_rope_f32")


def FusedRoPEF32(inp, cosA, sinA, output, batch: int, seqLen: int, headDim: int, halfRotary: int, cosStride: int) -> None:
    _lib.check(_lib.load().fused_rope_f32(_p(inp), _p(cosA), _p(sinA), _p(output), batch, seqLen, headDim, halfRotary, cosStride, _stream()), "fused_rope_f32")


def RoPESelect(cosTable, sinTable, cosOut, sinOut, counter, halfRotary: int) -> None:
    _lib.check(_lib.load().launch_rope_select(_p(cosTable), _p(sinTable), _p(cosOut), _p(sinOut), _p(counter), halfRotary, _stream()), "launch_rope_select")


def OffsetMemcpy(dst, src, counter, dim: int, maxSeqLen: int) -> None:
    _lib.check(_lib.load().launch_offset_memcpy(_p(dst), _p(src), _p(counter), dim, maxSeqLen, _stream()), "launch_offset_memcpy")


def OffsetMemcpyFP16(dst, src, counter, dim: int, maxSeqLen: int) -> None:
    _lib.check(_lib.load().launch_offset_memcpy_fp16(_p(dst), _p(src), _p(counter), dim, maxSeqLen, _stream()), "launch_offset_memcpy_fp16")


def IncrementCounter(counter, delta: int) -> None:
    _lib.check(_lib.load().launch_increment_counter(_p(counter), delta, _stream()), "launch_increment_counter")


def ResetCounter(counter, value: int) -> None:
    _lib.check(_lib.load().launch_reset_counter(_p(counter), value, _stream()), "launch_reset_counter")


def Argmax(inp, result, scratch, n: int) -> None:
    _lib.check(_lib.load().launch_argmax(_p(inp), _p(result), _p(scratch), n, _stream()), "launch_argmax")


def ScaledSoftmaxF32(inp, output, outer: int, inner: int, axisSize: int, scale: float) -> None:
    _lib.check(_lib.load().scaled_softmax_f32(_p(inp), _p(output), outer, inner, axisSize, _bits(scale), _stream()), "scaled_softmax_f32")


def GatherI32(table, indices, output, N: int, D: int, V: int) -> None:
    _lib.check(_lib.load().launch_gather_i32(_p(table), _p(indices), _p(output), N, D, V, _stream()), "launch_gather_i32")


def Gather(table, indices, output, N: int, D: int, V: int) -> None:
    _lib.check(_lib.load().launch_gather(_p(table), _p(indices), _p(output), N, D, V, _stream()), "launch_gather")


# ---- attention ---------------------------------------------------------------

def FlashDecodeSplitKVF32(Q, K, V, O, partialO, partialLSE, numBH: int, maxKVLen: int, headDim: int, kvLen: int, kvLenPtr, numQHeads: int,
                          numKVHeads: int, chunkSize: int) -> None:
    _lib.check(_lib.load().flash_decode_splitkv_f32(_p(Q), _p(K), _p(V), _p(O), _p(partialO), _p(partialLSE), numBH, maxKVLen, headDim, kvLen,
                                                    _p(kvLenPtr), numQHeads, numKVHeads, chunkSize, _stream()), "flash_decode_splitkv_f32")


def FlashAttentionDecodeF32(Q, K, V, O, numBH: int, maxKVLen: int, headDim: int, kvLen: int, kvLenPtr, numQHeads: int, numKVHeads: int) -> None:
    _lib.check(_lib.load().flash_attention_decode_f32(_p(Q), _p(K), _p(V), _p(O), numBH, maxKVLen, headDim, kvLen, _p(kvLenPtr), numQHeads, numKVHeads,
                                                      _stream()), "flash_attention_decode_f32")


def FlashAttentionForwardF32(Q, K, V, O, batch: int, heads: int, seqLen: int, headDim: int, causal: bool) -> None:
    _lib.check(_lib.load().flash_attention_forward_f32(_p(Q), _p(K), _p(V), _p(O), batch, heads, seqLen, headDim, int(causal), _stream()),
               "flash_attention_forward_f32")


# ---- boundary-only elementwise (a few, for the symbol/behaviour tests) -------

def Add(a, b, c, n: int) -> None:
    _lib.check(_lib.load().launch_add(_p(a), _p(b), _p(c), n, _stream()), "launch_add")


def MulScalar(a, scalar: float, c, n: int) -> None:
    _lib.check(_lib.load().launch_mul_scalar(_p(a), _bits(scalar), _p(c), n, _stream()), "launch_mul_scalar")


def Tanh(a, c, n: int) -> None:
    _lib.check(_lib.load().launch_tanh(_p(a), _p(c), n, _stream()), "launch_tanh")


def Transpose2D(inp, output, rows: int, cols: int) -> None:
    _lib.check(_lib.load().launch_transpose_2d(_p(inp), _p(output), rows, cols, _stream()), "launch_transpose_2d")


# ---- B200 streamed path (include/zb200.h: zb_gemv_stream_f32, zb_decode_attn_f32) -------------
import ctypes as _C


class StreamWeightC(_C.Structure):
    _fields_ = [("main", _C.c_void_p), ("aux", _C.c_void_p), ("qtype", _C.c_int), ("rows", _C.c_int), ("cols", _C.c_int),
                ("expert_sel", _C.c_void_p), ("n_sel", _C.c_int), ("y_slot_stride", _C.c_int),
                ("expert_main_stride", _C.c_int64), ("expert_aux_stride", _C.c_int64), ("epilogue", _C.c_int),
                ("n_peers", _C.c_int), ("site", _C.c_int), ("sites_per_step", _C.c_int), ("peer_out", _C.c_void_p * 8),
                ("peer_flag", _C.c_void_p * 8), ("epoch_base", _C.c_void_p), ("ticket", _C.c_void_p)]


class PrologueC(_C.Structure):
    _fields_ = [("a", _C.c_void_p), ("r", _C.c_void_p), ("w1", _C.c_void_p), ("w2", _C.c_void_p), ("sum_out", _C.c_void_p),
                ("mix_w", _C.c_void_p), ("mix_n", _C.c_int), ("mix_stride", _C.c_int), ("eps", _C.c_float), ("swiglu", _C.c_int),
                ("a_slot_stride", _C.c_int), ("wait_flags", _C.c_void_p), ("wait_epoch_base", _C.c_void_p), ("n_wait", _C.c_int),
                ("wait_site", _C.c_int), ("wait_sites_per_step", _C.c_int), ("a_replicas", _C.c_int), ("a_replica_stride", _C.c_int)]


class AttnArgsC(_C.Structure):
    _fields_ = [("qkv", _C.c_void_p), ("q_norm", _C.c_void_p), ("k_norm", _C.c_void_p), ("cos_tbl", _C.c_void_p), ("sin_tbl", _C.c_void_p),
                ("pos", _C.c_void_p), ("k_cache", _C.c_void_p), ("v_cache", _C.c_void_p), ("out", _C.c_void_p), ("part_o", _C.c_void_p),
                ("part_ml", _C.c_void_p), ("ticket", _C.c_void_p), ("eps", _C.c_float), ("head_dim", _C.c_int), ("n_q", _C.c_int),
                ("n_kv", _C.c_int), ("max_seq", _C.c_int), ("chunk", _C.c_int), ("max_splits", _C.c_int),
                ("batch", _C.c_int), ("qkv_stride", _C.c_int), ("out_stride", _C.c_int), ("block_table", _C.c_void_p),
                ("max_blocks", _C.c_int), ("page", _C.c_int), ("warps", _C.c_int), ("window", _C.c_int), ("window_on", _C.c_void_p), ("kv_f16", _C.c_int)]


class StreamWeight:
    """A quantized matrix in the stream layout on the device (UploadWeights for the B200 engine)."""

    def __init__(self, qtype: int, raw_gguf: np.ndarray, rows: int, cols: int, experts: int = 1):
        L = _lib.load()
        mb, ab = _C.c_int64(), _C.c_int64()
        _lib.check(L.zb_stream_layout(qtype, rows, cols, _C.byref(mb), _C.byref(ab)), "zb_stream_layout")
        raw = np.ascontiguousarray(raw_gguf).view(np.uint8).reshape(-1)
        hm = np.zeros(mb.value + 64, np.uint8)
        ha = np.zeros(ab.value + 64, np.uint8)
        _lib.check(L.zb_stream_repack_host(qtype, raw.ctypes.data, rows, cols, hm.ctypes.data, ha.ctypes.data if ab.value else None),
                   "zb_stream_repack_host")
        self.main = torch.from_numpy(hm).cuda()
        self.aux = torch.from_numpy(ha).cuda() if ab.value else None
        self.qtype, self.cols, self.experts = qtype, cols, experts
        self.rows = rows // experts
        self.main_stride = (mb.value // experts) if experts > 1 else 0
        self.aux_stride = ((self.rows * cols // 32 * 2) if qtype in (G.Q4_0, G.Q8_0) else (self.rows * cols // 256 * 2 if qtype == G.Q6_K else 0)) \
            if experts > 1 else 0


def gemv_stream(w: StreamWeight, a: torch.Tensor, *, r=None, w1=None, w2=None, sum_out=None, eps: float = 1e-5, swiglu: bool = False,
                mix_w=None, mix_n: int = 0, mix_stride: int = 0, sel=None, a_slot_stride: int = 0, pdl: bool = False, swiglu_pairs: bool = False,
                y: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = deq(W) . prologue(a, ...) through zb_gemv_stream_f32."""
    L = _lib.load()
    nsel = int(sel.numel()) if sel is not None else 0
    if y is None:
        y = torch.empty(max(nsel, 1) * (w.rows // 2 if swiglu_pairs else w.rows), dtype=torch.float32, device=a.device)
    sw = StreamWeightC(main=_p(w.main), aux=_p(w.aux), qtype=w.qtype, rows=w.rows, cols=w.cols, expert_sel=_p(sel), n_sel=nsel,
                       y_slot_stride=(w.rows // 2 if swiglu_pairs else w.rows), expert_main_stride=w.main_stride,
                       expert_aux_stride=w.aux_stride, epilogue=1 if swiglu_pairs else 0)
    pr = PrologueC(a=_p(a), r=_p(r), w1=_p(w1), w2=_p(w2), sum_out=_p(sum_out), mix_w=_p(mix_w), mix_n=mix_n, mix_stride=mix_stride,
                   eps=eps, swiglu=int(swiglu), a_slot_stride=a_slot_stride)
    _lib.check(L.zb_gemv_stream_f32(_C.byref(sw), _C.byref(pr), _p(y), 1 if pdl else 0, _stream()), "zb_gemv_stream_f32")
    return y


class MmaWeightC(_C.Structure):
    _fields_ = [("data", _C.c_void_p), ("qtype", _C.c_int), ("rows", _C.c_int), ("cols", _C.c_int), ("epilogue", _C.c_int),
                ("expert_sel", _C.c_void_p), ("n_sel", _C.c_int), ("y_slot_stride", _C.c_int), ("expert_stride", _C.c_int64)]


class MmaWeight:
    """A quantized matrix as 16-row x 256-weight block-tiles for the tensor-core GEMV (zb_mma_repack_host)."""

    def __init__(self, qtype: int, raw_gguf: np.ndarray, rows: int, cols: int, experts: int = 1):
        L = _lib.load()
        wb, sb = _C.c_int64(), _C.c_int64()
        _lib.check(L.zb_mma_layout(qtype, rows, cols, _C.byref(wb), _C.byref(sb)), "zb_mma_layout")
        self.experts, self.expert_stride = experts, 0
        if experts > 1:   # a stack of `experts` matrices of rows/experts rows each: tiles must not straddle experts
            assert (rows // experts) % 16 == 0
            self.expert_stride = wb.value // experts
        raw = np.ascontiguousarray(raw_gguf).view(np.uint8).reshape(-1)
        hw = np.zeros(wb.value, np.uint8)
        _lib.check(L.zb_mma_repack_host(qtype, raw.ctypes.data, rows, cols, hw.ctypes.data), "zb_mma_repack_host")
        self.data = torch.from_numpy(hw).cuda()
        self.scratch = torch.zeros(sb.value * 4, dtype=torch.uint8, device="cuda")
        self.qtype, self.rows, self.cols = qtype, rows // experts, cols


def gemv_mma(w: MmaWeight, a: torch.Tensor, *, r=None, w1=None, w2=None, sum_out=None, eps: float = 1e-5, swiglu: bool = False,
             pdl: bool = False, swiglu_pairs: bool = False, y: Optional[torch.Tensor] = None, a_replicas: int = 0,
             a_replica_stride: int = 0, sel=None, a_slot_stride: int = 0) -> torch.Tensor:
    """y = deq(W) . prologue(a, ...) through zb_gemv_mma_f32 (tensor-core batch-1 GEMV)."""
    L = _lib.load()
    nsel = int(sel.numel()) if sel is not None else 0
    nout = w.rows // 2 if swiglu_pairs else w.rows
    if y is None:
        y = torch.empty(max(nsel, 1) * nout, dtype=torch.float32, device=a.device)
    mw = MmaWeightC(data=_p(w.data), qtype=w.qtype, rows=w.rows, cols=w.cols, epilogue=1 if swiglu_pairs else 0,
                    expert_sel=_p(sel), n_sel=nsel, y_slot_stride=nout, expert_stride=w.expert_stride)
    pr = PrologueC(a=_p(a), r=_p(r), w1=_p(w1), w2=_p(w2), sum_out=_p(sum_out), eps=eps, swiglu=int(swiglu),
                   a_replicas=a_replicas, a_replica_stride=a_replica_stride, a_slot_stride=a_slot_stride)
    _lib.check(L.zb_gemv_mma_f32(_C.byref(mw), _C.byref(pr), _p(y), _p(w.scratch), 1 if pdl else 0, _stream()), "zb_gemv_mma_f32")
    return y


def decode_attn(qkv, q_norm, k_norm, cos_tbl, sin_tbl, pos, k_cache, v_cache, out, part_o, part_ml, ticket, eps: float, head_dim: int,
                n_q: int, n_kv: int, max_seq: int, chunk: int, max_splits: int, pdl: bool = False, batch: int = 0,
                qkv_stride: int = 0, out_stride: int = 0, block_table=None, max_blocks: int = 0, page: int = 16) -> None:
    L = _lib.load()
    a = AttnArgsC(qkv=_p(qkv), q_norm=_p(q_norm), k_norm=_p(k_norm), cos_tbl=_p(cos_tbl), sin_tbl=_p(sin_tbl), pos=_p(pos),
                  k_cache=_p(k_cache), v_cache=_p(v_cache), out=_p(out), part_o=_p(part_o), part_ml=_p(part_ml), ticket=_p(ticket),
                  eps=eps, head_dim=head_dim, n_q=n_q, n_kv=n_kv, max_seq=max_seq, chunk=chunk, max_splits=max_splits, batch=batch,
                  qkv_stride=qkv_stride, out_stride=out_stride, block_table=_p(block_table), max_blocks=max_blocks, page=page)
    _lib.check(L.zb_decode_attn_f32(_C.byref(a), 1 if pdl else 0, _stream()), "zb_decode_attn_f32")


def gemm_tc(w: StreamWeight, x: torch.Tensor, split_x: Optional[bool] = None) -> torch.Tensor:
    """Y[T, rows] = X[T, cols] . deq(W)^T on tcgen05 (zb_gemm_tc_prep_x + zb_gemm_tc_f32).  x: f32 [T, cols] on the device."""
    L = _lib.load()
    T, K = x.shape
    if split_x is None:
        split_x = T <= 64
    tp = (T + 15) // 16 * 16
    xhi = torch.zeros(tp, K, dtype=torch.bfloat16, device=x.device)
    xlo = torch.zeros(tp, K, dtype=torch.bfloat16, device=x.device) if split_x else None
    y = torch.empty(T, w.rows, dtype=torch.float32, device=x.device)
    _lib.check(L.zb_gemm_tc_prep_x(w.qtype, _p(x), T, K, K, _p(xhi), _p(xlo), tp, _stream()), "zb_gemm_tc_prep_x")
    sw = StreamWeightC(main=_p(w.main), aux=_p(w.aux), qtype=w.qtype, rows=w.rows, cols=w.cols)
    _lib.check(L.zb_gemm_tc_f32(_C.byref(sw), _p(xhi), _p(xlo), T, tp, _p(y), w.rows, _stream()), "zb_gemm_tc_f32")
    return y


def prefill_attn(qkv, q_norm, k_norm, cos_tbl, sin_tbl, p0: int, k_cache, v_cache, eps: float, head_dim: int, n_q: int, n_kv: int,
                 f32: bool = False, window: int = 0):
    """zb_prefill_attn_f32 on host arrays: QK-norm + RoPE + KV append of a prompt chunk at positions p0.., then causal attention
    over the cache.  qkv [T, (n_q+2n_kv)*hd]; caches [n_kv, max_seq, hd].  Returns (out [T, n_q*hd], k_cache, v_cache)."""
    import numpy as np
    L = _lib.load()
    dev = torch.device("cuda")
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    dq, dwq, dwk, dc, ds, dk, dv = t(qkv), t(q_norm), t(k_norm), t(cos_tbl), t(sin_tbl), t(k_cache), t(v_cache)
    T, ld = dq.shape
    max_seq = dk.shape[1]
    qrot = torch.empty(T, n_q * head_dim, dtype=torch.float32, device=dev)
    out = torch.empty(T, n_q * head_dim, dtype=torch.float32, device=dev)
    _lib.check(L.zb_prefill_attn_f32(_p(dq), ld, _p(dwq), _p(dwk), _p(dc), _p(ds), p0, T, _p(qrot), _p(dk), _p(dv), _p(out),
                                     _C.c_float(eps), head_dim, n_q, n_kv, max_seq, window, 1 if f32 else 0, _stream()), "zb_prefill_attn_f32")
    torch.cuda.synchronize()
    return out.cpu().numpy(), dk.cpu().numpy(), dv.cpu().numpy()
