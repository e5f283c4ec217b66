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
