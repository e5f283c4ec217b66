"""Host-side mirror of the reference's decode API over the C-ABI engine:

  * ``load_file``            inference.LoadFile (inference/load_gguf.go:17)
  * ``Generator.generate``   InferenceSession.Generate, temperature 0 (generate/session.go:84-268)
  * ``Generator.decode_step``  runDecodeStep + tryGPUArgmax (generate/decode_step.go:26-67)

Everything below is a thin ctypes call into ``zb_engine_*`` (include/zb200.h);
there is no CPU path -- without the CUDA library or a device, construction fails.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import lib as _lib


class EngineOpts(C.Structure):
    _fields_ = [("device", C.c_int), ("max_seq", C.c_int), ("use_graph", C.c_int), ("tp_rank", C.c_int), ("tp_size", C.c_int),
                ("batch", C.c_int), ("flags", C.c_int), ("reserved", C.c_int * 7)]


class ModelInfo(C.Structure):
    _fields_ = [("vocab", C.c_int), ("hidden", C.c_int), ("layers", C.c_int), ("n_q", C.c_int), ("n_kv", C.c_int), ("head_dim", C.c_int),
                ("ffn", C.c_int), ("max_seq", C.c_int), ("n_experts", C.c_int), ("top_k", C.c_int), ("tp_rank", C.c_int), ("tp_size", C.c_int),
                ("weight_bytes_per_token", C.c_int64), ("kv_bytes_per_pos", C.c_int64), ("launches_per_step", C.c_int), ("arch", C.c_char * 32)]


class GemvProfile(C.Structure):
    _fields_ = [("qtype", C.c_int), ("launches", C.c_int64), ("bytes", C.c_double), ("ms", C.c_double)]


class EngineError(RuntimeError):
    pass


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = _lib.load().zb_last_error()
        raise EngineError(f"{what} failed (rc {rc}): {msg.decode() if msg else ''}")


class Generator:
    """One loaded model + its KV cache + its captured decode-step graph."""

    def __init__(self, path: str, device: int = 0, max_seq: int = 0, use_graph: bool = True, tp_rank: int = 0, tp_size: int = 1,
                 batch: int = 1, nccl_id: Optional[bytes] = None, mega: Optional[bool] = None, kv_f16: bool = False):
        L = _lib.load()
        # mega=True asks for the persistent whole-token kernel (ZB_ENGINE_MEGA) instead of the CUDA-graph step of per-matrix launches
        opts = EngineOpts(device=device, max_seq=max_seq, use_graph=int(use_graph), tp_rank=tp_rank, tp_size=tp_size, batch=batch,
                          flags=(0 if mega is None else (2 if mega else 1)) | (4 if kv_f16 else 0))   # mega None: library default (graph step unless ZB_MEGA=1)
        self.batch = batch
        h = C.c_void_p()
        if tp_size > 1:
            if nccl_id is None or len(nccl_id) != 128:
                raise EngineError("tensor parallel needs the 128-byte NCCL unique id (tp_unique_id() on rank 0, broadcast to all ranks)")
            buf = (C.c_char * 128).from_buffer_copy(nccl_id)
            _check(L.zb_engine_create_tp(path.encode(), C.byref(opts), buf, C.byref(h)), "zb_engine_create_tp")
        else:
            _check(L.zb_engine_create(path.encode(), C.byref(opts), C.byref(h)), "zb_engine_create")
        self._h = h
        self._L = L
        self.tp_fused = False
        self.info = ModelInfo()
        _check(L.zb_engine_info(self._h, C.byref(self.info)), "zb_engine_info")

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.zb_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def refresh_info(self) -> ModelInfo:
        _check(self._L.zb_engine_info(self._h, C.byref(self.info)), "zb_engine_info")
        return self.info

    # -- cache / position ------------------------------------------------------
    def reset(self) -> None:
        _check(self._L.zb_engine_reset(self._h), "zb_engine_reset")

    @property
    def position(self) -> int:
        return self._L.zb_engine_position(self._h)

    @property
    def stream(self) -> int:
        return self._L.zb_engine_stream(self._h) or 0

    # -- forward ---------------------------------------------------------------
    def prefill(self, tokens: Sequence[int]) -> int:
        arr = (C.c_int32 * len(tokens))(*tokens)
        first = C.c_int32()
        _check(self._L.zb_engine_prefill(self._h, arr, len(tokens), C.byref(first)), "zb_engine_prefill")
        return first.value

    def prefill_chunked(self, tokens: Sequence[int]) -> Tuple[int, float]:
        """Prompt prefill in 256-token chunks through the tcgen05 dequant-GEMMs and the causal chunk attention
        (bf16 tiles: batched-path tolerance, not the bit-exact token-by-token `prefill`).  Returns (first token, device ms)."""
        arr = (C.c_int32 * len(tokens))(*tokens)
        first = C.c_int32()
        ms = C.c_float()
        _check(self._L.zb_engine_prefill_chunked(self._h, arr, len(tokens), C.byref(first), C.byref(ms)), "zb_engine_prefill_chunked")
        return first.value, ms.value

    def decode_step(self, token: int) -> int:
        nxt = C.c_int32()
        _check(self._L.zb_engine_decode_step(self._h, int(token), C.byref(nxt)), "zb_engine_decode_step")
        return nxt.value

    def decode_n(self, first_token: int, n: int) -> Tuple[List[int], float]:
        out = (C.c_int32 * n)()
        ms = C.c_float()
        _check(self._L.zb_engine_decode_n(self._h, int(first_token), n, out, C.byref(ms)), "zb_engine_decode_n")
        return list(out), ms.value

    def generate(self, prompt: Sequence[int], n_new: int) -> List[int]:
        arr = (C.c_int32 * len(prompt))(*prompt)
        out = (C.c_int32 * n_new)()
        _check(self._L.zb_engine_generate(self._h, arr, len(prompt), n_new, out), "zb_engine_generate")
        return list(out)

    def profile_gemv(self, steps: int):
        """[(qtype, launches, algorithmic bytes, summed device ms)] over `steps` eager decode steps."""
        arr = (GemvProfile * 8)()
        n = C.c_int()
        _check(self._L.zb_engine_profile_gemv(self._h, steps, arr, 8, C.byref(n)), "zb_engine_profile_gemv")
        return [(arr[i].qtype, arr[i].launches, arr[i].bytes, arr[i].ms) for i in range(n.value)]

    def profile_gemv_graph(self, qtype: int, reps: int = 4):
        """(launches, algorithmic bytes, device ms) of every GEMV of one block format replayed `reps` times as a PDL-chained graph."""
        r = GemvProfile()
        _check(self._L.zb_engine_profile_gemv_graph(self._h, qtype, reps, C.byref(r)), "zb_engine_profile_gemv_graph")
        return r.launches, r.bytes, r.ms

    def tp_allreduce_us(self, count: int, reps: int = 4) -> float:
        """Microseconds for `count` back-to-back all-reduces of one hidden-size vector on the engine's communicator
        (collective: every rank calls it).  Sets self.tp_fused (whether the decode step uses the fused exchange instead)."""
        us, fused = C.c_float(), C.c_int()
        _check(self._L.zb_engine_tp_allreduce_us(self._h, count, reps, C.byref(us), C.byref(fused)), "zb_engine_tp_allreduce_us")
        self.tp_fused = fused.value == 1
        self.tp_exchange = {0: "nccl all-reduce", 1: "fused into the GEMV epilogue / prologue (peer-memory LL slots)",
                            2: "one-shot push all-reduce over peer memory (LL pairs)"}.get(fused.value, "?")
        return us.value

    # -- batched decode over the paged KV cache (opts.batch > 1) ---------------
    def batch_reset(self) -> None:
        _check(self._L.zb_engine_batch_reset(self._h), "zb_engine_batch_reset")

    def batch_step(self, tokens: Sequence[int]) -> List[int]:
        assert len(tokens) == self.batch
        arr = (C.c_int32 * self.batch)(*tokens)
        out = (C.c_int32 * self.batch)()
        _check(self._L.zb_engine_batch_step(self._h, arr, out), "zb_engine_batch_step")
        return list(out)

    def batch_decode_n(self, first_tokens: Sequence[int], n: int):
        arr = (C.c_int32 * self.batch)(*first_tokens)
        out = (C.c_int32 * (n * self.batch))()
        ms = C.c_float()
        _check(self._L.zb_engine_batch_decode_n(self._h, arr, n, out, C.byref(ms)), "zb_engine_batch_decode_n")
        return np.array(out, dtype=np.int32).reshape(n, self.batch), ms.value

    def batch_logits(self) -> np.ndarray:
        out = np.empty((self.batch, self.info.vocab), dtype=np.float32)
        _check(self._L.zb_engine_batch_logits(self._h, out.ctypes.data_as(C.c_void_p)), "zb_engine_batch_logits")
        return out

    # -- taps -------------------------------------------------------------------
    def logits(self) -> np.ndarray:
        out = np.empty(self.info.vocab, dtype=np.float32)
        _check(self._L.zb_engine_logits(self._h, out.ctypes.data_as(C.c_void_p)), "zb_engine_logits")
        return out

    def hidden(self) -> np.ndarray:
        out = np.empty(self.info.hidden, dtype=np.float32)
        _check(self._L.zb_engine_hidden(self._h, out.ctypes.data_as(C.c_void_p)), "zb_engine_hidden")
        return out

    def kv(self, layer: int, n: int) -> Tuple[np.ndarray, np.ndarray]:
        d = self.info.n_kv * self.info.head_dim
        k = np.empty((n, d), dtype=np.float32)
        v = np.empty((n, d), dtype=np.float32)
        _check(self._L.zb_engine_kv(self._h, layer, n, k.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p)), "zb_engine_kv")
        return k, v


def tp_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0); broadcast the bytes to the other ranks."""
    buf = (C.c_char * 128)()
    _check(_lib.load().zb_tp_unique_id(buf), "zb_tp_unique_id")
    return bytes(buf.raw)


def load_file_tp(path: str, **kw) -> "Generator":
    """inference.LoadFile for one rank of a tensor-parallel group: reads RANK / WORLD_SIZE / LOCAL_RANK from the
    environment (torchrun), shares the NCCL id over torch.distributed, shards the weights at load."""
    import torch
    import torch.distributed as dist
    rank, world, local = int(__import__("os").environ["RANK"]), int(__import__("os").environ["WORLD_SIZE"]), int(__import__("os").environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    box = [tp_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return Generator(path, device=local, tp_rank=rank, tp_size=world, nccl_id=box[0], **kw)


def load_file(path: str, **kw) -> Generator:
    return Generator(path, **kw)
