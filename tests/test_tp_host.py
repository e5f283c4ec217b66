"""Host-side tensor-parallel logic on CPU: the shard math of zb_tp_shard_host (row split for column-parallel
layers, block-aligned K split for row-parallel layers) plus the all-reduce pattern, exercised with a world_size-2
gloo group: every rank contracts its shard with the oracle and the summed / gathered result must equal the unsharded
oracle GEMV (inference/parallel/tensor_parallel.go:40-48,151-163)."""
import ctypes as C
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zerfoo_b200 import gguf as G


def _shard(qt, raw, rows, cols, r0, r1, c0, c1):
    from zerfoo_b200 import lib
    L = lib.load()
    rb = (c1 - c0) // G.BLOCK_ELEMS[qt] * G.BLOCK_BYTES[qt]
    out = np.zeros((r1 - r0) * rb, np.uint8)
    raw = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
    rc = L.zb_tp_shard_host(qt, raw.ctypes.data, rows, cols, r0, r1, c0, c1, out.ctypes.data)
    assert rc == 0
    return out


def _worker(rank, world, port, qt, rows, cols, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    raw = G.quantize(rng.standard_normal((rows, cols), dtype=np.float32) * np.float32(0.02), qt)
    x = rng.standard_normal(cols, dtype=np.float32)
    full = O.gemv_f64(qt, raw, rows, cols, x)
    # column-parallel (q/k/v, gate/up, lm_head): row shards, all-gather
    rl = rows // world
    mine = _shard(qt, raw, rows, cols, rank * rl, (rank + 1) * rl, 0, cols)
    part = torch.from_numpy(O.gemv_f64(qt, mine, rl, cols, x).astype(np.float64))
    parts = [torch.zeros_like(part) for _ in range(world)]
    dist.all_gather(parts, part)
    ok1 = np.allclose(torch.cat(parts).numpy(), full, rtol=1e-6, atol=1e-7)
    # row-parallel (o_proj, down_proj): block-aligned K shards, all-reduce(sum)
    cl = cols // world
    mine = _shard(qt, raw, rows, cols, 0, rows, rank * cl, (rank + 1) * cl)
    part = torch.from_numpy(O.gemv_f64(qt, mine, rows, cl, x[rank * cl:(rank + 1) * cl]).astype(np.float64))
    dist.all_reduce(part)
    ok2 = np.allclose(part.numpy(), full, rtol=1e-5, atol=1e-6)
    if rank == 0:
        q.put((ok1, ok2))
    dist.destroy_process_group()


@pytest.mark.parametrize("qt,rows,cols", [(G.Q4_K, 64, 1024), (G.Q6_K, 32, 512), (G.Q4_0, 48, 128), (G.Q8_0, 16, 256), (G.Q5_K, 8, 2048)])
def test_tp_shards_recombine_world2_gloo(qt, rows, cols):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, qt, rows, cols, 7, q)) for r in range(2)]
    for p in procs: p.start()
    ok = q.get(timeout=120)
    for p in procs: p.join(timeout=60)
    assert ok == (True, True)


def test_tp_shard_rejects_unaligned_columns():
    from zerfoo_b200 import lib
    L = lib.load()
    raw = np.zeros(4 * 144, np.uint8); out = np.zeros(4 * 144, np.uint8)
    assert L.zb_tp_shard_host(G.Q4_K, raw.ctypes.data, 4, 256, 0, 4, 0, 128, out.ctypes.data) != 0   # 128 is not a Q4_K block boundary
    assert L.zb_tp_shard_host(G.Q4_K, raw.ctypes.data, 4, 256, 0, 2, 0, 256, out.ctypes.data) == 0
