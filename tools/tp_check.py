#!/usr/bin/env python
"""Tensor-parallel parity + timing, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py
Every rank shards the same synthetic GGUF, decodes greedily, and rank 0 checks tokens / logits against the CPU oracle."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")         # host-side plumbing only (id broadcast, barriers); the data path is NCCL inside the engine
    import modelzoo as Z
    from zerfoo_b200 import engine
    from oracle import oracle as O
    kinds = sys.argv[1:] or ["llama_tp_q4_k_m", "mixtral_tp_q4_k_m", "llama_tp_q8_0"]
    results = {}
    for kind in kinds:
        if rank == 0:
            path = Z.path(kind)
        dist.barrier()
        path = Z.path(kind)
        try:
            g = engine.load_file_tp(path)
        except engine.EngineError as ex:
            if rank == 0:
                results[kind] = f"skipped: {ex}"
            dist.barrier()
            continue
        got = g.generate(Z.PROMPT, 48)
        logits = g.logits()
        if rank == 0:
            om = O.Model(path)
            ref = om.generate(Z.PROMPT, 48)
            ok_tok = got == ref
            results[kind] = {"tokens_identical": ok_tok, "tp": world}
            assert ok_tok, (kind, got, ref)
        g.reset()
        first = g.prefill(Z.PROMPT)
        toks, ms = g.decode_n(first, 64)
        if rank == 0:
            results[kind]["ms_per_step"] = ms / 64
        g.close()
        dist.barrier()
    if rank == 0:
        print(json.dumps(results))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
