#!/usr/bin/env python
"""bench.py -- decode tok/s of the B200-native decode hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|...]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

One "step" is one greedy decode step (one token, batch 1) through the whole stack on a synthetic-weight GGUF of a
BASELINE config, KV cache primed with a fixed 17-token prompt.

Main line.  ONE model at every N, sharded N ways (tensor parallel: QKV / gate-up rows and o / down columns split across the
ranks, one all-reduce after o_proj and down_proj, lm_head rows split + all-gather) -- "scaling": "strong".  The model is
BASELINE config 4, the Llama-3-70B shape in Q4_K_M at full depth (80 layers, 42 GB: it fits one B200, so N = 1 runs the
same workload and the driver's 1 -> 8 curve measures the collective, not eight copies of a small model).  ZB_BENCH_MAIN /
--workload pick another config.  Weights are far larger than the 126 MB L2: every step streams them from HBM.

  value     whole-job tok/s, device-resident: K chained graph launches, the token never leaves the GPU, CUDA events on the
            engine stream, max over ranks.
  e2e       the same metric through the public per-token API (zb_engine_decode_step): every step copies the token id from
            pinned host memory to the device and reads the greedy argmax back (4 B + 4 B), host clock around K steps.
  roofline  the dominant kernel (the GEMV of the majority block format): algorithmic bytes per launch / CUDA-event launch
            time, against MEASURED_PEAKS.json; step_hbm_frac = all weight bytes of the step / step time.
  also      (N = 1) the single-GPU BASELINE configs as sub-records, each with value / ms_per_step / e2e / roofline:
            c2 (Llama-3.2-3B shape Q4_K_M, B=1, CUDA graph -- the config the round-1 headline was quoted on), c1 (Gemma-3-1B
            shape Q4_0), c3 B=32 over the paged KV cache (tcgen05 dequant-GEMMs), c2 at a 4096-token context (KV bytes in
            the roofline denominator).
  cpu_baseline  the restated reference CPU engine (oracle/, "port") timed on this box's host cores on a bounded sample of
            the main workload (N = 1 only).

--impl reference times that CPU restatement as the arm itself, on all host threads (the reference's Go engine cannot be
built here: no Go toolchain, arithmetic in un-vendored ztensor).  Same config keys as our arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PROMPT = [2] + list(range(100, 116))          # fixed prompt token ids
WORKLOADS = {
    "c1": "Gemma-3-1B-shape Q4_0 synthetic GGUF, greedy decode, batch 1",
    "c2": "Llama-3.2-3B-shape Q4_K_M synthetic GGUF, greedy decode, batch 1, CUDA graph",
    "c3": "Mistral-7B-shape Q5_K_M synthetic GGUF, greedy decode",
    "c4": "Llama-3-70B-shape Q4_K_M synthetic GGUF, greedy decode, batch 1, tensor parallel",
    "c5": "Mixtral-8x7B-shape Q4_K_M synthetic GGUF, greedy decode, batch 1, experts sharded",
}


def model_dir() -> str:
    d = os.environ.get("ZB_BENCH_MODEL_DIR") or os.path.join(tempfile.gettempdir(), "zb200_models")
    os.makedirs(d, exist_ok=True)
    return d


def model_path(workload: str, layers=None, ctx=None, fast=False) -> str:
    """Synthetic GGUF of a BASELINE config, cached under the temp dir.  fast=True: matmul tensors are random block bytes of
    the same scale (zerfoo_b200/gguf.py random_blocks) instead of quantized random floats -- 50x faster to write, which is
    what makes the 42 GB 70B shape usable; the timing does not depend on the weight values."""
    from zerfoo_b200 import gguf as G
    if os.environ.get("ZB_BENCH_QUANTIZED_WEIGHTS") == "1":   # A/B: weights quantized from N(0, 0.02) floats as in round 1
        fast = False
    tag = workload if layers is None else f"{workload}_l{layers}"
    if ctx is not None:
        tag += f"_c{ctx}"
    if fast:
        tag += "_fast"
    p = os.path.join(model_dir(), f"bench_{tag}_s1234.gguf")
    if not os.path.exists(p):
        spec = G.preset(workload, layers=layers, ctx=ctx)
        tmp = p + f".tmp{os.getpid()}"
        G.write_synthetic_gguf(tmp, spec, seed=1234, fast=fast)
        os.replace(tmp, p)
    return p


def main_workload(args) -> str:
    return args.workload or os.environ.get("ZB_BENCH_MAIN") or "c4"


def config_for(wl: str, world: int, layers=None) -> dict:
    """The `config` object of a decode line: identical for our arm and the reference arm (no GPU needed to build it)."""
    from zerfoo_b200 import gguf as G
    spec = G.preset(wl, layers=layers)
    return {"workload": WORKLOADS[wl], "batch": 1, "prompt_tokens": len(PROMPT), "parallelism": "single" if world == 1 else f"tp{world}",
            "cuda_graph": True, "l2": "weights >> 126 MB L2: every step streams them from HBM, no flush needed",
            "arch": spec.arch, "layers": spec.layers, "hidden": spec.hidden, "vocab": spec.vocab,
            "weights": "synthetic random quant blocks (std 0.02), seed 1234"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int = 0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 9:
                self.rows.append(parts)

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def cpu_reference(path: str, steps: int, warmup: int, budget_s: float, prompt=None):
    """The restated reference CPU engine on ALL host cores: prompt, `warmup` untimed decode
    tokens, then up to `steps` timed decode tokens (stops early when budget_s is spent)."""
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm sets its own thread count
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import oracle as O
    O.set_num_threads(os.cpu_count() or 1)
    prompt = list(PROMPT if prompt is None else prompt)
    om = O.Model(path, max_seq=len(prompt) + warmup + steps + 8)
    cores = O.num_threads()
    t0 = time.perf_counter()
    for t in prompt[:-1]:
        om.forward(t, want_logits=False)
    tok = O.argmax(om.forward(prompt[-1]))
    for _ in range(warmup):
        tok = O.argmax(om.forward(tok))
    prefill_s = time.perf_counter() - t0
    done = 0
    t1 = time.perf_counter()
    while done < steps:
        tok = O.argmax(om.forward(tok))
        done += 1
        if time.perf_counter() - t1 > budget_s:
            break
    dt = time.perf_counter() - t1
    om.close()
    return {"tok_s": done / dt, "steps": done, "seconds": dt, "cores": cores, "prefill_s": prefill_s}


CPU_PROMPT = PROMPT[:2]   # the CPU arm on the 70B shape runs ~3 s per token: a short prompt keeps the arm within minutes


def run_reference(args):
    """The reference arm: the restated reference CPU engine on the main workload, all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = main_workload(args)
    big = wl in ("c4", "c5")
    path = model_path(wl, layers=args.layers, fast=True)
    W = args.warmup                       # same warm-up as our arm; the timed part is bounded by wall clock instead
    r = cpu_reference(path, args.steps, W, budget_s=90.0, prompt=CPU_PROMPT if big else PROMPT)
    cfg = config_for(wl, 1 if args.gpus <= 1 else args.gpus, args.layers)
    line = {
        "impl": "reference", "metric": "decode_tok_per_s", "value": r["tok_s"], "unit": "tok/s", "n_gpus": args.gpus, "steps": r["steps"],
        "warmup": W, "ms_per_step": 1000.0 / r["tok_s"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": r["tok_s"], "unit": "tok/s", "cores": r["cores"], "kind": "port",
                         "sample": f"{r['steps']} timed decode tokens ({r['seconds']:.1f} s) after a {len(CPU_PROMPT if big else PROMPT)}-token prompt and "
                                   f"{W} warm-up token(s) on the same GGUF (CPU restatement of the reference engine, row-parallel over "
                                   f"{r['cores']} host threads; the Go engine cannot be built here)"},
        "e2e": {"value": r["tok_s"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


T0 = time.time()


def progress(msg: str) -> None:
    """Timestamped phase marker on stderr (the JSON line is the only thing on stdout)."""
    print(f"[bench +{time.time() - T0:6.1f}s rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


class Watchdog:
    """A multi-rank run that stalls (a peer died, a collective never completes) must not eat the caller's whole time limit:
    after `seconds` rank 0 prints the best line it has (or an error line) and every rank leaves with os._exit."""

    def __init__(self, seconds: float, rank: int):
        self.line = None
        self.rank = rank
        self.t = threading.Timer(seconds, self.fire)
        self.t.daemon = True
        self.t.start()

    def fire(self):
        progress("watchdog: time limit reached, leaving")
        if self.rank == 0:
            line = self.line or {"metric": "decode_tok_per_s", "value": None, "unit": "tok/s", "error": "bench.py watchdog: the run stalled before a measurement was complete"}
            line = dict(line)
            line["watchdog"] = "fired"
            print(json.dumps(line), flush=True)
        os._exit(0 if (self.line or self.rank != 0) else 3)   # the other ranks have nothing to print: leave quietly

    def cancel(self):
        self.t.cancel()


def dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        # host-side plumbing only (id broadcast, barriers, max over ranks): gloo.  The data path's collectives live inside the
        # engine (its own NCCL communicator / peer-memory exchange); a second, torch-owned NCCL communicator on the same GPUs is
        # not needed and is one more thing that can interleave badly with captured collectives.
        dist.init_process_group("gloo")
    return world, rank, local


def decode_record(g, wl, K, W, world, rank, local, with_clocks=True, roofline=True, ctx_note=None):
    """Device-timed value, end-to-end value and (rank 0) the roofline of an engine that already holds its prompt.
    Returns (record dict on rank 0 / None elsewhere, tokens)."""
    import torch
    import torch.distributed as dist
    from zerfoo_b200 import gguf as G

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    info = g.refresh_info()
    first = g.last_first
    toks, _ = g.decode_n(first, W)
    barrier()
    sampler = ClockSampler(local).start() if (rank == 0 and with_clocks) else None
    toks2, ms = g.decode_n(toks[-1], K)
    barrier()
    ms_max = max_over_ranks(ms)
    value = K / (ms_max / 1000.0)
    # e2e: public per-token API, host token in / argmax out every step
    tok = toks2[-1]
    for _ in range(W):
        tok = g.decode_step(tok)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        tok = g.decode_step(tok)
    torch.cuda.synchronize()
    e2e = K / max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if sampler else None
    # The eager profiling steps are whole decode steps: on a tensor-parallel engine they contain the exchange, so EVERY rank
    # runs them (rank 0 alone would wait for its peers forever); only rank 0 uses the numbers.
    prof = g.profile_gemv(4) if roofline else None
    dom_graph = None
    if roofline:
        dom = max(prof, key=lambda r: r[2])
        dom_graph = g.profile_gemv_graph(dom[0], 4)    # local GEMVs only, no collective
    if rank != 0:
        return None, [first] + toks + toks2
    pk = peaks()
    kv_bytes = info.kv_bytes_per_pos * (g.position - K // 2) if ctx_note else 0   # mean context over the timed steps
    step_bytes = info.weight_bytes_per_token + kv_bytes
    rec = {"value": value, "unit": "tok/s", "ms_per_step": ms_max / K, "steps": K, "warmup": W,
           "e2e": {"value": e2e, "unit": "tok/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 4},
           "gpu_launches": info.launches_per_step * K, "launches_per_step": info.launches_per_step, "kv_len_at_end": g.position}
    if clocks is not None:
        rec["clocks"] = clocks
    rf = {"bound": "hbm", "peak": pk["hbm_gbs"], "unit": "GB/s", "peak_source": pk["source"],
          "step_weight_bytes": info.weight_bytes_per_token, "step_kv_bytes": kv_bytes,
          "step_hbm_frac": (step_bytes / ((ms_max / K) / 1000.0) / 1e9) / pk["hbm_gbs"]}
    if roofline:
        # dominant kernel = the GEMV of the block format that carries most of the step's bytes.  Its average launch duration is
        # measured in steady state: all its launches of one step, PDL-chained in a CUDA graph exactly as in the decode step,
        # 4 replays between two CUDA events on the engine stream (weights of the class >> L2, every launch streams from HBM).
        gl, gb, gms = dom_graph
        ach = gb / (gms / 1000.0) / 1e9
        traffic = None   # dram__bytes_read+write per launch from the committed ncu --set full capture of this class, else null
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            ratio = json.load(open(tp)).get(f"{wl}:{G.TYPE_NAMES[dom[0]]}")
            if ratio and ratio.get("dram_bytes_over_algorithmic"):
                traffic = ratio["dram_bytes_over_algorithmic"] * gb / gl
        rf.update({"kernel": gemv_kernel_name(dom[0]), "achieved": ach, "frac": ach / pk["hbm_gbs"], "traffic": traffic,
                   "frac_of_8TBs_nominal": ach / 8000.0, "bytes_per_launch": gb / gl, "us_per_launch": 1000.0 * gms / gl, "launches_profiled": gl,
                   "timing": "CUDA events around 4 graph replays of all launches of this kernel class in one step (PDL-chained)",
                   "all_formats": [{"format": G.TYPE_NAMES[r[0]], "launches_per_step": r[1] // 4, "eager_GBps": r[2] / (r[3] / 1000.0) / 1e9}
                                   for r in prof]})
    else:
        ach = step_bytes / ((ms_max / K) / 1000.0) / 1e9
        rf.update({"kernel": "whole decode step (weight + KV bytes / step time)", "achieved": ach, "frac": ach / pk["hbm_gbs"], "traffic": None})
    rec["roofline"] = rf
    return rec, [first] + toks + toks2


def also_records(K, W):
    """N = 1: the single-GPU BASELINE configs as sub-records of the line (each a short, complete measurement)."""
    import torch
    from zerfoo_b200 import engine
    out = []
    Ka, Wa = min(K, 32), max(min(W, 5), 3)

    def guarded(name, fn):
        try:
            out.append(fn())
        except Exception as ex:   # a sub-record must never take the main line down
            out.append({"workload": name, "error": f"{type(ex).__name__}: {ex}"[:300]})

    def b1(wl, note=None):
        def run():
            path = model_path(wl, fast=True)
            g = engine.load_file(path, max_seq=max(512, len(PROMPT) + 3 * (Ka + Wa) + 64))
            g.last_first = g.prefill(PROMPT)
            rec, _ = decode_record(g, wl, Ka, Wa, 1, 0, 0, with_clocks=False)
            g.close()
            rec = {"workload": WORKLOADS[wl], "config": config_for(wl, 1), **rec}
            return rec
        return run

    def long_ctx(wl, ctx, kv_f16=False):
        def run():
            import numpy as np
            path = model_path(wl, ctx=ctx + 256, fast=True)
            g = engine.load_file(path, max_seq=ctx + 3 * (Ka + Wa) + 64, kv_f16=kv_f16)
            rng = np.random.default_rng(0)
            prompt = [int(t) for t in rng.integers(1, g.info.vocab, size=ctx)]
            if kv_f16:      # the chunked prefill writes an f32 cache: an fp16 engine takes its prompt through the decode kernels
                g.last_first = g.prefill(prompt)
            else:
                g.last_first, _ = g.prefill_chunked(prompt)
            rec, _ = decode_record(g, wl, Ka, Wa, 1, 0, 0, with_clocks=False, roofline=False, ctx_note=ctx)
            g.close()
            cfg = config_for(wl, 1)
            cfg["prompt_tokens"] = ctx
            cfg["kv"] = ("fp16" if kv_f16 else "f32") + " [n_kv][max_seq][hd]"
            return {"workload": WORKLOADS[wl] + f", {ctx}-token context, {'fp16' if kv_f16 else 'f32'} KV cache", "config": cfg, **rec}
        return run

    def batched(wl, B):
        def run():
            path = model_path(wl, fast=True)
            g = engine.load_file(path, batch=B, max_seq=max(256, len(PROMPT) + 2 * (Ka + Wa) + 32))
            info = g.refresh_info()
            g.batch_reset()
            last = None
            for t in PROMPT:
                last = g.batch_step([(t + b) % info.vocab for b in range(B)])
            o, _ = g.batch_decode_n(last, Wa)
            torch.cuda.synchronize()
            o2, ms = g.batch_decode_n(list(map(int, o[-1])), Ka)
            torch.cuda.synchronize()
            tok = list(map(int, o2[-1]))
            for _ in range(Wa):
                tok = g.batch_step(tok)
            t0 = time.perf_counter()
            for _ in range(Ka):
                tok = g.batch_step(tok)
            torch.cuda.synchronize()
            e2e = B * Ka / (time.perf_counter() - t0)
            pk = peaks()
            ach = info.weight_bytes_per_token / ((ms / Ka) / 1000.0) / 1e9
            g.close()
            cfg = config_for(wl, 1)
            cfg["batch"] = B
            cfg["kv"] = "paged, 16-position blocks"
            return {"workload": WORKLOADS[wl] + f", batch {B} over paged KV", "config": cfg, "value": B * Ka / (ms / 1000.0), "unit": "tok/s",
                    "ms_per_step": ms / Ka, "steps": Ka, "warmup": Wa, "dtype": "bf16 operands / f32 accumulate (tcgen05)",
                    "e2e": {"value": e2e, "unit": "tok/s", "h2d_bytes_per_step": 4 * B, "d2h_bytes_per_step": 4 * B},
                    "gpu_launches": info.launches_per_step * Ka, "launches_per_step": info.launches_per_step,
                    "roofline": {"bound": "hbm", "kernel": "gemm_tc_kernel (whole step: weight bytes once per step / step time)", "achieved": ach,
                                 "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None, "peak_source": pk["source"],
                                 "step_weight_bytes": info.weight_bytes_per_token}}
        return run

    def prefill(wl, n):
        def run():
            import numpy as np
            path = model_path(wl, fast=True)
            g = engine.load_file(path, max_seq=n + 64)
            info = g.refresh_info()
            rng = np.random.default_rng(0)
            prompt = [int(t) for t in rng.integers(1, info.vocab, size=n)]
            g.prefill_chunked(prompt)
            dev_ms, wall, reps = 0.0, 0.0, 2
            for _ in range(reps):
                g.reset()
                t0 = time.perf_counter()
                _, ms = g.prefill_chunked(prompt)
                wall += time.perf_counter() - t0
                dev_ms += ms
            g.close()
            cfg = config_for(wl, 1)
            cfg["prompt_tokens"] = n
            cfg["chunk"] = 256
            from zerfoo_b200 import gguf as G
            sp = G.preset(wl)
            qd, kvd = sp.n_q * sp.head_dim, sp.n_kv * sp.head_dim
            weights = sp.layers * (sp.hidden * (qd + 2 * kvd) + qd * sp.hidden + 3 * sp.hidden * sp.ffn)   # the head runs for the last token only
            flops = 2.0 * weights * n
            return {"workload": WORKLOADS[wl] + f", {n}-token prompt prefill in 256-token chunks (tcgen05 dequant-GEMMs)", "config": cfg,
                    "metric": "prefill_tok_per_s", "value": n * reps / (dev_ms / 1000.0), "unit": "tok/s", "ms_per_step": dev_ms / reps, "steps": reps,
                    "warmup": 1, "dtype": "bf16 operands / f32 accumulate (tcgen05), f32 attention",
                    "e2e": {"value": n * reps / wall, "unit": "tok/s", "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": 4},
                    "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (whole prompt: 2 * weights * tokens / time)",
                                 "achieved": flops / (dev_ms / reps / 1000.0) / 1e12, "unit": "TFLOP/s", "traffic": None}}
        return run

    guarded("c2", b1("c2"))
    guarded("c1", b1("c1"))
    guarded("c3 B=32", batched("c3", 32))
    guarded("c3 prefill 4096", prefill("c3", 4096))
    guarded("c2 ctx 4096", long_ctx("c2", 4096))
    guarded("c2 ctx 4096 fp16 KV", long_ctx("c2", 4096, kv_f16=True))   # last: the newest path goes where nothing depends on it
    return out


def also_records_tp(K, W, world, rank, local):
    """N = 8: BASELINE config 5 (Mixtral 8x7B shape, Q4_K_M, experts sharded across the 8 ranks) as a sub-record.  Collective:
    every rank runs it; errors and stalls must not take the main line down (try / except here, the watchdog above)."""
    import torch.distributed as dist
    from zerfoo_b200 import engine
    wl = "c5"
    out = []
    try:
        # rank 0 writes the 26 GB file (if the disk has room) and tells everybody whether to go on: no rank may be left
        # waiting at a collective that another rank never reaches
        go = [None]
        if rank == 0:
            try:
                import shutil
                p0 = os.path.join(model_dir(), f"bench_{wl}_fast_s1234.gguf")
                if not os.path.exists(p0) and shutil.disk_usage(model_dir()).free < 40e9:
                    go[0] = "skipped: less than 40 GB free under the model directory"
                else:
                    model_path(wl, fast=True)
                    go[0] = "ok"
            except Exception as ex:
                go[0] = f"skipped: {type(ex).__name__}: {ex}"[:200]
        dist.broadcast_object_list(go, src=0)
        if go[0] != "ok":
            return [{"workload": WORKLOADS[wl], "error": go[0]}] if rank == 0 else []
        path = model_path(wl, fast=True)
        Ka, Wa = min(K, 32), max(min(W, 5), 3)
        g = engine.load_file_tp(path, max_seq=max(512, len(PROMPT) + 3 * (Ka + Wa) + 64))
        progress("c5 engine loaded")
        g.last_first = g.prefill(PROMPT)
        rec, _ = decode_record(g, wl, Ka, Wa, world, rank, local, with_clocks=False, roofline=False)
        g.close()
        progress("c5 measured")
        if rank == 0:
            out.append({"workload": WORKLOADS[wl], "config": config_for(wl, world), **rec})
    except Exception as ex:
        if rank == 0:
            out.append({"workload": WORKLOADS[wl], "error": f"{type(ex).__name__}: {ex}"[:300]})
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from zerfoo_b200 import engine

    world, rank, local = dist_setup()
    wl = main_workload(args)
    dog = Watchdog(float(os.environ.get("ZB_BENCH_LIMIT_S", "780")), rank) if world > 1 else None
    progress(f"start: workload {wl}, world {world}")
    if rank == 0:
        model_path(wl, layers=args.layers, fast=True)
        progress("model file ready")
    if world > 1:
        dist.barrier()
    path = model_path(wl, layers=args.layers, fast=True)
    K, W = args.steps, max(args.warmup, 3)
    max_seq = max(512, len(PROMPT) + 3 * (K + W) + 64)
    g = engine.load_file(path, device=local, max_seq=max_seq) if world == 1 else engine.load_file_tp(path, max_seq=max_seq)
    progress("engine loaded")
    g.last_first = g.prefill(PROMPT)
    progress("prompt done")
    rec, toks = decode_record(g, wl, K, W, world, rank, local)
    progress("decode measured")

    def make_line(extra, cpu=None, also=None, identical=None):
        line = {"metric": "decode_tok_per_s", "value": rec["value"], "unit": "tok/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {**config_for(wl, world, args.layers), "kv_len_at_end": rec["kv_len_at_end"]},
                "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"], "launches_per_step": rec["launches_per_step"], "clocks": rec.get("clocks"),
                "roofline": rec["roofline"], "cpu_baseline": cpu, "tokens_identical_to_n1": identical, **extra}
        if also is not None:
            line["also"] = also
        return line

    extra = {}
    if dog and rank == 0:
        dog.line = make_line(extra)          # from here on a stall still leaves a complete measurement
    if world > 1:
        ar = g.tp_allreduce_us(2 * g.info.layers, 4)
        progress("all-reduce timed")
        if rank == 0:
            extra["allreduce_us_per_step"] = ar
            extra["exchange"] = g.tp_exchange
    g.close()
    also_tp = None
    if world == 8 and not args.no_also and wl == "c4" and args.layers is None:
        if dog and rank == 0:
            dog.line = make_line(dict(extra))    # the main line is complete: whatever happens to the sub-record, it gets printed
        also_tp = also_records_tp(K, W, world, rank, local)
    if rank != 0:
        if dog:
            dog.cancel()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # greedy tokens must not depend on the sharding: the N = 1 run leaves its tokens for the N > 1 runs of the same box
    tok_file = os.path.join(model_dir(), f"tokens_{wl}_{args.layers}_{K}_{W}.json")
    identical = None
    if world == 1:
        json.dump(toks, open(tok_file, "w"))
    elif os.path.exists(tok_file):
        ref = json.load(open(tok_file))
        identical = ref[:len(toks)] == toks[:len(ref)]

    cpu = None
    if world == 1 and not args.no_cpu:
        big = wl in ("c4", "c5")
        r = cpu_reference(path, steps=16, warmup=1, budget_s=20.0, prompt=CPU_PROMPT if big else PROMPT)
        cpu = {"value": r["tok_s"], "unit": "tok/s", "cores": r["cores"], "kind": "port",
               "sample": f"{r['steps']} timed decode tokens ({r['seconds']:.1f} s) after a {len(CPU_PROMPT if big else PROMPT)}-token prompt on the same GGUF; "
                         "CPU restatement of the reference engine (oracle/), row-parallel over all host threads"}
    also = also_records(K, W) if (world == 1 and not args.no_also) else also_tp

    line = make_line(extra, cpu, also, identical)
    if dog:
        dog.cancel()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def gemv_kernel_name(qtype):
    """Which batch-1 GEMV kernel the engine runs for this block format (engine.cu upload policy; ZB_GEMV_TC=0 forces the CUDA-core kernel)."""
    from zerfoo_b200 import gguf as G
    name = G.TYPE_NAMES[qtype]
    if os.environ.get("ZB_GEMV_TC", "1") != "0" and qtype in (G.Q4_K, G.Q6_K, G.Q4_0):
        path = "IMMA m16n8k32 u8 x s8, exact integer dot products" if os.environ.get("ZB_MMA_I8", "1") != "0" else "HMMA m16n8k16 f16"
        note = " (Q4_0: rows >= 4096 wide or matrices >= 32 MB; the rest on gemv_stream_kernel)" if qtype == G.Q4_0 else ""
        return f"gemv_mma_kernel<{name}> (tensor-core batch-1 GEMV, {path}){note}"
    return f"gemv_stream_kernel<{name}> (CUDA-core batch-1 GEMV)"


def run_batched(args):
    """B sequences decode in lock-step (BASELINE config 3 is B=32 on the Mistral-7B shape, Q5_K_M)."""
    import torch
    from zerfoo_b200 import engine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(0)
    wl = args.workload or "c3"
    path = model_path(wl, layers=args.layers)
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    g = engine.load_file(path, batch=B, max_seq=max(256, len(PROMPT) + 2 * (K + W) + 32))
    info = g.refresh_info()
    g.batch_reset()
    last = None
    for t in PROMPT:                       # every sequence gets the fixed prompt, shifted by its index so the batch is not degenerate
        last = g.batch_step([(t + b) % info.vocab for b in range(B)])
    out, _ = g.batch_decode_n(last, W)
    torch.cuda.synchronize()
    sampler = ClockSampler(0).start()
    out2, ms = g.batch_decode_n(list(map(int, out[-1])), K)
    torch.cuda.synchronize()
    value = B * K / (ms / 1000.0)
    tok = list(map(int, out2[-1]))
    for _ in range(W):
        tok = g.batch_step(tok)
    t0 = time.perf_counter()
    for _ in range(K):
        tok = g.batch_step(tok)
    torch.cuda.synchronize()
    e2e = B * K / (time.perf_counter() - t0)
    clocks = sampler.stop()
    pk = peaks()
    ach = info.weight_bytes_per_token / ((ms / K) / 1000.0) / 1e9
    line = {
        "metric": "decode_tok_per_s", "value": value, "unit": "tok/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 operands / f32 accumulate (tcgen05)", "data": "synthetic",
        "config": {"workload": WORKLOADS[wl] + f", batch {B} over paged KV", "batch": B, "prompt_tokens": len(PROMPT), "cuda_graph": True,
                   "l2": "weights >> 126 MB L2: every step streams them from HBM, no flush needed", "arch": info.arch.decode(),
                   "layers": info.layers, "hidden": info.hidden, "vocab": info.vocab},
        "e2e": {"value": e2e, "unit": "tok/s", "h2d_bytes_per_step": 4 * B, "d2h_bytes_per_step": 4 * B},
        "gpu_launches": info.launches_per_step * K, "launches_per_step": info.launches_per_step, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "gemm_tc_kernel (whole step: weight bytes once per step / step time)", "achieved": ach,
                     "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None, "peak_source": pk["source"],
                     "step_weight_bytes": info.weight_bytes_per_token},
        "cpu_baseline": None,
    }
    print(json.dumps(line))
    g.close()
    return 0


def run_prefill(args):
    """Prompt prefill of `--prefill N` tokens (BASELINE config 3: 4k tokens on the Mistral-7B shape) through the chunked
    tcgen05 path; a step = one whole prompt.  Reported as prompt tokens per second; flops = 2 * weights * tokens."""
    import torch
    from zerfoo_b200 import engine
    import numpy as np
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(0)
    wl = args.workload or "c3"
    n = args.prefill
    from zerfoo_b200 import gguf as G
    path = model_path(wl, layers=args.layers, ctx=None if n + 64 <= G.preset(wl).ctx else n + 64)
    g = engine.load_file(path, max_seq=n + 64)
    info = g.refresh_info()
    rng = np.random.default_rng(0)
    prompt = [int(t) for t in rng.integers(1, info.vocab, size=n)]
    K, W = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    for _ in range(W):
        g.reset()
        g.prefill_chunked(prompt)
    sampler = ClockSampler(0).start()
    dev_ms, wall = 0.0, 0.0
    for _ in range(K):
        g.reset()
        t0 = time.perf_counter()
        _, ms = g.prefill_chunked(prompt)
        wall += time.perf_counter() - t0
        dev_ms += ms
    clocks = sampler.stop()
    pk = peaks()
    value = n * K / (dev_ms / 1000.0)
    line = {
        "metric": "prefill_tok_per_s", "value": value, "unit": "tok/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": dev_ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 operands / f32 accumulate (tcgen05), f32 attention",
        "data": "synthetic",
        "config": {"workload": WORKLOADS[wl] + f", {n}-token prompt in 256-token chunks", "prompt_tokens": n, "chunk": 256,
                   "l2": "weights >> 126 MB L2: every chunk streams them from HBM", "arch": info.arch.decode(), "layers": info.layers,
                   "hidden": info.hidden, "vocab": info.vocab},
        "e2e": {"value": n * K / wall, "unit": "tok/s", "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": 4},
        "clocks": clocks, "cpu_baseline": None,
    }
    print(json.dumps(line))
    g.close()
    return 0


def g_position_note(prompt, w, k):
    return prompt + w + k


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-also", action="store_true", help="skip the single-GPU sub-records (N = 1)")
    ap.add_argument("--layers", type=int, default=None, help="override the layer count of the workload (reported in config)")
    ap.add_argument("--batch", type=int, default=1, help="decode batch (sequences in lock-step over the paged KV cache, tcgen05 GEMMs)")
    ap.add_argument("--prefill", type=int, default=0, help="time a chunked prefill of this many prompt tokens instead of decode")
    args = ap.parse_args()
    if args.prefill > 0 and args.impl != "reference":
        return run_prefill(args)
    if args.impl == "reference":
        return run_reference(args)
    if args.batch > 1:
        return run_batched(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
