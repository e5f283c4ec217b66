"""Host-side logic of bench.py that needs no GPU: both arms share one config object, the fast synthetic-weight generator is
deterministic and statistically what it says, the watchdog leaves a line instead of hanging."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from zerfoo_b200 import gguf as G  # noqa: E402
import refdata as R  # noqa: E402


def test_both_arms_build_the_same_config():
    for wl in ("c1", "c2", "c3", "c4", "c5"):
        a, b = bench.config_for(wl, 1), bench.config_for(wl, 1)
        assert a == b and {"workload", "arch", "layers", "hidden", "vocab", "parallelism", "batch"} <= set(a)
    assert bench.config_for("c4", 8)["parallelism"] == "tp8"
    assert bench.config_for("c4", 1)["layers"] == 80 and bench.config_for("c4", 1, layers=8)["layers"] == 8


def test_fast_random_blocks_are_deterministic_and_scaled(tmp_path):
    for qt in (G.Q4_0, G.Q8_0, G.Q4_K, G.Q5_K, G.Q6_K):
        a = G.random_blocks(qt, 2048, np.random.Generator(np.random.SFC64([1, 2, 3])), 0.02)
        b = G.random_blocks(qt, 2048, np.random.Generator(np.random.SFC64([1, 2, 3])), 0.02)
        assert np.array_equal(a, b)
        w = R.np_dequant(qt, a)
        assert np.isfinite(w).all() and 0.015 < float(w.std()) < 0.025 and abs(float(w.mean())) < 0.004
    spec = G.ModelSpec("llama", 256, 256, 2, 4, 2, 64, 512, ctx=128, base_type=G.Q4_K, more_bits_type=G.Q6_K, embed_type=G.Q6_K)
    p1, p2 = str(tmp_path / "a.gguf"), str(tmp_path / "b.gguf")
    G.write_synthetic_gguf(p1, spec, seed=7, fast=True)
    G.write_synthetic_gguf(p2, spec, seed=7, fast=True)
    assert open(p1, "rb").read() == open(p2, "rb").read()
    # and the CPU engine runs on such a file with finite logits
    from oracle import oracle as O
    om = O.Model(p1)
    lg = om.forward(3)
    assert np.isfinite(lg).all()
    om.close()


def test_watchdog_prints_the_best_line_and_exits():
    code = (
        "import sys, time; sys.path.insert(0, %r); import bench\n"
        "d = bench.Watchdog(0.3, 0); d.line = {'metric': 'decode_tok_per_s', 'value': 1.0}\n"
        "time.sleep(5); print('not reached')\n" % ROOT
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "not reached" not in out.stdout
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["value"] == 1.0 and line["watchdog"] == "fired"


def test_reference_arm_under_torchrun_uses_all_host_threads(tmp_path):
    """The driver launches the reference arm like ours (torchrun for N > 1, which exports OMP_NUM_THREADS=1): rank 0 alone
    works, with every host core, and prints our arm's config; the other ranks leave quietly.  (Round 1 reported cores: 1 here.)"""
    env = dict(os.environ, ZB_BENCH_MODEL_DIR=str(tmp_path), OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29733",
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "c1", "--layers", "1", "--steps", "3", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                   # rank 0 only
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["warmup"] == 1
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and line["cpu_baseline"]["kind"] == "port"
    assert line["config"] == bench.config_for("c1", 2, 1)
    assert line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0


def test_c5_subrecord_skips_collectively_when_the_disk_is_short(tmp_path):
    """N = 8 sub-record (Mixtral shape, experts sharded): when rank 0 cannot write the 26 GB file every rank must learn it
    (one broadcast) and skip together -- nobody may be left at a collective the others never reach.  gloo, world 2, CPU."""
    script = tmp_path / "skip.py"
    script.write_text(
        "import os, sys, json, shutil, collections\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import torch.distributed as dist\n"
        "import bench\n"
        "dist.init_process_group('gloo')\n"
        "U = collections.namedtuple('U', 'total used free')\n"
        "shutil.disk_usage = lambda p: U(10, 9, 1)\n"
        "out = bench.also_records_tp(20, 5, 2, dist.get_rank(), 0)\n"
        "print('RANK', dist.get_rank(), json.dumps(out), flush=True)\n"
        "dist.barrier(); dist.destroy_process_group()\n")
    env = dict(os.environ, ZB_BENCH_MODEL_DIR=str(tmp_path / "models"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29734", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    recs = {int(l.split()[1]): json.loads(l.split(" ", 2)[2]) for l in out.stdout.splitlines() if l.startswith("RANK")}
    assert recs[1] == [] and len(recs[0]) == 1 and "40 GB" in recs[0][0]["error"]
