"""bench.py's N = 1 control flow with a stand-in engine (no GPU): the main line carries every contract key, the sub-records
are complete, a failing sub-record becomes an `error` entry instead of taking the line down, the token file is written."""
import json
import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


class FakeInfo:
    vocab, hidden, layers = 1000, 64, 2
    launches_per_step = 11
    weight_bytes_per_token = 10_000_000
    kv_bytes_per_pos = 1024


class FakeGen:
    def __init__(self, path, **kw):
        self.kw, self.info, self.position, self.batch = kw, FakeInfo(), 0, kw.get("batch", 1)
        if kw.get("kv_f16"):
            self.info = FakeInfo()
            self.info.kv_bytes_per_pos = 512
        if "boom" in path:
            raise RuntimeError("stand-in failure")

    def refresh_info(self):
        return self.info

    def prefill(self, toks):
        self.position += len(toks)
        return 7

    def prefill_chunked(self, toks):
        if self.kw.get("kv_f16"):
            raise RuntimeError("fp16 engines have no chunked prefill")
        self.position += len(toks)
        return 7, 12.5

    def reset(self):
        self.position = 0

    def decode_n(self, first, n):
        self.position += n
        return [first + i + 1 for i in range(n)], 1.5 * n

    def decode_step(self, tok):
        self.position += 1
        return tok + 1

    def profile_gemv(self, steps):
        self.position += steps
        return [(12, 98 * steps, 98 * steps * 14e6, 0.7 * steps), (14, 29 * steps, 29 * steps * 20e6, 0.3 * steps)]

    def profile_gemv_graph(self, qtype, reps):
        return 98 * reps, 98 * reps * 14e6, 0.65 * reps

    def batch_reset(self):
        pass

    def batch_step(self, toks):
        return [t + 1 for t in toks]

    def batch_decode_n(self, first, n):
        import numpy as np
        return np.tile(np.array(first, dtype=np.int32), (n, 1)), 7.5 * n

    def close(self):
        pass


@pytest.fixture()
def fake(monkeypatch, tmp_path):
    import torch
    from zerfoo_b200 import engine
    monkeypatch.setenv("ZB_BENCH_MODEL_DIR", str(tmp_path))
    monkeypatch.setattr(bench, "dist_setup", lambda: (1, 0, 0))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(engine, "load_file", lambda path, **kw: FakeGen(path, **kw))
    monkeypatch.setattr(bench, "model_path", lambda wl, layers=None, ctx=None, fast=False: f"/nonexistent/{wl}_{ctx}.gguf")
    monkeypatch.setattr(bench, "cpu_reference", lambda path, steps, warmup, budget_s, prompt=None:
                        {"tok_s": 0.2, "steps": 3, "seconds": 15.0, "cores": 16, "prefill_s": 1.0})
    return monkeypatch


def run(capsys, argv):
    old = sys.argv
    sys.argv = ["bench.py"] + argv
    try:
        assert bench.main() == 0
    finally:
        sys.argv = old
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_main_line_and_subrecords(fake, capsys, tmp_path):
    line = run(capsys, ["--gpus", "1", "--steps", "20", "--warmup", "5"])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "also"):
        assert k in line, k
    assert line["scaling"] == "strong" and line["n_gpus"] == 1 and line["config"]["parallelism"] == "single" and line["config"]["layers"] == 80
    assert line["value"] == pytest.approx(20 / (1.5 * 20 / 1000.0)) and line["gpu_launches"] == 11 * 20
    rf = line["roofline"]
    assert rf["bound"] == "hbm" and 0 < rf["frac"] == pytest.approx(rf["achieved"] / rf["peak"]) and "step_hbm_frac" in rf
    assert line["e2e"]["h2d_bytes_per_step"] == 4 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 16
    names = [a["workload"] for a in line["also"]]
    assert len(names) == 6 and not any("error" in a for a in line["also"]), line["also"]
    for a in line["also"]:
        assert a["value"] > 0 and "roofline" in a and "e2e" in a and "config" in a
    assert any("fp16 KV" in n for n in names) and any("prefill" in n for n in names) and any("batch 32" in n for n in names)
    assert json.load(open(os.path.join(str(tmp_path), "tokens_c4_None_20_5.json")))[0] == 7


def test_a_failing_subrecord_does_not_take_the_line_down(fake, capsys):
    fake.setattr(bench, "model_path", lambda wl, layers=None, ctx=None, fast=False: f"/nonexistent/{'boom' if wl == 'c1' else wl}.gguf")
    line = run(capsys, ["--gpus", "1", "--steps", "8", "--warmup", "3"])
    errs = [a for a in line["also"] if "error" in a]
    assert len(errs) == 1 and "stand-in failure" in errs[0]["error"] and line["value"] > 0


def test_no_also_no_cpu(fake, capsys):
    line = run(capsys, ["--gpus", "1", "--steps", "8", "--warmup", "3", "--no-also", "--no-cpu", "--workload", "c2"])
    assert "also" not in line and line["cpu_baseline"] is None and line["config"]["layers"] == 28


@pytest.mark.parametrize("world", [2, 8])
def test_multi_rank_flow_over_gloo(tmp_path, world):
    """N = 2 / N = 8 control flow with the stand-in engine on both ranks (real gloo collectives, no GPU): rank 0 prints one line with the
    tensor-parallel keys, the others print nothing, the watchdog is armed and cancelled, tokens are compared with the N = 1 file;
    at N = 8 the Mixtral-shape sub-record runs collectively after the main measurement."""
    import subprocess
    script = tmp_path / "flow.py"
    script.write_text(
        "import os, sys, json\n"
        f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\n"
        "import torch, torch.distributed as dist\n"
        "import bench, test_bench_flow as T\n"
        "from zerfoo_b200 import engine\n"
        "def setup():\n"
        "    dist.init_process_group('gloo')\n"
        "    return dist.get_world_size(), dist.get_rank(), 0\n"
        "bench.dist_setup = setup\n"
        "torch.cuda.synchronize = lambda *a, **k: None\n"
        "class G(T.FakeGen):\n"
        "    tp_exchange = 'stand-in exchange'\n"
        "    def tp_allreduce_us(self, count, reps=4):\n"
        "        t = torch.zeros(1); dist.all_reduce(t); return 123.0\n"
        "engine.load_file_tp = lambda path, **kw: G(path, **kw)\n"
        "bench.model_path = lambda wl, layers=None, ctx=None, fast=False: '/nonexistent/x.gguf'\n"
        f"sys.argv = ['bench.py', '--gpus', '{world}', '--steps', '20', '--warmup', '5']\n"
        "sys.exit(bench.main())\n")
    models = tmp_path / "models"
    models.mkdir()
    # what an N = 1 run of the same box would have left: first token, 5 warm-up tokens, 20 timed tokens of the stand-in
    toks = [7] + [8 + i for i in range(5)] + [13 + i for i in range(20)]
    (models / "tokens_c4_None_20_5.json").write_text(json.dumps(toks))
    env = dict(os.environ, ZB_BENCH_MODEL_DIR=str(models), ZB_BENCH_LIMIT_S="120")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(29735 + world), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["n_gpus"] == world and line["config"]["parallelism"] == f"tp{world}" and line["scaling"] == "strong"
    assert line["allreduce_us_per_step"] == 123.0 and line["exchange"] == "stand-in exchange"
    assert line["tokens_identical_to_n1"] is True and line["cpu_baseline"] is None
    assert "watchdog" not in line and "all-reduce timed" in out.stderr
    if world == 8:
        (c5,) = line["also"]
        assert "Mixtral" in c5["workload"] and c5["value"] > 0 and c5["config"]["parallelism"] == "tp8" and "c5 measured" in out.stderr
    else:
        assert "also" not in line
