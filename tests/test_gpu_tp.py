"""Tensor-parallel decode across GPUs of one box (NCCL all-reduce inside the captured step): every rank shards the same
synthetic GGUF; greedy tokens must be identical to the CPU oracle.  Needs >= 2 GPUs (skipped otherwise); the ranks are
launched with torch.distributed.run exactly as bench.py is."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [2, 4, 8])
def test_tp_greedy_tokens_match_oracle(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    kinds = ["llama_tp_q4_k_m", "mixtral_tp_q4_k_m", "llama_tp_q8_0"] if n == 2 else ["llama_tp8_q4_k_m"]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + n), os.path.join(ROOT, "tools", "tp_check.py")] + kinds
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    for kind in kinds:
        assert isinstance(res[kind], dict) and res[kind]["tokens_identical"], res
