"""Data-parallel equivalence on real GPUs: overlapped path (asynchronous weight gradients + bucketed all-reduce behind
the side stream) == plain path (synchronous gradients, one all-reduce).  Needs two GPUs; the world_size-2 gloo tests in
tests/test_parallel_cpu.py cover the host logic on the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_overlapped_data_parallel_gradients_equal_plain_all_reduce():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")],
                       capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith('{"dp_check"')]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    assert json.loads(lines[-1])["dp_check"] == "ok"
