"""Data-parallel GNN training on NCCL (needs >= 2 GPUs; skipped on a 1-GPU box — the builder's 2- and 8-GPU runs of the same
script are kept under profiles/r2_dp_grad_check_*.json): the bucket all-reduced with ReduceOp.AVG over N ranks x batch B equals
the gradient of the concatenated N*B batch to 1e-5 (fp32 summation order)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nccl_averaged_gradient_equals_big_batch_gradient():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(ROOT, "tools", "dp_grad_check.py")], capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads(lines[-1])
    assert r["ok"] and r["max_rel_err"] < 1e-5
