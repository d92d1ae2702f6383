"""Multi-GPU device paths (skipped with fewer than 2 devices): the row-sharded unsupervised step must reproduce the single-GPU
step on the union batch (tools/check_unsup_dp.py under torchrun, 2 ranks, NCCL)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_row_sharded_unsup_step_equals_single_gpu_step():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tools", "check_unsup_dp.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert r.stdout.count("loss rel err") == 2
