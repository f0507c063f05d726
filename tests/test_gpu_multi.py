"""Sharded global merge on real GPUs (>= 2 visible): torchrun + NCCL, compared with the oracle.
Skipped on a single-GPU box; the host-side plan is covered on CPU by test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_merge_over_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "scripts", "multigpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "multigpu_check ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
