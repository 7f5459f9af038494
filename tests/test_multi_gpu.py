"""GPU, >= 2 devices: the public API under torchrun (NCCL table broadcast + sharded subjects)
reproduces the reference's files.  Skipped on single-GPU boxes; the same host logic is covered
on CPU by test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["g2_edges", "g3_pop3_messy"])
def test_torchrun_two_gpus_matches_golden(case, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300), os.path.join(HERE, "multi_gpu_worker.py"),
           case, str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIFFERENT" not in r.stdout
