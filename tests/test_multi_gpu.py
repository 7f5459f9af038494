"""GPU, >= 2 devices: the public API under torchrun (NCCL table broadcast + sharded subjects)
reproduces the reference's files.  Skipped on single-GPU boxes; the same host logic is covered
on CPU by test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.parametrize("case,chunk", [("g2_edges", "0"), ("g3_pop3_messy", "700"), ("g5_typed_cau", "3000"),
                                        ("g8_plan_a_blocks_planb_off", "1500")])
def test_torchrun_two_gpus_matches_golden(case, chunk, tmp_path, monkeypatch):
    """chunk: GRIMB_FILE_CHUNK -- small values cut the input into many chunks, so that both ranks stream interleaved
    pieces into the shared files (line indices and offsets through the board, grimb_impute_file_sharded)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("GRIMB_FILE_CHUNK", chunk)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300), os.path.join(HERE, "multi_gpu_worker.py"),
           case, str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIFFERENT" not in r.stdout
