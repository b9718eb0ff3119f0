"""Multi-GPU (one process per GPU, NCCL) check of the fused gather: the forward kernel's epilogue stores O straight
into the destination rank's symmetric buffer over NVLink.  Needs >= 2 GPUs on the box (skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_store_gather_matches_nccl_gather_on_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tools", "check_peer_gather.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PEER GATHER OK" in r.stdout, r.stdout[-2000:]


@pytest.mark.gpu
def test_sequence_parallel_single_prompt_matches_single_gpu_on_two_gpus():
    """UlyssesLiteAttention: all_to_all in, O scattered to the owning ranks by the forward epilogue (peer stores)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_ulysses.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ULYSSES OK" in r.stdout, r.stdout[-2000:]
