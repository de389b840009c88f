"""GPU, >= 2 devices (skipped otherwise): the N > 1 product path end to end under torchrun -- dist.generate_batch with the fused
frame all-gather (csrc/vae_ops.cu::frames_to_u8_allgather_kernel through FusedFrameGather) against the NCCL path."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs on one node")
def test_generate_batch_fused_gather_torchrun():
    n = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "_mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=540, cwd=ROOT)
    assert r.returncode == 0 and f"MGPU_OK {n}" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
