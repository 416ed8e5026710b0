"""bvc_b200.DistributedDataParallel with the REAL CUDA model on two GPUs over NCCL (tools/ddp_parity.py holds the
checks: gradients bitwise identical across ranks and equal to torch DDP's and to the mean of the per-rank gradients,
fwd-fwd-bwd-bwd, no_sync, parameter equality after optimizer steps, and BASELINE.json config 3's NT-Xent over
embeddings gathered across GPUs).  Skipped on boxes with fewer than two GPUs; the committed log of a 2 x B200 run is
profiles/r02_ddp_parity_2gpu.log."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_bvc_ddp_parity_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29655", os.path.join(ROOT, "tools", "ddp_parity.py"), "--config", "base",
           "--batch", "8"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith(("PASS", "FAIL", "DDP-PARITY"))]
    fails = [ln for ln in lines if ln.startswith("FAIL")]
    assert r.returncode == 0 and not fails and any("ALL PASS" in ln for ln in lines), \
        "\n".join(fails or lines[-8:]) + "\n" + r.stderr[-3000:]
