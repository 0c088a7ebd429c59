"""GPU, world_size 2: the NCCL-free DDP path against the oracle (needs >= 2 GPUs)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("model,steps,overlap,prec", [("net", 10, "1", "fp32"), ("net", 6, "0", "fp32"),
                                                      ("resnet18", 4, "1", "fp32"), ("resnet18", 6, "1", "bf16")])
def test_ddp_two_ranks_matches_oracle(model, steps, overlap, prec):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "ddp_check.py"),
           model, str(steps), overlap, prec]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DDP_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
