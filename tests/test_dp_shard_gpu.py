"""Row-sharded data-parallel optimizer == replicated data parallel, bit for bit (needs 2 GPUs; skipped on a 1-GPU box).
Runs tools/dp_shard_check.py under torchrun: three stage-3 steps with lr = 1e-3 in both modes on identical models and
batches, 96 tensors compared (every parameter, the fc1 bf16 shadow, consolidated fp32 master and Adam moments)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_dp_matches_replicated_dp_two_ranks():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577",
                          os.path.join(ROOT, "tools", "dp_shard_check.py")], capture_output=True, text=True, env=env,
                         timeout=900)
    assert "DP_SHARD_CHECK PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
