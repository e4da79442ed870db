"""Real NCCL run of the row-sharded path (all-gathered negatives) against the single-process oracle on the
concatenated batch.  Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_infonce_nccl(tmp_path):
    world = min(torch.cuda.device_count(), 8)
    out = tmp_path / "res.txt"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(HERE, "dist_worker.py"), "nccl",
           str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = eval(out.read_text())
    assert len(res) == 6 and res["fused"] == "ok" and res["kmeans"] == "ok" and res["pipe"] == "ok" and res["logits_peer"] == "ok"
