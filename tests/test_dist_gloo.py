"""World-size-2 gloo run (CPU) of the multi-rank host logic: global gather order [all first ; all second],
row-block offsets, the 3-floats-per-row statistics gather of the backward, and the DDP mean-over-ranks
convention -- with the oracle standing in for the CUDA kernels.  The same worker runs with NCCL + real
kernels in tests/test_multi_gpu.py."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_infonce_host_logic_gloo(tmp_path, world):
    out = tmp_path / "res.txt"
    port = 29500 + (os.getpid() % 500) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "dist_worker.py"), "gloo",
           str(out)]
    env = dict(os.environ, OMP_NUM_THREADS="2", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = eval(out.read_text())
    assert res.pop("kmeans") == "ok"          # N4 drop-in: sharded bank, identical clustering on every rank
    assert len(res) == 2
    for loss, ref, e1, e2 in res.values():
        assert abs(loss - ref) < 1e-6 * abs(ref) and e1 < 1e-5 and e2 < 1e-5
