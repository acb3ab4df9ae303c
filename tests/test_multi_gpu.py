"""Multi-GPU exchange paths inside libpcdb200 over real NCCL (pytest -m gpu; skipped with fewer than 2 GPUs): launches
tests/run_sharded_nccl.py with one process per GPU.  The script compares, on every rank, the row-sharded codebook and
the keypoint-sharded scene against an unsharded context on the same GPU (rows, distance bits, votes, maxima, labels)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
def test_sharded_codebook_and_scene_over_nccl(tmp_path):
    n = min(_n_gpus(), 8)
    out = tmp_path / "sharded.json"
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "run_sharded_nccl.py"), "--words",
           "30000", "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(out))
    assert res["world"] == n and all(res["identical_to_unsharded"].values()), res
