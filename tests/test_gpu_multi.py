"""Data-parallel parity on real GPUs: 2 ranks (NCCL + the NVLink peer exchange of the BatchNorm sums) against the
single-GPU result of the concatenated batch -- grad / Hv / vGHv / lambda_max, ragged shards, batch-global weighted-BCE
counts, the K-FAC preconditioned variant.  Needs >= 2 visible GPUs (``gpurun --gpus 2``); skipped on one."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_minibatch_equals_the_single_gpu_batch(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "mgpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert "mgpu_check ok" in r.stdout
