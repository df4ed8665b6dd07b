"""GPU test with world_size >= 2 (needs >= 2 GPUs on the box; skipped otherwise): view sharding with an NCCL broadcast of
the Gaussians and the [P,4] sum all-reduce of the densify vjp reproduce the single-GPU results (tests/mp_gpu_worker.py)."""
import os
import subprocess
import sys

import pytest
import torch

import util as U

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2])
def test_view_sharding_matches_single_gpu(world, device):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(U.ROOT, "tests", "mp_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "images bit-identical=True" in r.stdout
