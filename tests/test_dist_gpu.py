"""Hardware N-GPU correctness: spawns tests/dist_worker.py with one process per GPU (NCCL) and checks that the
replicas stay bit-identical and that the all-reduced gradient is the mean of the shard gradients.  Skipped on a 1-GPU box
(run with `gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_gpu_data_parallel_step():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs (found {n})")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DIST_OK" in r.stdout
