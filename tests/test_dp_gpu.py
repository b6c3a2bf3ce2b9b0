"""Data parallelism on real GPUs: N ranks x B images == 1 rank x N*B images (tests/dp_equivalence_gpu.py),
spawned under torchrun when the box has at least two GPUs; skipped on a single-GPU box."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("script", ["dp_equivalence_gpu.py"])
def test_two_rank_step_equals_single_rank_at_global_batch(script):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", script)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-4000:]
    assert res.stdout.count("dp equivalence ok") == 2, res.stdout[-4000:]       # eager and CUDA-graph replay
