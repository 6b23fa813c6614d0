"""Data parallel over NCCL on real GPUs: the DP result equals the single-GPU result row for row on every rank (also with an uneven
split and with a rank that holds no request), with the token exchange captured inside the decode CUDA graphs.
Skips itself on a box with one GPU; world_size-2 gloo covers the host logic on CPU (tests/test_dp_gloo.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4])
def test_nccl_dp_equals_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert "DP == single-GPU on every rank: True" in r.stdout
    assert "in-graph exchange: True" in r.stdout
