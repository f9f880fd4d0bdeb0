"""Multi-GPU pieces of the path on a box with at least two GPUs (skipped on a one-GPU box; the driver's N > 1 bench runs
exercise them as well): the peer-memory gradient all-reduce (phx_peer_allreduce / _nvls) and the exact-global-norm mode of
row-sharded batched dopri5 solves.  Both tools print one PASS / FAIL line on rank 0."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _torchrun(script, nproc):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(REPO, "tools", script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    return res.returncode, res.stdout + res.stderr


@pytest.mark.parametrize("script", ["check_peer_allreduce.py", "check_global_norm.py"])
def test_two_gpu_tools_pass(script):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    rc, out = _torchrun(script, 2)
    assert rc == 0 and "PASS" in out and "FAIL" not in out, out[-2000:]
