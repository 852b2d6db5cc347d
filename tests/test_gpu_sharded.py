"""Cell-block sharding (SURVEY 8e level 2, BASELINE config 5) on real GPUs: launches tests/sharded_worker.py
under torchrun with two ranks.  Needs >= 2 B200s on the box; skipped (not passed silently) otherwise."""

import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(900)
def test_cell_block_sharding_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("cell-block sharding needs two GPUs on the box (run under `gpurun --gpus 2`)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=850)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0
    assert r.stdout.count("SHARDED_OK") == 2
