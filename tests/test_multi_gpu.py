"""Multi-rank GPU parity ("Multi-GPU test"): the gathered output of a clip sharded over the GPUs
of the box must be byte-identical to the one-GPU run.  Needs >= 2 GPUs (skipped on a single-GPU box); the check itself
is tools/mgpu_check.py, launched with torchrun, one rank per GPU."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_clip_is_byte_identical_to_one_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["world"] == world and res["byte_identical_to_one_gpu_on_every_rank"], res
    assert res["cfg3_geometry"]["kernel"] == "general2" and res["cfg2_geometry"]["kernel"].startswith("p1")
