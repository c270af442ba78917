"""Multi-GPU parity (needs >= 2 GPUs on the box): launches tests/dist_check.py under torchrun, one rank per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_partitioned_parity(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "dist_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(p.stdout[-4000:])
    sys.stderr.write(p.stderr[-4000:])
    assert p.returncode == 0 and "DIST_CHECK OK" in p.stdout
