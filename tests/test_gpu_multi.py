"""Multi-GPU path on real hardware: the product's resynthesize_sharded under torchrun, one
process per GPU over NCCL.  Skipped on a single-GPU box (the host logic is covered by the gloo tests in
tests/test_host_cpu.py, which run everywhere)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_sharded_resynthesis_under_torchrun(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + os.getpid() % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "sharded_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(res.stdout[-4000:])
    sys.stderr.write(res.stderr[-4000:])
    assert res.returncode == 0 and "SHARDED-OK" in res.stdout
