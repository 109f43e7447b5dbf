"""GPU (>= 2 devices): METIS-partitioned assembly + SpMV with NCCL halo against the serial oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nproc", [2, 4])
def test_partitioned_assembly_and_spmv_over_nccl(nproc):
    import torch

    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(29620 + nproc), os.path.join(ROOT, "tests", "nccl_worker.py")]
    proc = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-4000:]
    for r in range(nproc):
        assert f"rank {r} ok" in proc.stdout, proc.stdout[-4000:]
