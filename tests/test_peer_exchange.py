"""Multi-GPU: the moments exchange over NVLink peer memory (dist.PeerExchange) against the NCCL all-reduce.
Needs >= 2 GPUs on one node; the driver's single-GPU `pytest -m gpu` run skips it."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nproc", [2, 4])
def test_peer_exchange_equals_nccl_allreduce(nproc):
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(29530 + nproc), os.path.join(ROOT, "tests", "peer_exchange_worker.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert f"peer exchange OK {nproc}" in proc.stdout
