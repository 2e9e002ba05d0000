"""Row-sharded search over the fused NVLink exchange with REAL peers: one process per GPU, CUDA IPC handles all-gathered
through torch.distributed.  Needs two GPUs; on a one-GPU box the single-process variant in test_hamming_gpu.py
(test_peer_exchange_fused_kernel_single_process) covers the kernel."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_row_sharded_search_over_nvlink_peer_memory_two_processes():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "peer_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PEER_EXCHANGE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
