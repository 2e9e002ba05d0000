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


def test_two_devices_in_one_process_hamming_and_l2():
    """One process, indexes on cuda:0 and cuda:1 (needs two GPUs): the dynamic shared-memory opt-in of the tensor-core
    kernels is per device and is set on every launch, the block cache is per device, and the results on the second device
    equal those on the first."""
    import numpy as np
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from rag_snvbert_b200 import WindowedHammingIndex, WindowedL2Index

    rng = np.random.default_rng(12)
    W, N, Q, d, k = 2, 3000, 300, 1030, 8
    panel = (rng.random((W, N, d)) < 0.3).astype(np.uint8)
    q = (rng.random((W, Q, d)) < 0.3).astype(np.uint8)
    refs = rng.standard_normal((1, 2008, 256)).astype(np.float32)
    fq = rng.standard_normal((1, 64, 256)).astype(np.float32)
    out = []
    for dev in (0, 1, 0, 1):  # alternate: create / free on one device while the other one's cache holds blocks
        h = WindowedHammingIndex(d, W, dev)
        h.add(panel)
        D, I = h.search(q, k)                      # tensor-core engine (>= 512 rows, >= 32 queries)
        l2 = WindowedL2Index(256, 1, dev)
        l2.add(refs)
        Df, If = l2.search(fq, 4)
        out.append((D.copy(), I.copy(), np.asarray(If).copy()))
        del h, l2
    for o in out[1:]:
        np.testing.assert_array_equal(o[0], out[0][0])
        np.testing.assert_array_equal(o[1], out[0][1])
        np.testing.assert_array_equal(o[2], out[0][2])
