"""Worker of tests/test_peer_exchange_multigpu.py (one process per GPU under torch.distributed.run): the row-sharded
search over the fused NVLink exchange (CUDA IPC peer memory) and over NCCL, both against the unsharded search."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_snvbert_b200 import WindowedHammingIndex  # noqa: E402
from rag_snvbert_b200.sharding import RowShardedSearch, shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rng = np.random.default_rng(5)
    W, N, Q, d = 3, 4000, 64 * world, 1030
    panel = (rng.random((W, N, d)) < 0.3).astype(np.uint8)
    for k in (8, 32):
        q = torch.from_numpy((rng.random((W, Q, d)) < 0.3).astype(np.uint8)).cuda()
        full = WindowedHammingIndex(d, W, local)
        full.add(panel)
        De, Ie = full.search(q, k)
        lo, hi = shard_range(N, world, rank)
        shard = WindowedHammingIndex(d, W, local)
        shard.add(np.ascontiguousarray(panel[:, lo:hi]))
        for transport in ("peer", "nccl"):
            s = RowShardedSearch(shard, lo, world=world, transport=transport)
            for rep in range(3):  # both receive slots, growing epochs
                q_lo, q_hi, D, I = s.search(q, k, sync=(rep != 1))
                s.wait()
                torch.cuda.synchronize()
                assert torch.equal(D, De[:, q_lo:q_hi]) and torch.equal(I, Ie[:, q_lo:q_hi]), (transport, k, rep, rank)
            if transport == "peer":
                assert s._peer is not None, "the GPUs of this box have peer access: the NVLink route must be the one that ran"
                assert "NVLink" in s.describe()
    dist.barrier()
    if rank == 0:
        print("PEER_EXCHANGE_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
