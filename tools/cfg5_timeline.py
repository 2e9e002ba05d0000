#!/usr/bin/env python
"""Event timeline of the pipelined row-sharded search (cfg 5 shapes) under torchrun: per batch, when the scan and the
exchange start and end on their streams.  Diagnostic (what overlaps what); REFS = panel rows over all ranks."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from rag_snvbert_b200 import WindowedHammingIndex, _lib
from rag_snvbert_b200.sharding import RowShardedSearch, shard_range

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
N, Q, W, S, k = int(os.environ.get("REFS", 200000)), 10000, 8, 1030, 32
lo, hi = shard_range(N, world, rank)
g = torch.Generator(device="cuda"); g.manual_seed(1)
stride = _lib.packed_stride(S)
panel = torch.randint(-2**31, 2**31 - 1, (W, hi - lo, stride), device="cuda", generator=g, dtype=torch.int32)
panel[:, :, _lib.packed_words(S) - 1] &= (1 << (S % 32)) - 1 if S % 32 else -1
panel[:, :, _lib.packed_words(S):] = 0
g.manual_seed(2)
q = torch.randint(-2**31, 2**31 - 1, (W, Q, stride), device="cuda", generator=g, dtype=torch.int32)
q[:, :, _lib.packed_words(S) - 1] &= (1 << (S % 32)) - 1 if S % 32 else -1
q[:, :, _lib.packed_words(S):] = 0
idx = WindowedHammingIndex(S, W, local); idx.add(panel); del panel
s = RowShardedSearch(idx, lo, world=world)
for _ in range(6):
    s.search(q, k, sync=False)
s.wait(); torch.cuda.synchronize()
if world > 1: dist.barrier()
steps = int(os.environ.get("STEPS", 8))
main = torch.cuda.current_stream()
ev = []
orig = idx.search
def timed_search(*args, **kw):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); r = orig(*args, **kw); b.record(); ev.append((a, b)); return r
idx.search = timed_search
s.search_fn = lambda qq, kk, w0: idx.search(qq, kk, w0=w0, id_offset=lo)
t0 = torch.cuda.Event(enable_timing=True); t0.record()
side_ev = []
for i in range(steps):
    s.search(q, k, sync=False)
    e = torch.cuda.Event(enable_timing=True); e.record(s._side) if s._side is not None else e.record(); side_ev.append(e)
s.wait(); t1 = torch.cuda.Event(enable_timing=True); t1.record(); torch.cuda.synchronize()
rows = [{"scan_start": round(t0.elapsed_time(a), 3), "scan_end": round(t0.elapsed_time(b), 3), "exchange_end": round(t0.elapsed_time(e), 3)}
        for (a, b), e in zip(ev, side_ev)]
out = {"rank": rank, "world": world, "rows_per_gpu": hi - lo, "total_ms": round(t0.elapsed_time(t1), 3), "per_step": round(t0.elapsed_time(t1) / steps, 3),
       "transport": s.describe()[:60], "timeline": rows}
if rank == 0: print(json.dumps(out))
if world > 1: dist.barrier(); dist.destroy_process_group()
