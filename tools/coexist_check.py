#!/usr/bin/env python
"""Does the fused exchange kernel (128-thread blocks, <= 32 registers) run NEXT TO a resident scan CTA?  One process, one
GPU: start a 1.3 ms scan, then the exchanges of two pointer-wired peers on two other streams; if they end long before
the scan does, they were co-resident."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_snvbert_b200 import WindowedHammingIndex, _lib
from rag_snvbert_b200.sharding import PeerExchange

N, Q, W, S, k = 25000, 10000, 8, 1030, 32
g = torch.Generator(device="cuda"); g.manual_seed(1)
stride = _lib.packed_stride(S)
def rows(n):
    x = torch.randint(-2**31, 2**31 - 1, (W, n, stride), device="cuda", generator=g, dtype=torch.int32)
    x[:, :, _lib.packed_words(S) - 1] &= (1 << (S % 32)) - 1
    x[:, :, _lib.packed_words(S):] = 0
    return x
idx = WindowedHammingIndex(S, W, 0); idx.add(rows(N)); q = rows(Q)
D, I = idx.search(q, k)
world = 2
peers = PeerExchange.connect_local(0, world, W * Q * k * 8)
streams = [torch.cuda.Stream() for _ in range(world)]
outs = [(torch.empty((W, Q // world, k), dtype=torch.int32, device="cuda"), torch.empty((W, Q // world, k), dtype=torch.int64, device="cuda")) for _ in range(world)]
res = []
for mode in ("alone", "with_scan", "with_scan"):
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t0.record()
    scan_end = torch.cuda.Event(enable_timing=True)
    if mode == "with_scan":
        idx.search(q, k, out=(D, I)) if False else idx.search(q, k)
        scan_end.record()
    ends = []
    for r in range(world):
        streams[r].wait_event(t0)
        with torch.cuda.stream(streams[r]):
            peers[r].exchange(D, I, k, out=outs[r])
            e = torch.cuda.Event(enable_timing=True); e.record(); ends.append(e)
    torch.cuda.synchronize()
    res.append({"mode": mode, "scan_end_ms": round(t0.elapsed_time(scan_end), 3) if mode == "with_scan" else None,
                "exchange_end_ms": [round(t0.elapsed_time(e), 3) for e in ends]})
print(json.dumps(res))
