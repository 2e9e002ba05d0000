#!/usr/bin/env python
"""Times small Hamming searches on each engine (tuning aid for the auto engine choice)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_snvbert_b200 import WindowedHammingIndex, _lib
import bench

dev = torch.device("cuda", 0)
S, k = 1030, 8
for (W, N, Q) in [(1, 5008, 32), (1, 5008, 64), (1, 5008, 128), (1, 5008, 256), (1, 5008, 1024), (4, 5008, 64), (16, 5008, 64),
                  (64, 5008, 64), (1, 50000, 64), (1, 2008, 48), (8, 2008, 48)]:
    panel = bench.gen_windows_device(torch, dev, 2000, W, N, S, 777)
    queries = bench.gen_windows_device(torch, dev, 5000, W, Q, S, 777)
    idx = WindowedHammingIndex(S, W, 0); idx.add(panel)
    row = {"W": W, "N": N, "Q": Q}
    ref = None
    for eng in ("popc", "tc4", "tc4x2"):
        os.environ["SNV_HAMMING_ENGINE"] = eng
        for _ in range(3): D, I = idx.search(queries, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): D, I = idx.search(queries, k)
        e1.record(); torch.cuda.synchronize()
        row[eng + "_us"] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
        if ref is None: ref = (D.clone(), I.clone())
        else: assert torch.equal(ref[0], D) and torch.equal(ref[1], I)
    print(json.dumps(row), flush=True)
