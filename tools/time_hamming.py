#!/usr/bin/env python
"""Times the cfg-2 Hamming scan on a smaller window count (tuning aid; run under SNVKNN_LIB=...)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_snvbert_b200 import WindowedHammingIndex, _lib
import bench

W = int(os.environ.get("W", "296")); N, S, Q, k = (int(os.environ.get(n, d)) for n, d in (("N", 5008), ("S", 1030), ("Q", 2000), ("K", 8)))
dev = torch.device("cuda", 0)
panel = bench.gen_windows_device(torch, dev, 2000, W, N, S, 777)
queries = bench.gen_windows_device(torch, dev, 5000, W, Q, S, 777)
masks = bench.gen_masks_device(torch, dev, 8000, W, Q, S) if os.environ.get("MASKED") else None
idx = WindowedHammingIndex(S, W, 0); idx.add(panel)
for _ in range(3): D, I = idx.search(queries, k, observed=masks)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps): D, I = idx.search(queries, k, observed=masks)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
_lib.profile_enable(True)
D, I = idx.search(queries, k, observed=masks)
torch.cuda.synchronize()
kernel_ms = _lib.profile_last_ms()
_lib.profile_enable(False)
print(json.dumps({"lib": os.path.basename(_lib.so_path()), "env": {k: v for k, v in os.environ.items() if k.startswith("SNV_")},
                  "W": W, "N": N, "Q": Q, "k": k, "ms": ms, "kernel_ms": kernel_ms, "pairs_per_s": W * N * Q / ms * 1e3, "checksum": int(I.sum().item()), "dsum": int(D.sum().item())}))
