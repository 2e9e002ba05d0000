#!/usr/bin/env python
"""Times the cfg-4 L2 search kernel (tuning aid; run under SNVKNN_LIB=...)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rag_snvbert_b200 import WindowedL2Index, _lib
N, Q, d, k = int(os.environ.get("N", 5008)), int(os.environ.get("Q", 4096)), int(os.environ.get("D", 256)), 8
prec = os.environ.get("PREC", "tf32")
torch.manual_seed(0)
refs = torch.randn(N, d, device="cuda"); q = torch.randn(Q, d, device="cuda")
idx = WindowedL2Index(d, 1, 0, prec); idx.add(refs)
for _ in range(3): idx.search(q, k)
torch.cuda.synchronize()
_lib.profile_enable(True)
ks, ss = [], []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); idx.search(q, k); e1.record(); torch.cuda.synchronize()
    ss.append(e0.elapsed_time(e1)); ks.append(_lib.profile_last_ms())
print(json.dumps({"lib": os.path.basename(_lib.so_path()), "prec": prec, "N": N, "Q": Q, "d": d,
                  "kernel_us": float(np.median(ks)) * 1e3, "step_us": float(np.median(ss)) * 1e3}))
