#!/usr/bin/env python
"""cfg 1 (one faiss IndexFlatL2 per window from host arrays): where a build + search call spends its time."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rag_snvbert_b200 import faiss_compat as faiss
rng = np.random.default_rng(0)
panel = (rng.random((5008, 1030)) < 0.3).astype(np.float32); q = (rng.random((1000, 1030)) < 0.3).astype(np.float32)
out = []
for rep in range(6):
    t = [time.perf_counter()]
    index = faiss.IndexFlatL2(1030); torch.cuda.synchronize(); t.append(time.perf_counter())
    index.add(panel); torch.cuda.synchronize(); t.append(time.perf_counter())
    D, I = index.search(q, 1); torch.cuda.synchronize(); t.append(time.perf_counter())
    D, I = index.search(q, 1); torch.cuda.synchronize(); t.append(time.perf_counter())
    del index; torch.cuda.synchronize(); t.append(time.perf_counter())
    out.append([round((b - a) * 1e3, 3) for a, b in zip(t, t[1:])])
print(json.dumps({"columns": ["create", "add", "search#1", "search#2", "free"], "ms": out, "cores": os.cpu_count()}))
