#!/usr/bin/env python
"""The reference's real V18 retrieval shape (SURVEY.md cfg 4'): N=2008 refs, nq<=48, d = L*D = 1030*192.
Accuracy vs float64 and timing vs torch.cdist+topk (the reference's own GPU path,
src/dataset/embedding_rag_dataset.py:392-402).  Tuning/validation aid."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rag_snvbert_b200 import WindowedL2Index, _lib

L, D = int(os.environ.get("L", 1030)), int(os.environ.get("D", 192))
N, nq, k = int(os.environ.get("N", 2008)), int(os.environ.get("NQ", 48)), int(os.environ.get("K", 1))
prec = os.environ.get("PREC", "tf32x3")
center = os.environ.get("CENTER", "0") == "1"
g = torch.Generator(device="cuda"); g.manual_seed(0)
dev = "cuda"
T = torch.randn(7, D, device=dev, generator=g)                      # token embedding
P = torch.randn(L, D, device=dev, generator=g) * 1.0               # position + AF part (shared by all rows)
founders = (torch.rand(16, L, device=dev, generator=g) < 0.25)
def haps(n):
    f = torch.randint(0, 16, (n,), device=dev, generator=g)
    h = founders[f].clone()
    h ^= torch.rand(n, L, device=dev, generator=g) < 0.02
    return h
mask = torch.rand(L, device=dev, generator=g) < 0.3                 # window mask shared by refs and queries
def embed(h):
    tok = torch.where(h, 6, 5)
    tok[:, mask] = 4
    return (T[tok] + P[None]).reshape(h.shape[0], L * D).contiguous()
refs = embed(haps(N)); q = embed(haps(nq))
d64 = torch.cdist(q.double(), refs.double(), p=2) ** 2
D64, I64 = d64.topk(k, largest=False, dim=1)
kw = {"center": True} if center else {}
idx = WindowedL2Index(L * D, 1, 0, prec, **kw)
t0 = time.perf_counter(); idx.add(refs); torch.cuda.synchronize(); t_add = time.perf_counter() - t0
Dg, Ig = idx.search(q, k)
torch.cuda.synchronize()
_lib.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); Dg, Ig = idx.search(q, k); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1); kms = _lib.profile_last_ms()
for _ in range(3):  # warm the reference's GPU path (cuBLAS handles, workspaces) before timing it
    dd = torch.cdist(q, refs, p=2); _, It = dd.topk(k, largest=False, dim=1)
torch.cuda.synchronize()
e0.record(); dd = torch.cdist(q, refs, p=2); _, It = dd.topk(k, largest=False, dim=1); e1.record(); torch.cuda.synchronize()
ms_torch = e0.elapsed_time(e1)
got = torch.gather(d64, 1, Ig)
print(json.dumps({"d": L * D, "N": N, "nq": nq, "k": k, "prec": prec, "center": center, "add_s": t_add, "search_ms": ms, "kernel_ms": kms,
                  "torch_cdist_topk_ms": ms_torch, "ids_equal_f64": float((Ig == I64).float().mean()),
                  "torch_ids_equal_f64": float((It == I64).float().mean()),
                  "max_abs_dD": float((Dg.double() - got).abs().max()), "true_gap_min": float((d64.topk(2, largest=False, dim=1)[0].diff(dim=1)).min()),
                  "dist_scale": float(D64.mean()), "norm_scale": float((refs.double() ** 2).sum(1).mean())}))
