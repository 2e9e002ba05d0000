#!/usr/bin/env python
"""Randomised cross-check of the Hamming engines: every tensor-core variant must return exactly what the popcount
scan returns (which the parity tests pin to the oracle) over random shapes, mask modes and k."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rag_snvbert_b200 import WindowedHammingIndex
from rag_snvbert_b200.index import pack_rows

rng = np.random.default_rng(int(os.environ.get("SEED", "1234")))
dev = torch.device("cuda", 0)
n_cases = int(os.environ.get("CASES", "120"))
bad = 0
# ENGINES="tc4x2ta" (comma separated) checks other builds / bring-up engines against the popcount scan
TC_ENGINES = tuple(e for e in os.environ.get("ENGINES", "tc,tc4,tc4x2,tc4x2ta").split(",") if e)
# panel sizes around the tile widths (240 / 256 rows; 160 for the TMEM-operand engine)
N_CHOICES = [1, 7, 100, 239, 240, 241, 479, 481, 1000, 2500, 5008, 12000]
if "tc4x2ta" in TC_ENGINES:
    N_CHOICES += [159, 160, 161, 319, 321, 800, 2561, 20000]  # 20000 rows: k > 8 items long enough for the 4-tile fold windows
for case in range(n_cases):
    W = int(rng.choice([1, 1, 2, 3, 5, 9]))
    N = int(rng.choice(N_CHOICES))
    Q = int(rng.choice([1, 2, 31, 33, 127, 128, 129, 255, 257, 600]))
    d = int(rng.choice([5, 64, 255, 256, 257, 511, 513, 1024, 1030, 1057, 2049, 3000]))
    k = int(rng.choice([1, 2, 5, 8, 9, 17, 32]))
    mode = rng.choice(["none", "window", "query"])
    dens = float(rng.choice([0.02, 0.3, 0.5, 0.9]))
    g = torch.Generator(device=dev); g.manual_seed(case)
    panel = (torch.rand((W, N, d), device=dev, generator=g) < dens).to(torch.uint8)
    if rng.random() < 0.3 and N > 4:  # duplicates -> ties
        panel[:, N // 2:] = panel[:, : N - N // 2].clone()
    q = (torch.rand((W, Q, d), device=dev, generator=g) < dens).to(torch.uint8)
    obs = None
    if mode == "window":
        obs = (torch.rand((W, d), device=dev, generator=g) < 0.6).to(torch.uint8)
    elif mode == "query":
        obs = (torch.rand((W, Q, d), device=dev, generator=g) < 0.6).to(torch.uint8)
    idx = WindowedHammingIndex(d, W, 0)
    idx.add(panel)
    out = {}
    for eng in ("popc",) + TC_ENGINES:
        os.environ["SNV_HAMMING_ENGINE"] = eng
        D, I = idx.search(q, k, observed=obs)
        out[eng] = (D.clone(), I.clone())
    for eng in TC_ENGINES:
        if not (torch.equal(out["popc"][0], out[eng][0]) and torch.equal(out["popc"][1], out[eng][1])):
            bad += 1
            print("MISMATCH", eng, dict(case=case, W=W, N=N, Q=Q, d=d, k=k, mode=str(mode), dens=dens), flush=True)
print(json.dumps({"cases": n_cases, "mismatches": bad}))
sys.exit(1 if bad else 0)
