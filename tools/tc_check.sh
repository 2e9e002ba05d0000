#!/bin/bash
# GPU: guarded correctness + timing pass for the tensor-core Hamming engines after a kernel change.
# 1. tiny searches under a timeout (a wrong barrier protocol hangs), 2. fuzz against the popcount scan, 3. timings.
cd "$(dirname "$0")/.."
TAG=${TAG:-r2}; ENGINES=${ENGINES:-"tc4x2 tc4x2ta"}
mkdir -p gpurun_out
{
echo "== tiny"
ENGINES="$ENGINES" timeout 120 python - <<'PY'
import os
import numpy as np
from rag_snvbert_b200 import WindowedHammingIndex, _lib
rng = np.random.default_rng(0)
ok = True
for (W, N, Q, d, k) in [(1, 500, 200, 100, 8), (2, 1000, 300, 1030, 8), (3, 5008, 700, 1030, 32), (2, 700, 257, 256, 5), (1, 3000, 300, 512, 8)]:
    panel = (rng.random((W, N, d)) < 0.4).astype(np.uint8); q = (rng.random((W, Q, d)) < 0.4).astype(np.uint8)
    idx = WindowedHammingIndex(d, W); idx.add(panel)
    os.environ["SNV_HAMMING_ENGINE"] = "popc"; D0, I0 = idx.search(q, k)
    for e in os.environ["ENGINES"].split():
        os.environ["SNV_HAMMING_ENGINE"] = e; D, I = idx.search(q, k)
        same = bool((D == D0).all() and (I == I0).all()); ok &= same
        print((W, N, Q, d, k), e, "engine", _lib.last_hamming_engine(), "equal to the popcount scan:", same, flush=True)
        if not same:
            bad = np.argwhere((D != D0) | (I != I0)); print("  first mismatches", bad[:5].tolist(), D[tuple(bad[0])], D0[tuple(bad[0])], I[tuple(bad[0])], I0[tuple(bad[0])])
raise SystemExit(0 if ok else 1)
PY
rc=$?; echo "tiny rc=$rc"
if [ $rc = 0 ]; then
echo "== fuzz"
ENGINES=$(echo $ENGINES | tr ' ' ',') CASES=${CASES:-120} SEED=${SEED:-11} timeout 600 python tools/fuzz_engines.py 2>&1 | tail -8
fi
for e in $ENGINES; do
  echo "== $e cfg2"; SNV_HAMMING_ENGINE=$e W=296 timeout 60 python tools/time_hamming.py 2>&1 | tail -1 | cut -c1-330
  echo "== $e cfg2 masked"; SNV_HAMMING_ENGINE=$e MASKED=1 W=296 timeout 60 python tools/time_hamming.py 2>&1 | tail -1| cut -c1-330
  echo "== $e cfg5 shard k=32"; SNV_HAMMING_ENGINE=$e W=8 N=25000 Q=10000 K=32 timeout 60 python tools/time_hamming.py 2>&1 | tail -1| cut -c1-330
done
} > gpurun_out/${TAG}_check.txt 2>&1
cat gpurun_out/${TAG}_check.txt
