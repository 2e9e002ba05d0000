#!/bin/bash
# GPU bring-up of the TMEM-operand engine (SNV_HAMMING_ENGINE=tc4x2ta, hamming_tc_kernel<KT, MODE_FP4_2CTA_TA>):
# 1. one tiny search under a timeout (a wrong barrier protocol would hang: never run it unguarded),
# 2. parity against the popcount scan on random shapes, 3. timing next to the CTA-pair engine.
# Not part of the test-suite until step 2 has passed on a B200.  Run tools/build_tmema_variants.sh (CPU) first to also
# time the ring-depth / staged hand-over variants.   Output: gpurun_out/${TAG}_tmema.txt
cd "$(dirname "$0")/.."
TAG=${TAG:-r2}
mkdir -p gpurun_out
{
echo "== tiny"
timeout 60 python - <<'PY'
import os
import numpy as np
from rag_snvbert_b200 import WindowedHammingIndex, _lib
rng = np.random.default_rng(0)
for (W, N, Q, d, k) in [(1, 500, 200, 100, 8), (2, 1000, 300, 1030, 8), (3, 5008, 700, 1030, 32)]:
    panel = (rng.random((W, N, d)) < 0.4).astype(np.uint8); q = (rng.random((W, Q, d)) < 0.4).astype(np.uint8)
    idx = WindowedHammingIndex(d, W); idx.add(panel)
    os.environ["SNV_HAMMING_ENGINE"] = "popc"; D0, I0 = idx.search(q, k)
    os.environ["SNV_HAMMING_ENGINE"] = "tc4x2ta"; D, I = idx.search(q, k)
    print((W, N, Q, d, k), "engine", _lib.last_hamming_engine(), "equal to the popcount scan:", bool((D == D0).all() and (I == I0).all()), flush=True)
PY
echo "tiny rc=$?"
echo "== fuzz"
ENGINES=tc4x2ta CASES=${CASES:-40} SEED=11 timeout 300 python tools/fuzz_engines.py 2>&1 | tail -5
for e in tc4x2 tc4x2ta; do
  echo "== $e cfg2"; SNV_HAMMING_ENGINE=$e W=296 timeout 60 python tools/time_hamming.py 2>&1 | tail -1
  echo "== $e cfg2 masked"; SNV_HAMMING_ENGINE=$e MASKED=1 W=296 timeout 60 python tools/time_hamming.py 2>&1 | tail -1
  echo "== $e cfg5 shard k=32"; SNV_HAMMING_ENGINE=$e W=8 N=25000 Q=10000 K=32 timeout 60 python tools/time_hamming.py 2>&1 | tail -1
done
# variants built by tools/build_tmema_variants.sh (if present): parity proxy = the checksums must equal the lines above
for lib in tools/variants/libsnvknn_ta_*.so; do
  [ -f "$lib" ] || continue
  echo "== $(basename $lib) tc4x2ta cfg2"; SNVKNN_LIB=$PWD/$lib SNV_HAMMING_ENGINE=tc4x2ta W=296 timeout 60 python tools/time_hamming.py 2>&1 | tail -1
  echo "== $(basename $lib) tc4x2ta cfg5 shard k=32"; SNVKNN_LIB=$PWD/$lib SNV_HAMMING_ENGINE=tc4x2ta W=8 N=25000 Q=10000 K=32 timeout 60 python tools/time_hamming.py 2>&1 | tail -1
done
} > gpurun_out/${TAG}_tmema.txt 2>&1
cat gpurun_out/${TAG}_tmema.txt
