#!/bin/bash
# Round-end check: full GPU suite on the committed build, then parity proxy + timing of the SNV_TC_EPI_MIN variants.
cd "$(dirname "$0")/.."
TAG=${TAG:-r22}
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
{
for v in "" min4 min2; do
  lib=""; [ -n "$v" ] && lib=$PWD/tools/variants/libsnvknn_tc_$v.so
  echo "== ${v:-base} tc4x2"; SNVKNN_LIB=$lib SNV_HAMMING_ENGINE=tc4x2 W=296 timeout 60 python tools/time_hamming.py 2>&1 | tail -1
done
for v in min4 min2; do
  echo "== $v fuzz"; SNVKNN_LIB=$PWD/tools/variants/libsnvknn_tc_$v.so CASES=12 SEED=5 timeout 90 python tools/fuzz_engines.py 2>&1 | tail -1
done
echo "== min4 tc4x2 cfg5 k=32"; SNVKNN_LIB=$PWD/tools/variants/libsnvknn_tc_min4.so SNV_HAMMING_ENGINE=tc4x2 W=8 N=25000 Q=10000 K=32 timeout 60 python tools/time_hamming.py 2>&1 | tail -1
echo "== base tc4x2 cfg5 k=32"; SNV_HAMMING_ENGINE=tc4x2 W=8 N=25000 Q=10000 K=32 timeout 60 python tools/time_hamming.py 2>&1 | tail -1
} > gpurun_out/${TAG}_tc_min_variants.txt 2>&1
cat gpurun_out/${TAG}_tc_min_variants.txt
