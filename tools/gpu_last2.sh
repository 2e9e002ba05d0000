#!/bin/bash
# Share of the fold / of the whole epilogue in the CTA-pair kernel (diagnostic builds; results are wrong by design)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for v in nofold noepi; do
  echo "== $v tc4x2"; SNVKNN_LIB=$PWD/tools/variants/libsnvknn_tc_$v.so SNV_HAMMING_ENGINE=tc4x2 W=296 timeout 40 python tools/time_hamming.py 2>&1 | tail -1
done
echo "== nofold tc4x2 cfg5 k=32"; SNVKNN_LIB=$PWD/tools/variants/libsnvknn_tc_nofold.so SNV_HAMMING_ENGINE=tc4x2 W=8 N=25000 Q=10000 K=32 timeout 40 python tools/time_hamming.py 2>&1 | tail -1
} > gpurun_out/r23_pair_breakdown.txt 2>&1
cat gpurun_out/r23_pair_breakdown.txt
