#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for e in tc4; do
  echo "== base $e"; SNV_HAMMING_ENGINE=$e W=296 timeout 120 python tools/time_hamming.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['kernel_ms'], d['ms'])"
  for v in a3r6b3 a3r4b3 a2r8b3 a4r8b2 a2r4b4 a3r2b4 rawbulk rawbulk_a3r6; do
    echo "== $v $e"; SNVKNN_LIB=$PWD/tools/variants/libsnvknn_tc_$v.so SNV_HAMMING_ENGINE=$e W=296 timeout 120 python tools/time_hamming.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['kernel_ms'], d['ms'])"
  done
done
} > gpurun_out/r2_tc_variants.txt 2>&1
cat gpurun_out/r2_tc_variants.txt
