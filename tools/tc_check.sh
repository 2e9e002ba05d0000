#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_hamming_tc_gpu.py -q -m gpu 2>&1 | tail -6
for e in tc tc4 tc4x2; do
  echo "=== timing $e"; SNV_HAMMING_ENGINE=$e W=296 timeout 120 python tools/time_hamming.py 2>&1 | tail -1
done
} > gpurun_out/r2_tc_check.txt 2>&1
cat gpurun_out/r2_tc_check.txt
