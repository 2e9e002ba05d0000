#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "=== quick parity tc4x2"
timeout 90 python -m pytest tests/test_hamming_tc_gpu.py -x -q -m gpu -k "launches_tensor_core and tc4x2" 2>&1 | tail -15
rc=${PIPESTATUS[0]}
if [ "$rc" = 0 ]; then
  echo "=== full parity tc4x2"
  timeout 600 python -m pytest tests/test_hamming_tc_gpu.py -q -m gpu -k "tc4x2 or engines_agree" 2>&1 | tail -15
  for e in tc4 tc4x2; do
    echo "=== timing $e"; SNV_HAMMING_ENGINE=$e W=296 timeout 120 python tools/time_hamming.py 2>&1 | tail -1
  done
fi
} > gpurun_out/r2_tc_2cta.txt 2>&1
tail -40 gpurun_out/r2_tc_2cta.txt
