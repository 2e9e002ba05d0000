#!/bin/bash
# first GPU run of the tensor-core Hamming engine: parity tests (each under its own timeout), then timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
for e in tc4; do
  echo "=== quick parity, engine $e"
  timeout 120 python -m pytest tests/test_hamming_tc_gpu.py -x -q -m gpu -k "launches_tensor_core and $e" 2>&1 | tail -15
done
echo "=== full tc parity file"
timeout 900 python -m pytest tests/test_hamming_tc_gpu.py -q -m gpu 2>&1 | tail -40
for e in tc tc4; do
  echo "=== timing engine $e"
  SNV_HAMMING_ENGINE=$e W=296 timeout 300 python tools/time_hamming.py 2>&1 | tail -3
  SNV_HAMMING_ENGINE=$e W=296 MASKED=1 timeout 300 python tools/time_hamming.py 2>&1 | tail -3
done
} > gpurun_out/r2_tc_bringup.txt 2>&1
tail -60 gpurun_out/r2_tc_bringup.txt
