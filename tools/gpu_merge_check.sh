#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hamming_gpu.py -q -m gpu -k "merge" 2>&1 | tail -5
NS="" NS5="1 2" TAG=r6 bash tools/gpu_scale.sh
