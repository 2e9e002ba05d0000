#!/bin/bash
cd "$(dirname "$0")/.."
echo "== base"; SNV_HAMMING_ENGINE=tc4x2 W=296 timeout 120 python tools/time_hamming.py 2>&1 | tail -1
echo "== epi16 (G=32)"; SNVKNN_LIB=$PWD/tools/variants/libsnvknn_tc_epi16.so SNV_HAMMING_ENGINE=tc4x2 W=296 timeout 120 python tools/time_hamming.py 2>&1 | tail -1
SNVKNN_LIB=$PWD/tools/variants/libsnvknn_tc_epi16.so timeout 300 python -m pytest tests/test_hamming_tc_gpu.py -q -m gpu -k "tc4x2" 2>&1 | tail -2
