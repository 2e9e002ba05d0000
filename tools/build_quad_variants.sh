#!/bin/bash
# CPU: builds of hamming_tc.cu with / without the quad scan (tools/variants/libsnvknn_qd_*.so); time with
# PREFIX=qd_ CFG5=1 ENGINES=tc4x2ta tools/time_variants.sh
set -e
cd "$(dirname "$0")/../rag_snvbert_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../tools/variants
build() {  # name, flags
  local name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --threads 2 -diag-suppress 177 \
       "$@" -c hamming_tc.cu -o build/hamming_tc_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../tools/variants/libsnvknn_qd_$name.so \
       build/api.o build/hamming.o build/hamming_tc_$name.o build/misc_kernels.o build/l2_tcgen05.o
}
build off -DSNV_TC_QUAD_SCAN=0 -DSNV_TC_QUAD_SCAN_K32=0 &
build k8 -DSNV_TC_QUAD_SCAN=1 -DSNV_TC_QUAD_SCAN_K32=0 &
build both -DSNV_TC_QUAD_SCAN=1 -DSNV_TC_QUAD_SCAN_K32=1 &
wait
ls ../../tools/variants | grep qd_
