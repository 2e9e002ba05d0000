#!/bin/bash
# Diagnostic builds of the tensor-core Hamming kernel: tools/variants/libsnvknn_tc_<name>.so (select with SNVKNN_LIB)
set -e
cd "$(dirname "$0")/../rag_snvbert_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../tools/variants
build() {  # name, flags
  local name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --threads 2 \
       "$@" -c hamming_tc.cu -o build/hamming_tc_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../tools/variants/libsnvknn_tc_$name.so \
       build/api.o build/hamming.o build/hamming_tc_$name.o build/misc_kernels.o build/l2_tcgen05.o
}
build noepi -DTC_DEBUG_NO_EPI &
build nofold -DTC_DEBUG_NO_FOLD &
build nosts -DTC_DEBUG_NO_EXPAND_STS &
build nomma -DTC_DEBUG_NO_MMA &
wait
build noepi_nosts -DTC_DEBUG_NO_EPI -DTC_DEBUG_NO_EXPAND_STS &
build noepi_nomma -DTC_DEBUG_NO_EPI -DTC_DEBUG_NO_MMA &
build nosts_nomma -DTC_DEBUG_NO_EXPAND_STS -DTC_DEBUG_NO_MMA &
build epi16 -DSNV_TC_EPI16=1 &
wait
ls ../../tools/variants | grep tc_
