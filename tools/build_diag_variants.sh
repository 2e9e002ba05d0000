#!/bin/bash
# CPU: diagnostic builds of hamming_tc.cu (wrong results by design) that remove one role's work at a time, to find what
# bounds the tensor-core Hamming kernels.  Output: tools/variants/libsnvknn_dg_*.so; time with tools/time_variants.sh.
set -e
cd "$(dirname "$0")/../rag_snvbert_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../tools/variants
build() {  # name, flags
  local name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --threads 2 -diag-suppress 177 \
       "$@" -c hamming_tc.cu -o build/hamming_tc_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../tools/variants/libsnvknn_dg_$name.so \
       build/api.o build/hamming.o build/hamming_tc_$name.o build/misc_kernels.o build/l2_tcgen05.o
}
build e -DTC_DEBUG_NO_EPI &
build em -DTC_DEBUG_NO_EPI -DTC_DEBUG_NO_MMA &
build es -DTC_DEBUG_NO_EPI -DTC_DEBUG_NO_EXPAND_STS &
build er -DTC_DEBUG_NO_EPI -DTC_DEBUG_NO_RAW &
wait
build ems -DTC_DEBUG_NO_EPI -DTC_DEBUG_NO_MMA -DTC_DEBUG_NO_EXPAND_STS &
build emsr -DTC_DEBUG_NO_EPI -DTC_DEBUG_NO_MMA -DTC_DEBUG_NO_EXPAND_STS -DTC_DEBUG_NO_RAW &
build esr -DTC_DEBUG_NO_EPI -DTC_DEBUG_NO_EXPAND_STS -DTC_DEBUG_NO_RAW &
build m -DTC_DEBUG_NO_MMA &
wait
ls ../../tools/variants | grep dg_
