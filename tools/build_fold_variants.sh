#!/bin/bash
# CPU: builds of hamming_tc.cu with different deferred-fold policies (tools/variants/libsnvknn_fd_*.so); time with
# PREFIX=fd_ CFG5=1 ENGINES=tc4x2ta tools/time_variants.sh
set -e
cd "$(dirname "$0")/../rag_snvbert_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../tools/variants
build() {  # name, flags
  local name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --threads 2 -diag-suppress 177 \
       "$@" -c hamming_tc.cu -o build/hamming_tc_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../tools/variants/libsnvknn_fd_$name.so \
       build/api.o build/hamming.o build/hamming_tc_$name.o build/misc_kernels.o build/l2_tcgen05.o
}
build off -DSNV_TC_DEFER_FOLD=0 &
build w4_8 -DSNV_TC_FOLD_WARM=4 -DSNV_TC_FOLD_WINDOW=8 &
build w16_8 -DSNV_TC_FOLD_WARM=16 -DSNV_TC_FOLD_WINDOW=8 &
build w8_4 -DSNV_TC_FOLD_WARM=8 -DSNV_TC_FOLD_WINDOW=4 &
wait
build w8_2 -DSNV_TC_FOLD_WARM=8 -DSNV_TC_FOLD_WINDOW=2 &
build w16_4 -DSNV_TC_FOLD_WARM=16 -DSNV_TC_FOLD_WINDOW=4 &
build w32_8 -DSNV_TC_FOLD_WARM=32 -DSNV_TC_FOLD_WINDOW=8 &
build w2_8 -DSNV_TC_FOLD_WARM=2 -DSNV_TC_FOLD_WINDOW=8 &
wait
ls ../../tools/variants | grep fd_
