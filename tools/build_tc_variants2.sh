#!/bin/bash
# Ring-depth variants of the tensor-core Hamming kernel
set -e
cd "$(dirname "$0")/../rag_snvbert_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../tools/variants
rm -f ../../tools/variants/libsnvknn_tc_*.so
build() {  # name, flags
  local name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --threads 2 \
       "$@" -c hamming_tc.cu -o build/hamming_tc_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../tools/variants/libsnvknn_tc_$name.so \
       build/api.o build/hamming.o build/hamming_tc_$name.o build/misc_kernels.o build/l2_tcgen05.o
}
build a3r6b3 -DSNV_TC_ASTAGES=3 -DSNV_TC_RAWSTAGES=6 &
build a3r4b3 -DSNV_TC_ASTAGES=3 -DSNV_TC_RAWSTAGES=4 &
build a2r8b3 -DSNV_TC_ASTAGES=2 -DSNV_TC_RAWSTAGES=8 &
build a4r8b2 -DSNV_TC_ASTAGES=4 -DSNV_TC_RAWSTAGES=8 -DSNV_TC_BSTAGES=2 &
wait
build a2r4b4 -DSNV_TC_ASTAGES=2 -DSNV_TC_RAWSTAGES=4 -DSNV_TC_BSTAGES=4 &
build a3r2b4 -DSNV_TC_ASTAGES=3 -DSNV_TC_RAWSTAGES=2 -DSNV_TC_BSTAGES=4 &


wait
ls ../../tools/variants | grep tc_
