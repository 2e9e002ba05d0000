#!/bin/bash
# Tuning builds of the Hamming kernel (NW=33 only): tools/variants/libsnvknn_<name>.so
set -e
cd "$(dirname "$0")/../rag_snvbert_b200/csrc"
make -j8 >/dev/null
build() {  # name, flags
  local name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --threads 2 \
       -DSNV_TUNE_ONLY33 "$@" -c hamming.cu -o build/hamming_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../tools/variants/libsnvknn_$name.so \
       build/api.o build/hamming_$name.o build/misc_kernels.o build/l2_tcgen05.o
}
buildl2() {  # name, flags
  local name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --threads 2 \
       "$@" -c l2_tcgen05.cu -o build/l2_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../tools/variants/libsnvknn_l2_$name.so \
       build/api.o build/hamming.o build/misc_kernels.o build/l2_$name.o
}
buildl2 base &
buildl2 noepi -DL2_DEBUG_NO_EPILOGUE &
buildl2 nomma -DL2_DEBUG_NO_MMA &
buildl2 notma -DL2_DEBUG_NO_TMA &
wait
buildl2 notma_noepi -DL2_DEBUG_NO_TMA -DL2_DEBUG_NO_EPILOGUE &
buildl2 nomma_noepi -DL2_DEBUG_NO_MMA -DL2_DEBUG_NO_EPILOGUE &
wait
ls -la ../../tools/variants
