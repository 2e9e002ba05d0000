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
build g2_adds  -DSNV_POPC_MODE=0 -DSNV_ROW_GROUP=2 &
build g2_mad1  -DSNV_POPC_MODE=1 -DSNV_ROW_GROUP=2 &
build g2_mad4  -DSNV_POPC_MODE=2 -DSNV_ROW_GROUP=2 &
build g4_mad1  -DSNV_POPC_MODE=1 -DSNV_ROW_GROUP=4 &
wait
build g4_adds  -DSNV_POPC_MODE=0 -DSNV_ROW_GROUP=4 &
build g3_mad1  -DSNV_POPC_MODE=1 -DSNV_ROW_GROUP=3 &
build g1_mad1  -DSNV_POPC_MODE=1 -DSNV_ROW_GROUP=1 &
build g2_mad1_b256 -DSNV_POPC_MODE=1 -DSNV_ROW_GROUP=2 -DSNV_BLOCK=256 &
wait
ls -la ../../tools/variants
