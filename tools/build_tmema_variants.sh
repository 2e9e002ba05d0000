#!/bin/bash
# CPU: builds tools/variants/libsnvknn_ta_*.so, variants of the TMEM-operand engine (hamming_tc_kernel<KT, MODE_FP4_2CTA_TA>):
# ring depths and the staged hand-over of the query tile.  Time / check them on the GPU with tools/tc_tmema_bringup.sh.
set -e
cd "$(dirname "$0")/../rag_snvbert_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../tools/variants
build() {  # name, flags
  local name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --threads 2 \
       "$@" -c hamming_tc.cu -o build/hamming_tc_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../tools/variants/libsnvknn_ta_$name.so \
       build/api.o build/hamming.o build/hamming_tc_$name.o build/misc_kernels.o build/l2_tcgen05.o
}
build stagea -DSNV_TC_TA_STAGE_A=1 &
build b3r4 -DSNV_TC_TA_BSTAGES=3 -DSNV_TC_TA_RAWSTAGES=4 &
build b4r6 -DSNV_TC_TA_BSTAGES=4 -DSNV_TC_TA_RAWSTAGES=6 &
build b8r8 -DSNV_TC_TA_BSTAGES=8 -DSNV_TC_TA_RAWSTAGES=8 &
wait
# diagnostic builds (wrong results by design): what is left without the epilogue / without the fold
build noepi -DTC_DEBUG_NO_EPI &
build nofold -DTC_DEBUG_NO_FOLD &
build stagea_noepi -DSNV_TC_TA_STAGE_A=1 -DTC_DEBUG_NO_EPI &
wait
ls -la ../../tools/variants | grep ta_
