#!/bin/bash
# ncu capture of the tensor-core Hamming kernel (one GPU; plain run first)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export SNV_HAMMING_ENGINE=${ENGINE:-tc} W=${W:-148}
python tools/time_hamming.py > gpurun_out/r2_tc_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hamming_tc_kernel -s 3 -c 1 -f \
    -o gpurun_out/r2_tc_full python tools/time_hamming.py > gpurun_out/r2_tc_ncu.log 2>&1
cat gpurun_out/r2_tc_plain.log; tail -5 gpurun_out/r2_tc_ncu.log
