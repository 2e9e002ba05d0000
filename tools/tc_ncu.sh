#!/bin/bash
# ncu --set full capture (with source counters) of the tensor-core Hamming kernel, one GPU; the same command runs without ncu first.
# ENGINE=tc4x2|tc4x2ta|...  W=windows  TAG=output prefix (gpurun_out/${TAG}_tc_full.ncu-rep)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2}
export SNV_HAMMING_ENGINE=${ENGINE:-tc4x2} W=${W:-148}
python tools/time_hamming.py > gpurun_out/${T}_tc_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hamming_tc_kernel -s 3 -c 1 -f \
    -o gpurun_out/${T}_tc_full python tools/time_hamming.py > gpurun_out/${T}_tc_ncu.log 2>&1
echo rc=$?; cat gpurun_out/${T}_tc_plain.log | cut -c1-300; tail -3 gpurun_out/${T}_tc_ncu.log
