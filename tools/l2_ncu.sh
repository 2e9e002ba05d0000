#!/bin/bash
# ncu --set full capture (with source counters) of the float-L2 kernel inside the cfg-4 bench command (1 GPU; plain run first)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2}
CMD="python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${T}_l2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:l2_topk_kernel -s 4 -c 1 -f -o gpurun_out/${T}_l2_full $CMD > gpurun_out/${T}_l2_ncu.log 2>&1
echo rc=$?; tail -2 gpurun_out/${T}_l2_ncu.log
