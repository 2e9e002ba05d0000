#!/bin/bash
# ncu evidence for the bench command (1 GPU).  MODE=list: per-launch device times; MODE=full: one full-set capture
# of the dominant kernel.  The same command runs without ncu first.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r5}
CMD="python bench.py --windows ${W:-296} --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
if [ "${MODE:-list}" = list ]; then
  $CMD > gpurun_out/${T}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hamming_tc_kernel|tc_expand_queries_kernel|pack_sites_kernel|merge_keys_kernel|hamming_topk_kernel" -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_list.log 2>&1
else
  $CMD > gpurun_out/${T}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:hamming_tc_kernel -s 4 -c 1 -f -o gpurun_out/${T}_tc_full $CMD > gpurun_out/${T}_ncu_full.log 2>&1
fi
echo rc=$?; tail -2 gpurun_out/${T}_plain.log | cut -c1-300
