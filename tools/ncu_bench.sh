#!/bin/bash
# ncu evidence for the bench command (1 GPU): the launch list (per-launch device times of the library's kernels, namespace snv::;
# the synthetic-data generation is torch kernels and is left out) and one --set full capture of the
# dominant kernel.  The same command runs without ncu first (rule: profile only what has just exited 0).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2}
CMD="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:hamming_tc_kernel|hamming_topk|pack_|merge_|narrow_|restride_|exchange_|^gather_|l2_|tc_expand|intersect_mask|invert_packed" -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_list.log 2>&1
echo "list rc=$?"
[ -n "$LIST_ONLY" ] && exit 0
$CMD > gpurun_out/${T}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hamming_tc_kernel -s 4 -c 1 -f -o gpurun_out/${T}_bench_tc_full $CMD > gpurun_out/${T}_ncu_full.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/${T}_plain.log | cut -c1-300
