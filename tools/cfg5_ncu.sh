#!/bin/bash
# cfg5 launch lists at 1 GPU: the full 200,000-row panel and the 25,000-row shard one of 8 GPUs scans
# (which kernels make up the step besides the scan).  Plain run first, then the same command under ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2}
K="regex:hamming_tc_kernel|hamming_topk|pack_|merge_|narrow_|restride_|exchange_|^gather_|l2_|tc_expand"
for R in 200000 25000; do
  CMD="python bench.py --workload cfg5 --refs $R --steps 3 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/${T}_cfg5_${R}_plain.json 2> gpurun_out/${T}_cfg5_${R}_plain.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/${T}_cfg5_${R}_launches.csv $CMD > gpurun_out/${T}_cfg5_${R}_ncu.log 2>&1
  echo "refs=$R rc=$?"
done
