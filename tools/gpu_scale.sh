#!/bin/bash
# multi-GPU bench lines (one box): cfg2 window-sharded (weak scaling) and cfg5 row-sharded + NCCL merge (strong scaling)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r5}
run() {  # n, outfile, extra args
  local n=$1 out=$2; shift 2
  if [ "$n" = 1 ]; then timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; fi
  echo "$out rc=$?"; tail -1 gpurun_out/$out.json | cut -c1-220
}
for n in ${NS:-2 8}; do
  run $n ${T}_cfg2_${n}gpu --no-cpu-baseline --steps 30
done
for n in ${NS5:-1 2 8}; do
  run $n ${T}_cfg5_${n}gpu --workload cfg5 --windows 8 --steps 10
done
