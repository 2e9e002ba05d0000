#!/bin/bash
# 2-GPU check of the fused NVLink exchange: parity worker, then cfg5 with both transports (checksums must agree).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2}
N=${NGPU:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/peer_worker.py 2>&1 | tail -15
for X in peer nccl; do
  SNV_EXCHANGE=$X timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus $N --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_cfg5_${N}gpu_$X.json 2> gpurun_out/${T}_cfg5_${N}gpu_$X.err
  echo "$X rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${T}_cfg5_${N}gpu_$X.json") if l.startswith("{")][0])
    print("$X", d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["checksum"], d["config"]["exchange"][:90])
except Exception as e:
    print("no line", e); print(open("gpurun_out/${T}_cfg5_${N}gpu_$X.err").read()[-3000:])
PY
done
