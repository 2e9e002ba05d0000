#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_l2_gpu.py -x -q -m gpu -k "cfg4 or golden" 2>&1 | tail -4
timeout 600 python -m pytest tests/test_l2_gpu.py tests/test_refdb_gpu.py -q -m gpu 2>&1 | tail -4
for pr in 0 1; do
  SNV_L2_PAIR=$pr timeout 300 python bench.py --workload cfg4 --steps 50 --no-cpu-baseline > gpurun_out/r14_cfg4_pair$pr.json 2> gpurun_out/r14_cfg4_pair$pr.err; echo "cfg4 pair=$pr rc=$?"
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/r14_cfg4_pair$pr.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline'].get('issued_frac'), d['value'])"
done
