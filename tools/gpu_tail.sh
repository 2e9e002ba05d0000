#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hamming_tc_gpu.py -q -m gpu 2>&1 | tail -4
for v in "" 1; do
  echo "== cfg5 W=8 no_tail_split=${v:-0}"
  SNV_TC_NO_TAIL_SPLIT=$v timeout 300 python bench.py --workload cfg5 --windows 8 --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['checksum'])"
done
