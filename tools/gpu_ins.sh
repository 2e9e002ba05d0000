#!/bin/bash
cd "$(dirname "$0")/.."
for v in 0.02 0.1 0.2 0.4; do
  echo "== ins $v"; SNV_TC_INS=$v timeout 300 python bench.py --workload cfg5 --windows 8 --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'])"
done
