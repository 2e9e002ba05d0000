#!/bin/bash
cd "$(dirname "$0")/.."
for cfg in "N=50000 Q=8192 D=256 PREC=tf32x3" "N=50000 Q=8192 D=256 PREC=tf32" "N=20000 Q=4096 D=1030 PREC=tf32x3" "N=5008 Q=4096 D=256 PREC=tf32x3"; do
  for pr in 0 1; do
    echo "== $cfg pair=$pr"; env $cfg SNV_L2_PAIR=$pr timeout 120 python tools/time_l2.py 2>&1 | tail -1
  done
done
