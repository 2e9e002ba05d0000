#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "pipeline or many_windows or engines_agree" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r10_bench_cfg2.json 2> gpurun_out/r10_bench_cfg2.err; echo rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r10_bench_cfg2.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'])"
