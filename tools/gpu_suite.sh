#!/bin/bash
# full GPU parity suite + cfg2/cfg3 bench lines (1 GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 600 python bench.py > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench_cfg2.json; tail -3 gpurun_out/${T}_bench_cfg2.err
timeout 600 python bench.py --workload cfg3 --masked --no-cpu-baseline > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err; echo "bench cfg3 rc=$?"
cat gpurun_out/${T}_bench_cfg3.json; tail -3 gpurun_out/${T}_bench_cfg3.err
