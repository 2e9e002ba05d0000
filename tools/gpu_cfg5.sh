#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r4}
timeout 900 python bench.py --workload cfg5 --steps 10 > gpurun_out/${T}_cfg5_1gpu.json 2> gpurun_out/${T}_cfg5_1gpu.err; echo "cfg5 rc=$?"; cat gpurun_out/${T}_cfg5_1gpu.json; tail -3 gpurun_out/${T}_cfg5_1gpu.err
SNV_HAMMING_ENGINE=popc timeout 900 python bench.py --workload cfg5 --steps 3 > gpurun_out/${T}_cfg5_1gpu_popc.json 2> gpurun_out/${T}_cfg5_1gpu_popc.err; echo "cfg5 popc rc=$?"; cat gpurun_out/${T}_cfg5_1gpu_popc.json
SNV_HAMMING_ENGINE=tc timeout 900 python bench.py --workload cfg5 --steps 10 > gpurun_out/${T}_cfg5_1gpu_fp8.json 2> gpurun_out/${T}_cfg5_1gpu_fp8.err; echo "cfg5 fp8 rc=$?"; cat gpurun_out/${T}_cfg5_1gpu_fp8.json
timeout 600 python bench.py --workload cfg4 --steps 20 > gpurun_out/${T}_cfg4.json 2> gpurun_out/${T}_cfg4.err; echo "cfg4 rc=$?"; cat gpurun_out/${T}_cfg4.json
