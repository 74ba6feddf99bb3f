#!/bin/bash
# round 2, GPU call 18: scaled form on the real axis: parity at full size, the slow rank of 8, N = 1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_bench_constants_gpu.py -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r02r_pytest.log
SETTINGS="X=0 CHAOS_EMULATE_PART=1:8 CHAOS_EMULATE_PART=0:8 CHAOS_EMULATE_PART=1:4 CHAOS_EMULATE_PART=0:4 CHAOS_EMULATE_PART=1:2 CHAOS_EMULATE_PART=0:2" WORKLOADS="c2" STEPS=20 tools/sweep_env.sh > gpurun_out/r02r_axis.txt 2>&1
SETTINGS="X=0" WORKLOADS="c2f32 c5 c4" STEPS=10 tools/sweep_env.sh >> gpurun_out/r02r_axis.txt 2>&1
echo done
