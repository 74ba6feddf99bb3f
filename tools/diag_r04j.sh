#!/bin/bash
# GPU call 63: one-launch frames rendered in parts when they go to host memory, each part composed while the next is rendered (CHAOS_HOST_PARTS)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SETTINGS="CHAOS_HOST_PARTS=1 CHAOS_HOST_PARTS=4 CHAOS_HOST_PARTS=2 CHAOS_HOST_PARTS=8" WORKLOADS="c5 c4 c1" STEPS=10 tools/sweep_env.sh 2>&1 | tee gpurun_out/r04j.txt
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "full_size or matches_golden or device_output" 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r04j_pytest.log
