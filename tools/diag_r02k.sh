#!/bin/bash
# round 2, GPU call 11: hot-first in pass A (survivors of boundary tiles), occupancy thresholds 2,4,8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r02k_pytest.log
SETTINGS="X=0 CHAOS_HOT_FIRST=0 CHAOS_HOT_FIRST=1 CHAOS_HOT_FIRST=2 CHAOS_EMULATE_PART=0:8 CHAOS_EMULATE_PART=0:4 CHAOS_EMULATE_PART=0:2" WORKLOADS="c2 c2ex2" STEPS=10 tools/sweep_env.sh > gpurun_out/r02k_hot.txt 2>&1
SETTINGS="CHAOS_ENGINE=2 CHAOS_ENGINE=2+CHAOS_HOT_FIRST=1" WORKLOADS="c4 c5" STEPS=10 tools/sweep_env.sh >> gpurun_out/r02k_hot.txt 2>&1
echo done
