#!/bin/bash
# round 2, GPU call 10: parity after the classify/export-all/occupancy changes; occupancy thresholds at N = 1 and for one rank of 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_api_gpu.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02j_pytest.log
SETTINGS="X=0 CHAOS_LONG_OCC=0,0,0 CHAOS_LONG_OCC=2,4,8 CHAOS_LONG_OCC=8,16,32 CHAOS_LONG_OCC=4,8,1000 CHAOS_LONG_OCC=16,32,64" WORKLOADS="c2 c2ex2" STEPS=10 tools/sweep_env.sh > gpurun_out/r02j_occ.txt 2>&1
SETTINGS="CHAOS_EMULATE_PART=0:8 CHAOS_EMULATE_PART=0:8+CHAOS_LONG_OCC=0,0,0 CHAOS_EMULATE_PART=0:8+CHAOS_LONG_OCC=8,16,32 CHAOS_EMULATE_PART=0:8+CHAOS_LONG_OCC=16,32,64 CHAOS_EMULATE_PART=0:8+CHAOS_STRANDS=1 CHAOS_EMULATE_PART=0:8+CHAOS_STRANDS=1+CHAOS_LONG_OCC=16,32,64 CHAOS_EMULATE_PART=0:4 CHAOS_EMULATE_PART=0:4+CHAOS_LONG_OCC=16,32,64 CHAOS_EMULATE_PART=0:2 CHAOS_EMULATE_PART=0:2+CHAOS_LONG_OCC=16,32,64" WORKLOADS="c2" STEPS=10 tools/sweep_env.sh >> gpurun_out/r02j_occ.txt 2>&1
echo done
