#!/bin/bash
# GPU call 36: pool looks backed off in the drain (CHAOS_DRAIN_LOOK_MAX 64 / 8 / 2 = as before), with and without the length guess
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/quick.py --settings "X=0 CHAOS_KERNELS_DIR=tools/variants/look8 CHAOS_KERNELS_DIR=tools/variants/look2 CHAOS_PREDICT=0 CHAOS_STRANDS=1 CHAOS_STRANDS=1+CHAOS_KERNELS_DIR=tools/variants/look2" --workloads "c2 c2f32" --steps 8 2>&1 | tee gpurun_out/r03j_quick.txt
( CHAOS_STRANDS=1 LS_FRAMES=2 timeout 300 python tools/lane_stats.py c2 2>&1 | grep -v "pass main" | tail -8 ) 2>&1 | tee gpurun_out/r03j_ls.txt
