#!/bin/bash
# round 2, GPU call 26: strands side by side instead of one after the other -- long kernels at half / quarter occupancy (CHAOS_LOOP_WARPS_PER_SM) x strands
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/quick.py --settings "X=0 CHAOS_LOOP_WARPS_PER_SM=16 CHAOS_LOOP_WARPS_PER_SM=16+CHAOS_STRANDS=3 CHAOS_LOOP_WARPS_PER_SM=16+CHAOS_STRANDS=4 CHAOS_LOOP_WARPS_PER_SM=8+CHAOS_STRANDS=4 CHAOS_LOOP_WARPS_PER_SM=24 CHAOS_LOOP_WARPS_PER_SM=16+CHAOS_TIMELINE=gpurun_out/r02z_tl.txt CHAOS_LOOP_WARPS_PER_SM=16+CHAOS_KERNELS_DIR=tools/variants/ce32" --workloads "c2 c2f32 c2ex2" --steps 8 2>&1 | tee gpurun_out/r02z_quick.txt
cat gpurun_out/r02z_tl.txt
