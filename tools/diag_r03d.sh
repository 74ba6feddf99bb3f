#!/bin/bash
# GPU call 30: the long kernel's drain by occupancy: one strand, 8 / 16 / 32 warps per SM (lane-stats build prints dry / end times)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in 8 16 32; do
  echo "== CHAOS_LOOP_WARPS_PER_SM=$w one strand"
  ( CHAOS_STRANDS=1 CHAOS_LOOP_WARPS_PER_SM=$w LS_FRAMES=2 timeout 300 python tools/lane_stats.py c2 2>&1 | grep -v "pass main" | tail -6 )
done 2>&1 | tee gpurun_out/r03d_ls.txt
timeout 600 python tools/quick.py --settings "CHAOS_STRANDS=1 CHAOS_STRANDS=1+CHAOS_LOOP_WARPS_PER_SM=16 CHAOS_STRANDS=1+CHAOS_LOOP_WARPS_PER_SM=8 CHAOS_STRANDS=1+CHAOS_LONG_OCC=0,0,0 CHAOS_STRANDS=1+CHAOS_POOL_MIN=28 CHAOS_STRANDS=1+CHAOS_POOL_MIN=0" --workloads "c2" --steps 8 2>&1 | tee gpurun_out/r03d_quick.txt
