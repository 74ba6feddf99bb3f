#!/bin/bash
# GPU call 34: how long the orbits that run all the way take and when the last one starts (lane-stats build), with and without the length guess
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for s in "CHAOS_PREDICT=1" "CHAOS_PREDICT=0" "CHAOS_PREDICT=1 CHAOS_LOOP_WARPS_PER_SM=16"; do
  echo "== $s, one strand"
  ( env $s CHAOS_STRANDS=1 LS_FRAMES=2 timeout 300 python tools/lane_stats.py c2 2>&1 | grep -v "pass main" | tail -8 )
done 2>&1 | tee gpurun_out/r03h_ls.txt
