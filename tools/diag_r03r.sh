#!/bin/bash
# GPU call 44: tiny frame, no pool, lane-stats build: when do the orbits that run all the way start?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( CHAOS_POOL_MIN=0 CHAOS_STRANDS=1 LS_FRAMES=2 timeout 300 python tools/lane_stats.py c2@256x144 2>&1 | tail -40 ) 2>&1 | tee gpurun_out/r03r.txt
