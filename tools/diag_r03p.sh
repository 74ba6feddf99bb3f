#!/bin/bash
# GPU call 42: cycles per iteration of the long kernel's main loop once the list is dry, with and without a pass (lane-stats build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( CHAOS_STRANDS=1 LS_FRAMES=2 timeout 300 python tools/lane_stats.py c2@256x144 c2 2>&1 | grep "cycles\|render_ms\|dry after" | tail -14 ) 2>&1 | tee gpurun_out/r03p.txt
( CHAOS_POOL_MIN=0 CHAOS_STRANDS=1 LS_FRAMES=2 timeout 300 python tools/lane_stats.py c2@256x144 2>&1 | grep "cycles\|render_ms\|dry after" | tail -6 ) 2>&1 | tee -a gpurun_out/r03p.txt
