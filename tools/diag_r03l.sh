#!/bin/bash
# GPU call 38: the long kernels over time (warps alive, blocks per warp, lanes per block), one strand
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( CHAOS_STRANDS=1 LS_FRAMES=2 timeout 300 python tools/lane_stats.py c2 2>&1 | grep -v "pass main" | tail -75 ) 2>&1 | tee gpurun_out/r03l_ls.txt
