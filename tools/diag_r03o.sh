#!/bin/bash
# GPU call 41: cycles per trip inside the long kernel's loop, full frame and a tiny one (lane-stats build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( CHAOS_STRANDS=1 LS_FRAMES=2 timeout 300 python tools/lane_stats.py c2@256x144 c2 2>&1 | grep "cycles per trip\|render_ms\|dry after" ) 2>&1 | tee gpurun_out/r03o.txt
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I chaos-ultra_b200/csrc tools/loopbench.cu -o /tmp/loopbench 2>&1 | tail -2
timeout 120 /tmp/loopbench 2>&1 | tail -30 | tee -a gpurun_out/r03o.txt
