#!/bin/bash
# GPU call 28: lane counters and the long kernels' dry/end times on c2 (modules built with -DCHAOS_LANE_STATS), one strand and two
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( CHAOS_STRANDS=1 timeout 300 python tools/lane_stats.py c2 ) > gpurun_out/r03b_ls1.txt 2>&1
( timeout 300 python tools/lane_stats.py c2 ) > gpurun_out/r03b_ls2.txt 2>&1
cat gpurun_out/r03b_ls1.txt gpurun_out/r03b_ls2.txt
