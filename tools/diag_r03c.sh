#!/bin/bash
# GPU call 29: the long kernels' dry/end times on c2; trips per block of the long kernel (CHAOS_BLOCK_ITERS)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( CHAOS_STRANDS=1 timeout 300 python tools/lane_stats.py c2 2>&1 | grep -v "pass main" ) > gpurun_out/r03c_ls1.txt 2>&1
( timeout 300 python tools/lane_stats.py c2 2>&1 | grep -v "pass main" ) > gpurun_out/r03c_ls2.txt 2>&1
cat gpurun_out/r03c_ls1.txt gpurun_out/r03c_ls2.txt
timeout 600 python tools/quick.py --settings "X=0 CHAOS_BLOCK_ITERS=64 CHAOS_BLOCK_ITERS=32 CHAOS_BLOCK_ITERS=64+CHAOS_SCHED_IDLE=5,16 CHAOS_BLOCK_ITERS=32+CHAOS_SCHED_IDLE=5,16" --workloads "c2 c2f32 c2ex2" --steps 8 2>&1 | tee gpurun_out/r03c_quick.txt
