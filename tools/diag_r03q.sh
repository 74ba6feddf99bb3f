#!/bin/bash
# GPU call 43: is the 0.4 ms of a tiny frame's long kernel a matter of clocks?  3000 frames back to back, timeline of the last; SM clock sampled meanwhile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader -lms 500 > gpurun_out/r03q_clocks.txt & echo $! > /tmp/smi.pid )
TINY_FRAMES=3000 CHAOS_STRANDS=1 timeout 300 python tools/tiny_timeline.py 256 144 2>&1 | grep "frame\|LongDouble" | tee gpurun_out/r03q.txt
TINY_FRAMES=3000 CHAOS_POOL_MIN=0 CHAOS_STRANDS=1 timeout 300 python tools/tiny_timeline.py 256 144 2>&1 | grep "frame\|LongDouble" | tee -a gpurun_out/r03q.txt
kill $(cat /tmp/smi.pid)
sort gpurun_out/r03q_clocks.txt | uniq -c | sort -rn | head -8
