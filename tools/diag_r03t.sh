#!/bin/bash
# GPU call 46: one rank of an 8-GPU c2 frame, on and off the axes: how much of its chain is the warps that hold orbits of both forms?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in "-0.5,0.0" "-0.50013,0.00007"; do
  for p in 0:8 3:8; do
  echo "== centre $c rank $p"
  TINY_PART=$p TINY_CENTER=$c TINY_FRAMES=6 timeout 300 python tools/tiny_timeline.py 3840 2160 2>&1 | grep "frame\|LongDouble\|ProbeDouble\|compose\|Finish"
  done
done 2>&1 | tee gpurun_out/r03t.txt
