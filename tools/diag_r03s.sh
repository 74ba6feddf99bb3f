#!/bin/bash
# GPU call 45: is the long kernels' tail the warps that hold orbits of both forms (pixels with c.x == 0 exactly)?  c2's view and the same view moved off the axes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in "-0.5,0.0" "-0.50013,0.00007"; do
  echo "== centre $c"
  TINY_CENTER=$c CHAOS_STRANDS=1 timeout 300 python tools/tiny_timeline.py 256 144 3840 2160 2>&1 | grep "frame\|LongDouble\|ProbeDouble"
done 2>&1 | tee gpurun_out/r03s.txt
