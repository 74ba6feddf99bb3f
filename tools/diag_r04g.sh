#!/bin/bash
# GPU call 60: one rank of an 8-GPU (and 4-GPU) c2 frame: length guess (CHAOS_PREDICT 0 / 2) x long-kernel occupancy thresholds (CHAOS_LONG_OCC)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for p in 0:8 3:8 1:4; do
for s in "CHAOS_PREDICT=0" "CHAOS_PREDICT=2" "CHAOS_PREDICT=2 CHAOS_LONG_OCC=4,8,16" "CHAOS_PREDICT=2 CHAOS_LONG_OCC=8,16,32" "CHAOS_PREDICT=0 CHAOS_LONG_OCC=8,16,32"; do
  echo "== rank $p $s"
  env $s TINY_PART=$p TINY_FRAMES=8 timeout 300 python tools/tiny_timeline.py 3840 2160 2>&1 | grep "frame\|LongDouble\|ProbeDouble"
done; done 2>&1 | tee gpurun_out/r04g.txt
