#!/bin/bash
# GPU call 40: latency floor of the long kernel at 256x144 by pool / block knobs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for s in "X=0" "CHAOS_POOL_MIN=0" "CHAOS_BLOCK_ITERS=512" "CHAOS_POOL_MIN=0 CHAOS_BLOCK_ITERS=1024" "CHAOS_DENSE_COMPARE=0 CHAOS_POOL_MIN=0" "CHAOS_SHORTCUTS=1 CHAOS_POOL_MIN=0" "CHAOS_LONG_OCC=0,0,0"; do
  echo "== $s"
  env $s CHAOS_STRANDS=1 timeout 300 python tools/tiny_timeline.py 256 144 2>&1 | grep "frame\|LongDouble"
done 2>&1 | tee gpurun_out/r03n.txt
