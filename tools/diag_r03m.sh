#!/bin/bash
# GPU call 39: latency floor of the long kernel: c2's view at small sizes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CHAOS_STRANDS=1 timeout 300 python tools/tiny_timeline.py 128 72 256 144 640 360 1280 720 2>&1 | grep -v "Classify\|Order\|Export\|PassB\|Replay\|compose" | tee gpurun_out/r03m.txt
