#!/bin/bash
# GPU call 64: launch timeline of c4 and c5 frames that go to host memory in parts
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( TINY_HOST=1 TINY_WL=c4 TINY_FRAMES=3 timeout 300 python tools/tiny_timeline.py 8192 8192; TINY_HOST=1 TINY_WL=c5 TINY_FRAMES=5 timeout 300 python tools/tiny_timeline.py 3840 2160 ) 2>&1 | tee gpurun_out/r04k.txt
