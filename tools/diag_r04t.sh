#!/bin/bash
# GPU call 75: memcheck over the frames-in-parts path (4K frames into host memory)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( SANITIZE_ENGINES=3 SANITIZE_PARTS=1 timeout 50 compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize_small.py 2>&1 | tail -12 ) > gpurun_out/r04t_memcheck.txt 2>&1
cat gpurun_out/r04t_memcheck.txt
