#!/bin/bash
# round 2, GPU call 20: compute-sanitizer memcheck over every kernel family (small frames)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py 2>&1 | tail -40 ) > gpurun_out/r02t_memcheck.log
echo done
