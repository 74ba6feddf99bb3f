#!/bin/bash
# GPU call 73: memcheck over small frames through the default engine choice and the streams (the lists now carry the orbits' state)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( SANITIZE_ENGINES=2 timeout 95 compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize_small.py 2>&1 | tail -14 ) 2>&1 | tee gpurun_out/r04r_memcheck.txt
