#!/bin/bash
# GPU call 74: memcheck (all engines) and racecheck (streams) over small frames on the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 40 compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize_small.py 2>&1 | tail -22 ) > gpurun_out/r04s_memcheck.txt 2>&1
( SANITIZE_ENGINES=2,3 timeout 45 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize_small.py 2>&1 | tail -16 ) > gpurun_out/r04s_racecheck.txt 2>&1
tail -3 gpurun_out/r04s_memcheck.txt gpurun_out/r04s_racecheck.txt
