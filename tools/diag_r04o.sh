#!/bin/bash
# GPU call 69: final records of the build as committed: both bench arms, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r04o_smoke.log 2>&1
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r04o_bench_ref.json 2> gpurun_out/r04o_bench_ref.err
( time timeout 900 python bench.py ) > gpurun_out/r04o_bench.json 2> gpurun_out/r04o_bench.err
tail -n 1 gpurun_out/r04o_smoke.log
