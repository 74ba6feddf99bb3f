#!/bin/bash
# GPU call 48: records of the build as it stands: both bench arms (default command), then the ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r03v_bench_ref.json 2> gpurun_out/r03v_bench_ref.err
( time timeout 900 python bench.py ) > gpurun_out/r03v_bench.json 2> gpurun_out/r03v_bench.err
tail -3 gpurun_out/r03v_bench.err gpurun_out/r03v_bench_ref.err
tools/profile_r03.sh
ls -la gpurun_out/r03_*
