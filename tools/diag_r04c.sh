#!/bin/bash
# GPU call 55: records of the build as committed (cap only for frames that stay on the device): both bench arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r04c_bench_ref.json 2> gpurun_out/r04c_bench_ref.err
( time timeout 900 python bench.py ) > gpurun_out/r04c_bench.json 2> gpurun_out/r04c_bench.err
B="python bench.py --no-extras --no-cpu-baseline --no-full-trips"
timeout 300 $B --workload c2f32 > gpurun_out/r04c_c2f32.json 2> /dev/null
timeout 300 $B --workload c2ex2 > gpurun_out/r04c_c2ex2.json 2> /dev/null
tail -n 3 gpurun_out/r04c_bench.err
