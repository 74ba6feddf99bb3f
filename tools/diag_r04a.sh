#!/bin/bash
# GPU call 53: records of the build as committed: both bench arms (default command), c2 FP32 and c2 "M ex 2" end to end, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r04a_bench_ref.json 2> gpurun_out/r04a_bench_ref.err
( time timeout 900 python bench.py ) > gpurun_out/r04a_bench.json 2> gpurun_out/r04a_bench.err
B="python bench.py --no-extras --no-cpu-baseline --no-full-trips"
timeout 300 $B --workload c2f32 > gpurun_out/r04a_c2f32.json 2> /dev/null
timeout 300 $B --workload c2ex2 > gpurun_out/r04a_c2ex2.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03_launches_c2.csv $B --workload c2 --steps 2 --warmup 3 > /dev/null 2>&1
tail -n 3 gpurun_out/r04a_bench.err
