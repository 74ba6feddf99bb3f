#!/bin/bash
# round 2, GPU call 5: static probe; group-size variants; multi-process frame assembly tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r02e_pytest.log
SETTINGS="X=0 CHAOS_STRANDS=1 CHAOS_KERNELS_DIR=tools/variants/g64 CHAOS_KERNELS_DIR=tools/variants/g32u2 CHAOS_KERNELS_DIR=tools/variants/g64u2 CHAOS_KERNELS_DIR=tools/variants/g16 CHAOS_PASS_THREADS=128 CHAOS_PASS_THREADS=64 CHAOS_STRANDS=3" WORKLOADS="c2 c4 c5 c2ex2 c1" STEPS=10 tools/sweep_env.sh > gpurun_out/r02e_knobs.txt 2>&1
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size
CHAOS_STRANDS=1 timeout 600 ncu --metrics $M --clock-control none --launch-skip 36 --launch-count 12 --csv --log-file gpurun_out/r02e_ncu_c2.csv python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-full-trips > /dev/null 2>&1
echo done
