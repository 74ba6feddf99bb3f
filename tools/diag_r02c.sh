#!/bin/bash
# round 2, GPU call 3: engine 2 after the counter fix: parity, per-kernel breakdown of a c2 frame, refill policy sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/r02c_pytest.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size
for w in c2 c5; do
CHAOS_STRANDS=1 timeout 600 ncu --metrics $M --clock-control none --launch-skip 36 --launch-count 12 --csv --log-file gpurun_out/r02c_ncu_$w.csv python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-full-trips > /dev/null 2>&1
done
CHAOS_ENGINE=1 CHAOS_STRANDS=1 timeout 600 ncu --metrics $M --clock-control none --launch-skip 24 --launch-count 8 --csv --log-file gpurun_out/r02c_ncu_c2_e1.csv python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-full-trips > /dev/null 2>&1
SETTINGS="X=0 CHAOS_STRANDS=1 CHAOS_SCHED_IDLE=2,16 CHAOS_SCHED_IDLE=5,16 CHAOS_SCHED_IDLE=20,16 CHAOS_SCHED_IDLE=5,16+CHAOS_STRANDS=1 CHAOS_SCHED_IDLE=5,16+CHAOS_BLOCK_ITERS=64 CHAOS_SCHED_IDLE=5,16+CHAOS_BLOCK_ITERS=32" WORKLOADS="c2 c4 c5 c2ex2" STEPS=10 tools/sweep_env.sh > gpurun_out/r02c_knobs.txt 2>&1
echo done
