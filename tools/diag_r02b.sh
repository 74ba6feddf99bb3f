#!/bin/bash
# round 2, GPU call 2: engine 2 (orbit streams): parity suite, then engine 1 vs 2 on every workload, then knob sweeps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/r02b_pytest.log
SETTINGS="CHAOS_ENGINE=1 CHAOS_ENGINE=2" WORKLOADS="c2 c2ex2 c2f32 c4 c1 c5" STEPS=10 tools/sweep_env.sh > gpurun_out/r02b_engines.txt 2>&1
SETTINGS="CHAOS_PROBE_TRIPS=32 CHAOS_PROBE_TRIPS=128 CHAOS_SCHED_IDLE=2,16 CHAOS_SCHED_IDLE=4,16 CHAOS_SCHED_IDLE=8,16 CHAOS_BLOCK_ITERS=64 CHAOS_BLOCK_ITERS=256 CHAOS_LOOP_WARPS_PER_SM=24 CHAOS_LOOP_WARPS_PER_SM=16 CHAOS_STRANDS=1 CHAOS_STRANDS=3 CHAOS_STRANDS=4" WORKLOADS="c2 c4 c5" STEPS=10 tools/sweep_env.sh > gpurun_out/r02b_knobs.txt 2>&1
echo done
