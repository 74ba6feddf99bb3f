#!/bin/bash
# round 2, GPU call 9: what one rank of an N-GPU c2 / c4 frame costs (partition emulated on one GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SETTINGS="CHAOS_EMULATE_PART=0:2 CHAOS_EMULATE_PART=0:4 CHAOS_EMULATE_PART=0:8 CHAOS_EMULATE_PART=3:8 CHAOS_EMULATE_PART=0:8+CHAOS_STRANDS=1 CHAOS_EMULATE_PART=0:4+CHAOS_STRANDS=1 CHAOS_EMULATE_PART=0:8+CHAOS_BENCH_BAND_ROWS=16 CHAOS_EMULATE_PART=0:8+CHAOS_BENCH_BAND_ROWS=64 CHAOS_EMULATE_PART=0:8+CHAOS_POOL_MIN=0 CHAOS_EMULATE_PART=0:8+CHAOS_ENGINE=1 CHAOS_EMULATE_PART=0:8+CHAOS_LOOP_WARPS_PER_SM=16 CHAOS_EMULATE_PART=0:8+CHAOS_LOOP_WARPS_PER_SM=8+CHAOS_STRANDS=1" WORKLOADS="c2" STEPS=10 tools/sweep_env.sh > gpurun_out/r02i_part.txt 2>&1
M=gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,launch__grid_size
CHAOS_EMULATE_PART=0:8 timeout 600 ncu --metrics $M --clock-control none --launch-skip 72 --launch-count 24 --csv --log-file gpurun_out/r02i_ncu_c2_part8.csv python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-full-trips --no-extras > /dev/null 2>&1
echo done
