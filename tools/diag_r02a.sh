#!/bin/bash
# round 2, GPU call 1: where c2 stands, occupancy sweep, one rank of an 8-GPU frame on its own
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r02a_pytest.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt
timeout 300 python bench.py --workload c2 --steps 20 > gpurun_out/r02a_bench_c2.json 2> gpurun_out/r02a_bench_c2.err
( timeout 200 python tools/lane_stats.py c2 c2ex2 c5 ) > gpurun_out/r02a_lane_stats.txt 2>&1
SETTINGS="none CHAOS_LOOP_WARPS_PER_SM=24 CHAOS_LOOP_WARPS_PER_SM=16 CHAOS_LOOP_WARPS_PER_SM=12 CHAOS_LOOP_WARPS_PER_SM=8 CHAOS_LOOP_WARPS_PER_SM=16,CHAOS_STRANDS=1 CHAOS_LOOP_WARPS_PER_SM=16,CHAOS_STRANDS=3 CHAOS_LOOP_WARPS_PER_SM=16,CHAOS_POOL_MIN=0" WORKLOADS="c2 c2ex2 c4" STEPS=10 tools/sweep_env.sh > gpurun_out/r02a_sweep_occ.txt 2>&1
SETTINGS="CHAOS_EMULATE_PART=0:8 CHAOS_EMULATE_PART=3:8 CHAOS_EMULATE_PART=0:8,CHAOS_LOOP_WARPS_PER_SM=16 CHAOS_EMULATE_PART=0:8,CHAOS_LOOP_WARPS_PER_SM=8 CHAOS_EMULATE_PART=0:8,CHAOS_STRANDS=1 CHAOS_EMULATE_PART=0:8,CHAOS_STRANDS=1,CHAOS_LOOP_WARPS_PER_SM=8 CHAOS_EMULATE_PART=0:2 CHAOS_EMULATE_PART=0:4" WORKLOADS="c2 c4" STEPS=10 tools/sweep_env.sh > gpurun_out/r02a_sweep_part.txt 2>&1
SETTINGS="none CHAOS_SYNC_BELOW=0 CHAOS_SYNC_BELOW=0,CHAOS_LOOP_WARPS_PER_SM=16" WORKLOADS="c1 c5" STEPS=20 tools/sweep_env.sh > gpurun_out/r02a_sweep_low.txt 2>&1
echo done
