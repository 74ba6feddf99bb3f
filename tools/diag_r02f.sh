#!/bin/bash
# round 2, GPU call 6: lock-step voted blocks in the long kernel, hot-first pass C; workload constants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "engine2 or engines_agree or strands or exported or shortcuts or full_size" 2>&1 | tail -15 ) > gpurun_out/r02f_pytest.log
( CHAOS_STRANDS=1 timeout 200 python tools/lane_stats.py c2 c2ex2 ) > gpurun_out/r02f_lane_stats.txt 2>&1
SETTINGS="X=0 CHAOS_ENGINE=1 CHAOS_HOT_FIRST=0 CHAOS_KERNELS_DIR=tools/variants/mb3 CHAOS_PASS_THREADS=128 CHAOS_PASS_THREADS=128+CHAOS_STRANDS=3 CHAOS_LOOP_WARPS_PER_SM=24 CHAOS_SCHED_IDLE=5,16 CHAOS_SCHED_IDLE=20,16 CHAOS_BLOCK_ITERS=256" WORKLOADS="c2 c2ex2 c4 c5 c1" STEPS=10 tools/sweep_env.sh > gpurun_out/r02f_knobs.txt 2>&1
( timeout 600 python tests/golden/make_workload_constants.py 2>&1 | tail -12 ) > gpurun_out/r02f_constants.log
cp tests/golden/workloads.json gpurun_out/workloads.json 2>/dev/null
echo done
