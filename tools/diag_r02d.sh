#!/bin/bash
# round 2, GPU call 4: engine 2 with the orbit pool; lane counters and source-level profile of the long and probe kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "engine2 or engines_agree or strands or exported" 2>&1 | tail -15 ) > gpurun_out/r02d_pytest.log
( CHAOS_STRANDS=1 timeout 200 python tools/lane_stats.py c2 c2ex2 ) > gpurun_out/r02d_lane_stats.txt 2>&1
SETTINGS="X=0 CHAOS_STRANDS=1 CHAOS_POOL_MIN=0 CHAOS_POOL_MIN=0+CHAOS_STRANDS=1 CHAOS_POOL_MIN=28 CHAOS_POOL_MIN=12" WORKLOADS="c2 c4 c5 c2ex2" STEPS=10 tools/sweep_env.sh > gpurun_out/r02d_knobs.txt 2>&1
CHAOS_STRANDS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"chaosLongDouble|chaosProbeDouble" --launch-skip 12 --launch-count 2 -o gpurun_out/r02d_long_probe_c2 -f python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-full-trips > /dev/null 2>&1
echo done
