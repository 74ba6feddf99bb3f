#!/bin/bash
# GPU call 31: the marathon list (CHAOS_EVICT_TRIPS): parity, thresholds, dry/end times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "set_aside or full_lists or shortcuts or engines_agree or exported or full_size_frame" 2>&1 | tail -8 ) > gpurun_out/r03e_pytest.log
cat gpurun_out/r03e_pytest.log
timeout 600 python tools/quick.py --settings "X=0 CHAOS_EVICT_TRIPS=0 CHAOS_EVICT_TRIPS=512 CHAOS_EVICT_TRIPS=2048 CHAOS_EVICT_TRIPS=256 CHAOS_STRANDS=1 CHAOS_STRANDS=1+CHAOS_EVICT_TRIPS=512 CHAOS_STRANDS=3 CHAOS_KERNELS_DIR=tools/variants/ce32" --workloads "c2 c2f32" --steps 8 2>&1 | tee gpurun_out/r03e_quick.txt
( CHAOS_STRANDS=1 LS_FRAMES=2 timeout 300 python tools/lane_stats.py c2 2>&1 | grep -v "pass main" | tail -6 ) 2>&1 | tee gpurun_out/r03e_ls.txt
